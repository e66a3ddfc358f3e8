#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k whole_model 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_kernels_gpu.py::test_whole_model_native_path_matches_library_path 2>&1 | tail -5 | cut -c1-300
