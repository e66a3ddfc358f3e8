#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --fp8 --no-e2e 2>&1 | grep -E '^\{"metric"|Error|error' | cut -c1-1600
