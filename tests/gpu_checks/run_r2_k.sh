#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | cut -c1-1800
