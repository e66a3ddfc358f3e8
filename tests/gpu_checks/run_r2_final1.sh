#!/bin/bash
# round 2, final 1-GPU call: launch list of the 4-layer Llama-3-8B step (ncu, duration only), compute-sanitizer, smoke()
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r2_4L.csv \
  python bench.py --layers 4 --steps 1 --warmup 1 --no-e2e > gpurun_out/final1_launch.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_r2_4L.csv
bash tests/gpu_checks/run_r2_sanitize.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 3
