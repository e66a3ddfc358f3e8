#!/bin/bash
# First GPU call of the next round (1 GPU, ~4 GPU-minutes): everything that was written after round 1's GPU budget was
# spent, in priority order, each under its own timeout so one hang cannot eat the call.
#   gpurun --timeout 900 -- 'bash tests/gpu_checks/run_round2_first.sh'
mkdir -p gpurun_out
run() {  # name, timeout, command...
  name=$1; t=$2; shift 2
  timeout "$t" "$@" > gpurun_out/r2_$name.log 2>&1
  echo "== $name exit $? : $(grep -E '_OK|rel err|TFLOP|TB/s|passed|failed|Error|error|assert' gpurun_out/r2_$name.log | tail -4 | cut -c1-220 | tr '\n' '|')"
}
run pytest_gpu 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu
run flash_bwd 120 python tests/gpu_checks/check_flash_bwd.py
run flash_fwd 120 python tests/gpu_checks/check_flash.py
NXDT_FLASH_SPLIT=2 run flash_fwd_split2 120 python tests/gpu_checks/check_flash.py      # eight-softmax-warp variant
run fp8 180 python tests/gpu_checks/check_fp8.py
run adam_v1 120 python tests/gpu_checks/check_adam.py
NXDT_ADAM_V2=1 run adam_v2 120 python tests/gpu_checks/check_adam.py
NXDT_ADAM_V2=1 run adam_v2_numerics 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "adamw_flat"
run smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
NXDT_FLASH_BWD_PREVIEW=1 run bench_flash_native 200 python bench.py --steps 2 --warmup 3 --layers 8 --no-e2e
run bench_default_8L 200 python bench.py --steps 2 --warmup 3 --layers 8 --no-e2e
grep -h '^{"metric"' gpurun_out/r2_bench_flash_native.log gpurun_out/r2_bench_default_8L.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tokens/s', round(d['value']), 'loss', d['config'].get('final_loss'))"
