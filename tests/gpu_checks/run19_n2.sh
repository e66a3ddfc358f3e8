#!/bin/bash
# 2-GPU validation of the device-side epoch + CUDA-graph replay of the micro-batch with fused TP kernels.
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=${N:-2}
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dropout or embedding or eager" > gpurun_out/pytest_gpu_r19.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r19.log
FUSED_TIMELINE=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/gpu_checks/check_fused_tp.py > gpurun_out/check_fused_tp_n${N}_dev_epoch.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_tp_n${N}_dev_epoch.log
grep -E "^AG|^GEMM|FUSED|^exit|Error" gpurun_out/check_fused_tp_n${N}_dev_epoch.log
for g in 1 0; do
  NXDT_CUDA_GRAPH_DEBUG=1 NXDT_CUDA_GRAPH=$g timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$g bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_graph$g.log 2>&1; echo "exit $?" >> gpurun_out/bench_n${N}_graph$g.log
  echo "== N=$N graph=$g"; grep -h '^{"metric"' gpurun_out/bench_n${N}_graph$g.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tokens/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 1), 'e2e', round(d.get('e2e', {}).get('value', 0)), 'loss', d['config'].get('final_loss'), 'mem', d['config'].get('peak_mem_gb'), 'launches', d.get('gpu_launches'), d['clocks'])"
  grep -E "CUDA-graph|captured|Error|error" gpurun_out/bench_n${N}_graph$g.log | head -5
  tail -n 2 gpurun_out/bench_n${N}_graph$g.log | cut -c1-300
done
