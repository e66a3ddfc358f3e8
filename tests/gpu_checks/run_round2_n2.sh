#!/bin/bash
# Second GPU call of the next round (2 GPUs, ~8 GPU-minutes): graph replay + fused TP reproducer and the two overlap modes.
#   gpurun --gpus 2 --timeout 900 -- 'bash tests/gpu_checks/run_round2_n2.sh'
mkdir -p gpurun_out
export NCCL_DEBUG=WARN PYTHONFAULTHANDLER=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; t=$2; shift 2; timeout "$t" "$@" > gpurun_out/r2n2_$name.log 2>&1; echo "== $name exit $? : $(grep -E '_OK|replay|nxdt\]|tokens/s|Error' gpurun_out/r2n2_$name.log | tail -3 | cut -c1-200 | tr '\n' '|')"; }
run graph_ag_rs 60 $TR --master-port 29701 tests/gpu_checks/check_fused_graph.py --pattern ag,rs --replays 40
NXDT_RS_CLEAR_KERNEL=1 run graph_ag_rs_clearkernel 60 $TR --master-port 29702 tests/gpu_checks/check_fused_graph.py --pattern ag,rs --replays 40
NXDT_RS_SELF_RESET=1 run graph_ag_rs_selfreset 60 $TR --master-port 29705 tests/gpu_checks/check_fused_graph.py --pattern ag,rs --replays 40
NXDT_RS_ARRIVE_FLAGS=1 run graph_ag_rs_flags 60 $TR --master-port 29707 tests/gpu_checks/check_fused_graph.py --pattern ag,rs --replays 40
NXDT_RS_ARRIVE_FLAGS=1 run eager_flags_check 120 $TR --master-port 29708 tests/gpu_checks/check_fused_tp.py
NXDT_RS_BLOCKED=1 run eager_blocked_check 120 $TR --master-port 29709 tests/gpu_checks/check_fused_tp.py     # 16 KB contiguous peer writes
NXDT_RS_BLOCKED=1 NXDT_RS_ARRIVE_FLAGS=1 run eager_blocked_flags_check 120 $TR --master-port 29710 tests/gpu_checks/check_fused_tp.py
NXDT_RS_SELF_RESET=1 run eager_selfreset_check 120 $TR --master-port 29706 tests/gpu_checks/check_fused_tp.py
run graph_rs_only 60 $TR --master-port 29703 tests/gpu_checks/check_fused_graph.py --pattern rs,rs --replays 40
run graph_ag_only 60 $TR --master-port 29704 tests/gpu_checks/check_fused_graph.py --pattern ag,ag --replays 40
port=29711
for mode in default overlap keepgather; do
  port=$((port + 1))
  env $( [ $mode = overlap ] && echo NXDT_MB_OVERLAP=1 ) $( [ $mode = keepgather ] && echo NXDT_KEEP_GATHERED=1 ) timeout 200 $TR --master-port $port bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2n2_bench_$mode.log 2>&1
  echo "== bench $mode exit $?"; grep -h '^{"metric"' gpurun_out/r2n2_bench_$mode.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tokens/s', round(d['value']), 'loss', d['config'].get('final_loss'), 'overlap', d['config'].get('microbatch_overlap'))"
done
