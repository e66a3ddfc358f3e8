#!/bin/bash
# N=1 recompute-policy sweep on the full Llama-3-8B bench (memory vs. recompute FLOPs).
mkdir -p gpurun_out
for r in swiglu_norm mlp_only swiglu; do
  timeout 240 python bench.py --steps 2 --warmup 3 --recompute $r > gpurun_out/bench_n1_$r.log 2>&1
  echo "== $r exit $?"; grep -h '^{"metric"' gpurun_out/bench_n1_$r.log | cut -c1-400; tail -2 gpurun_out/bench_n1_$r.log | cut -c1-300
done
