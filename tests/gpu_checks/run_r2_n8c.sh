#!/bin/bash
# round 2, N=8 call C: headline config with the round's final defaults (SM-pair GEMM, keep_gathered auto) and the NVLS norm + all-gather A/B
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
    bench.py --gpus 8 --steps 3 --warmup 3 "$@" > gpurun_out/n8c_$name.log 2>&1
  echo "== $name exit $?"; grep '^{"metric' gpurun_out/n8c_$name.log | tail -1 | cut -c1-1500; grep "NVLS\|nvls" gpurun_out/n8c_$name.log | head -3
}
run default --no-e2e
run nvls --nvls --no-e2e
