#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file gpurun_out/launches_native_4L.csv python bench.py --impl native --layers 4 --steps 1 --warmup 1 --global-batch 1 --no-e2e --recompute none > gpurun_out/ncu_launches.log 2>&1; echo "exit $?" >> gpurun_out/ncu_launches.log
tail -n 2 gpurun_out/ncu_launches.log | cut -c1-300
wc -l gpurun_out/launches_native_4L.csv
