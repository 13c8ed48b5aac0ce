#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_r02_train.csv python bench.py --workload train --batch 16 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_y.log 2>&1; echo "rc=$?"
