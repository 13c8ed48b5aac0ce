#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-comparator > gpurun_out/ncu_m.log 2>&1; echo "rc=$?"
