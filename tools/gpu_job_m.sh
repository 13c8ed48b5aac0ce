#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_f8c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --math f16f8c > gpurun_out/ncu_m.log 2>&1; echo "rc=$?"
