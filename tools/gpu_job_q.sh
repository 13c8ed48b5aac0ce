#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vox_ -s 9 -c 3 -o gpurun_out/ncu_r02_voxelize_c python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-gpu-comparator > gpurun_out/ncu_q.log 2>&1; echo "rc=$?"
