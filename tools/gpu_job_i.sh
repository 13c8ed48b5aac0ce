#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_depth.py tests/test_voxelize_gpu.py tests/test_act.py -m gpu -q -x > gpurun_out/pytest_i.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_i.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vox_ -s 4 -c 2 -o gpurun_out/ncu_r02_voxelize_b \
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1
