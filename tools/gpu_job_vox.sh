#!/bin/bash
# voxelizer check: bit-exact tests, timing by occupancy (tools/vox_time.py), ncu --set full of both kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxelize_gpu.py tests/test_depth.py -m gpu -q -x > gpurun_out/pytest_vox.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_vox.log | cut -c1-200
timeout 300 python tools/vox_time.py 2>&1 | tail -4 | tee gpurun_out/vox_time_r02_final.log
for k in vox_scatter vox_fill_rows; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_${k}_v3 python tools/vox_time.py > /dev/null 2>&1; echo "$k rc=$?"
done
ls -la gpurun_out/*_v3.ncu-rep
