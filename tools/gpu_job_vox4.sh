#!/bin/bash
# voxelizer v4 (fused scatter + background launch, table-scan patch): bit-exact tests, A/B timing against v3, ncu of the new kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_voxelize_gpu.py tests/test_depth.py -m gpu -q -x > gpurun_out/pytest_vox4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_vox4.log | cut -c1-200
{
for p in 3 4 5; do echo "== v4 fill_per_sm=$p"; VXB_VOX_FILL_PER_SM=$p timeout 200 python tools/vox_time.py 2>&1 | tail -4; done
echo "== v3"; VXB_VOX_PATH=v3 timeout 200 python tools/vox_time.py 2>&1 | tail -4
} | tee gpurun_out/vox_time_r02_v4.log
for k in vox_scatter_fill vox_patch; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_${k}_v4 python tools/vox_time.py > /dev/null 2>&1; echo "$k rc=$?"
done
