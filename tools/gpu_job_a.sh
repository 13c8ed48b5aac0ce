#!/bin/bash
# round-2 job A: full GPU test-suite, default bench, ncu launch list + full captures of the voxelizer kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_a.log
tail -5 gpurun_out/pytest_gpu_r02_a.log
timeout 600 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -c 1500 gpurun_out/bench_r02_a.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vox_ -s 4 -c 4 -o gpurun_out/ncu_r02_voxelize \
  python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_a.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out
