#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxelize_gpu.py tests/test_depth.py tests/test_act.py -m gpu -q -x > gpurun_out/pytest_p.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_p.log | cut -c1-300
python tools/vox_time.py
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_p.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_p.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value']); print(d['roofline']['voxelize'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vox_ -s 8 -c 4 -o gpurun_out/ncu_r02_voxelize_e python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-gpu-comparator > gpurun_out/ncu_q.log 2>&1; echo "rc=$?"
