#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py tests/test_train_ops.py -m gpu -q > gpurun_out/pytest_x.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_x.log | cut -c1-300
timeout 900 python bench.py --workload train --batch 16 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_x_train.json 2> gpurun_out/bench_x_train.err; echo "rc=$?"; tail -3 gpurun_out/bench_x_train.err | cut -c1-300; python -c "
import json
d=json.loads(open('gpurun_out/bench_x_train.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['stages_ms'])"
