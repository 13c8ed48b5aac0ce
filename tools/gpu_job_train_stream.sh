#!/bin/bash
# training backward check: gradient tests + agent-level update test, then the training bench line (1 GPU, batch 16)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_train_ops.py tests/test_agent_dropin.py -m gpu -q -x > gpurun_out/pytest_train.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_train.log | cut -c1-300
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|FAILED" gpurun_out/pytest_train.log | head -20 | cut -c1-300; exit 1; fi
timeout 300 python bench.py --workload train --batch 16 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02_train_1gpu_stream2.json 2> gpurun_out/bench_train_stream.err; echo "train rc=$?"; tail -3 gpurun_out/bench_train_stream.err | cut -c1-300; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_train_1gpu_stream2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('stages_ms'))"
