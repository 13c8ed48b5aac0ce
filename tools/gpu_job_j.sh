#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_qnet_gpu.py -m gpu -q -x -k "f16_fp8 or golden or tracks or full_size" > gpurun_out/pytest_j.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_j.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "rc=$?"; tail -3 gpurun_out/bench_j.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'], d['clocks'])"
