#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02_n.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r02_n.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_r02_n.json 2> gpurun_out/bench_r02_n.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r02_n.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_n.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'], d.get('cpu_baseline'))"
