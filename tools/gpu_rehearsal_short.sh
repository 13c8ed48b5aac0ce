#!/bin/bash
# short round-end check on the final binary: GPU tests, default bench line, smoke
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02_final.log | cut -c1-200
timeout 200 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err | cut -c1-200; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='mode'}); print(d['roofline']['frac'], d['roofline']['whole_forward_frac'], d['roofline']['voxelize']); print(d['clocks']); print(d.get('gpu_torch_comparator')); print(d.get('cpu_baseline'))"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02.log | cut -c1-300
