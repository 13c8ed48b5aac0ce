#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/report_errors.py > gpurun_out/errors_u.log 2>&1; echo "rc=$?"; grep "mode 2" gpurun_out/errors_u.log | cut -c1-250; tail -3 gpurun_out/errors_u.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; echo "rc=$?"; tail -3 gpurun_out/bench_u.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_u.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms'])"
