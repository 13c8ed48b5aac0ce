#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/report_errors.py > gpurun_out/errors_l.log 2>&1; echo "rc=$?"; grep "mode 2" gpurun_out/errors_l.log | cut -c1-250; tail -3 gpurun_out/errors_l.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --math f16f8c > gpurun_out/bench_l_f8c.json 2> gpurun_out/bench_l_f8c.err; echo "rc=$?"; tail -3 gpurun_out/bench_l_f8c.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_l_f8c.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'])"
