#!/bin/bash
mkdir -p gpurun_out
for m in 1 2; do
VXB_TRANSFORMER_F8C=$m timeout 900 python tools/report_errors.py > gpurun_out/errors_u$m.log 2>&1; echo "mask=$m rc=$?"; grep "mode 2" gpurun_out/errors_u$m.log | cut -c1-250
VXB_TRANSFORMER_F8C=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_u$m.json 2> gpurun_out/bench_u$m.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_u$m.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['stages_ms']['transformer'])"
done
