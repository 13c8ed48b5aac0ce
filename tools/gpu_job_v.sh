#!/bin/bash
mkdir -p gpurun_out
for wl in single dual; do
timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_v_$wl.json 2> gpurun_out/bench_v_$wl.err; echo "rc=$?"; tail -3 gpurun_out/bench_v_$wl.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_v_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='mode'})"
done
