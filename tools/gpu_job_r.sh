#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_ops_gpu.py tests/test_qnet_gpu.py tests/test_act.py -m gpu -q -x > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r.log | cut -c1-300
for wl in single; do
timeout 600 python bench.py --steps 20 --warmup 5 --batch 1 --workload $wl --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_r_b1_$wl.json 2> gpurun_out/bench_r_b1_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_r_b1_$wl.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r_b1_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('act_latency')); print(d['stages_ms'])"
done
timeout 600 python bench.py --steps 10 --warmup 3 --batch 4 --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_r_b4.json 2> gpurun_out/bench_r_b4.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r_b4.json').read().strip().splitlines()[-1])
print('B=4', d['value'], d['ms_per_step'], d['e2e']['value'])"
