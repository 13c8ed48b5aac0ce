#!/bin/bash
# batch-1 latency line and the training step on the final binary
mkdir -p gpurun_out
timeout 600 python bench.py --batch 1 --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_r02_batch1.json 2> gpurun_out/bench_b1.err; echo "b1 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_batch1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('act_latency')); print(d['stages_ms'])"
timeout 900 python bench.py --workload train --steps 5 --warmup 2 > gpurun_out/bench_r02_train_1gpu_final.json 2> gpurun_out/bench_train.err; echo "train rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_train_1gpu_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('stages_ms'))"
