#!/bin/bash
mkdir -p gpurun_out
for sk in 0 4 8 0; do
VXB_CONV_DEBUG_SKIP=$sk timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/bench_k_sk$sk.json 2> gpurun_out/bench_k_sk$sk.err; echo "skip=$sk rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_k_sk$sk.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['stages_ms']['final_conv'], d['clocks'])"
done
