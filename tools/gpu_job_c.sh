#!/bin/bash
mkdir -p gpurun_out
for b in 2 16; do
  timeout 900 python bench.py --workload train --batch $b --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_train_b$b.json 2> gpurun_out/bench_train_b$b.err; echo "rc=$?"; tail -c 1800 gpurun_out/bench_train_b$b.json; tail -5 gpurun_out/bench_train_b$b.err
done
nvidia-smi --query-gpu=memory.used --format=csv
