#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -s > gpurun_out/pytest_train_r02.log 2>&1; echo "train rc=$?" >> gpurun_out/pytest_train_r02.log
grep -n "vs fp64\|activation-gradient\|passed\|failed\|^FAILED\|Error" gpurun_out/pytest_train_r02.log | cut -c1-180 | head -70
timeout 900 python bench.py --workload train --batch 16 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_train_b16.json 2> gpurun_out/bench_train_b16.err; echo "rc=$?"; tail -c 900 gpurun_out/bench_train_b16.json; tail -5 gpurun_out/bench_train_b16.err
