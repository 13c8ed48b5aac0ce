#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -s > gpurun_out/pytest_train_r02.log 2>&1; echo "train rc=$?" >> gpurun_out/pytest_train_r02.log
grep -n "vs fp64\|activation-gradient\|passed\|failed\|^FAILED\|Error" gpurun_out/pytest_train_r02.log | head -90
timeout 1200 python -m pytest tests/test_qnet_gpu.py -m gpu -q -s -k "baseline_configs or two_robots_full" > gpurun_out/pytest_qnet_r02.log 2>&1; echo "qnet rc=$?" >> gpurun_out/pytest_qnet_r02.log
grep -n "fraction outside\|passed\|failed\|^FAILED" gpurun_out/pytest_qnet_r02.log | head -40
