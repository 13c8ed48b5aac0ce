#!/bin/bash
# f16 + fp8-corrected final convolution: per-op parity, full-network errors per math mode, bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_act.py -m gpu -q -x -k "f16_fp8 or act_tail or test_conv3d" > gpurun_out/pytest_j.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_j.log | cut -c1-300
timeout 900 python tools/report_errors.py > gpurun_out/errors_j.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/errors_j.log | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --math f16f8c > gpurun_out/bench_j_f8c.json 2> gpurun_out/bench_j_f8c.err; echo "rc=$?"; tail -3 gpurun_out/bench_j_f8c.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_j_f8c.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --math bf16x3 > gpurun_out/bench_j_x3.json 2> gpurun_out/bench_j_x3.err; echo "rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_j_x3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['stages_ms'])"
