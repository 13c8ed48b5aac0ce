#!/bin/bash
# streaming kernels of the training backward (register-resident softmax + fused dropout, vectorised amax / argmax / transposing
# splits, trans_decoder adjoint): full GPU test suite, training bench line, launch list of one training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02_train_stream.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02_train_stream.log | cut -c1-300
timeout 600 python bench.py --workload train --batch 16 --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r02_train_1gpu_stream.json 2> gpurun_out/bench_train_stream.err; echo "train rc=$?"; tail -3 gpurun_out/bench_train_stream.err | cut -c1-300; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_train_1gpu_stream.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('stages_ms'))"
