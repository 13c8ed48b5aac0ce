#!/bin/bash
# multi-GPU training step: data-parallel, NCCL gradient all-reduce through the library's own communicator
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload train --batch 16 --steps 3 --warmup 1 > gpurun_out/bench_train_${N}gpu.json 2> gpurun_out/bench_train_${N}gpu.err; echo "rc=$?"
tail -c 1500 gpurun_out/bench_train_${N}gpu.json; tail -5 gpurun_out/bench_train_${N}gpu.err | cut -c1-300
