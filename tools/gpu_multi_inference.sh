#!/bin/bash
# multi-GPU inference: batch-sharded, no data-path collective (default single-arm workload and config 4's VLM-crop workload)
mkdir -p gpurun_out
N=${1:-8}
for wl in single crop; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps 20 --warmup 3 > gpurun_out/bench_r02_${wl}_${N}gpu.json 2> gpurun_out/bench_r02_${wl}_${N}gpu.err; echo "$wl rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_${wl}_${N}gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['config']['global_batch'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_${N}gpu.json
