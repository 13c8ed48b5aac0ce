#!/bin/bash
# per-launch times (ncu, gpu__time_duration only) of the reworked streaming kernels of the training backward; the kernel filter keeps
# ncu off the other ~1 000 launches of a step (an unfiltered list of the 30 GB training step takes > 5 minutes)
mkdir -p gpurun_out
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:trans_bwd|pad_transpose_split|amax_kernel|softmax_rows|softmax_bwd_rows|channel_argmax|split_transpose_scaled|wgrad_tall64|dropout_rows|unscale_kernel' \
  --csv --log-file gpurun_out/launches_r02_train_reworked_kernels.csv python bench.py --workload train --batch 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_train_k.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r02_train_reworked_kernels.csv')) if len(r) > 10]
h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split('(')[0][:60]; a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]): print('%-62s %5d %10.1f us' % (n, c, t / 1000 if t > 1e6 else t))
P
