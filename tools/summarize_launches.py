#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/launches_*.csv):
launch count, total time and share of the LAST complete forward in the file (between its input_preprocess launch and
the next one; the capture's launch limit usually cuts the final forward short).

    python tools/summarize_launches.py profiles/launches_r01_final.csv

ncu serialises the kernels and runs them cold-cache, so only the SHARES are comparable with the CUDA-event stage
times bench.py reports."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[1:]]
    starts = [i for i, (n, _) in enumerate(seq) if 'input_preprocess' in n]
    if len(starts) >= 2:
        a, b = starts[-2], starts[-1]
        # the forward ends where the next voxelizer call begins; its own voxelizer launches sit right before its start
        nxt = [i for i in range(a, b) if 'vox_scatter' in seq[i][0]]
        body = seq[a:(nxt[0] if nxt else b)]
        head = [x for x in seq[max(0, a - 4):a] if 'vox_' in x[0]][-2:]
        seq = head + body
    elif starts:
        seq = seq[starts[-1]:]
    agg = collections.OrderedDict()
    for n, v in seq:
        k = n.split('(')[0][:64]
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
    tot = sum(v for _, v in agg.values())
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-66s %4d %10.1f us %5.1f%%' % (k, c, v / 1000, 100 * v / tot))
    print('%-66s %4d %10.1f us' % ('total', sum(c for c, _ in agg.values()), tot / 1000))


if __name__ == '__main__':
    main(sys.argv[1])
