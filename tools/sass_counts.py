#!/usr/bin/env python
"""Counts of the tensor-core / TMA / tensor-memory SASS instructions per kernel of the shipped library
(`cuobjdump -sass voxactb_b200/libvoxactb.so`): UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, UTMALDG = TMA tensor
load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit.   python tools/sass_counts.py > profiles/sass_r02_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'voxactb_b200', 'libvoxactb.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCOMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|UTCCP|SYNCS|ATOMG|REDG|RED)\b')
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        counts.setdefault(name, collections.Counter())
        continue
    if name:
        for k in pat.findall(line):
            counts[name][k] += 1
tot = collections.Counter()
print('%-90s %s' % ('kernel', 'instruction counts'))
for n, c in counts.items():
    if any(k in c for k in ('UTCHMMA', 'UTCQMMA', 'UTMALDG', 'LDTM', 'STTM')):
        print('%-90s %s' % (n[:90], ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
        tot.update(c)
print('%-90s %s' % ('TOTAL (tensor kernels)', ' '.join('%s=%d' % kv for kv in sorted(tot.items()))))
