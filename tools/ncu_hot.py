#!/usr/bin/env python
"""tools/ncu_hot.py <ncu source-page csv> [top] : instruction classes by executed count and the stall
reasons of the hottest instructions (ncu -i x.ncu-rep --page source --csv > x.csv)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) == len(h)]
tot = sum(int(r[ix['Instructions Executed']]) for r in body)
samp = sum(int(r[ix['# Samples']]) for r in body)
print('total warp instructions %d, samples %d' % (tot, samp))
cls = collections.Counter()
scls = collections.Counter()
for r in body:
    op = r[ix['Source']].split()
    op = [o for o in op if not o.startswith('@')][0].split('.')[0]
    cls[op] += int(r[ix['Instructions Executed']])
    scls[op] += int(r[ix['# Samples']])
print('%-12s %8s %8s' % ('opcode', 'inst %', 'samples %'))
for k, v in cls.most_common(28):
    print('%-12s %7.2f%% %7.2f%%' % (k, 100.0 * v / tot, 100.0 * scls[k] / samp))
stall_cols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
agg = collections.Counter()
for r in body:
    for k in stall_cols:
        agg[k] += int(r[ix[k]])
print('stall reasons (all samples):', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / samp) for k, v in agg.most_common(8)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print('hottest instructions by samples:')
for r in sorted(body, key=lambda r: -int(r[ix['# Samples']]))[:top]:
    st = sorted(((int(r[ix[k]]), k[6:]) for k in stall_cols), reverse=True)[:2]
    print('%6.2f%%  x%-10s %-60s %s' % (100.0 * int(r[ix['# Samples']]) / samp, r[ix['Instructions Executed']],
                                       r[ix['Source']].strip()[:60], ' '.join('%s:%d' % (k, v) for v, k in st)))
