#!/usr/bin/env python
"""tools/sass_hist.py : per-kernel SASS opcode histogram of sctda_b200/csrc/libsctda_b200.so (cuobjdump -sass),
the mnemonics that tell which hardware paths a kernel uses: DMMA (FP64 tensor), UTCHMMA / LDTM / UTCBAR (tcgen05 +
TMEM), UBLKCP / UTMALDG (TMA), LDGSTS (cp.async), STAS (st.async to a peer CTA), SYNCS (mbarrier), LDG.256, REDG, ...
    python tools/sass_hist.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'sctda_b200', 'csrc', 'libsctda_b200.so')
KEYS = ['DMMA', 'UTCHMMA', 'UTCQMMA', 'LDTM', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'STAS', 'SYNCS', 'LDG', 'LDG.256',
        'STG', 'LDS', 'STS', 'SHFL', 'DFMA', 'DADD', 'DMUL', 'MUFU', 'ATOMG', 'REDG', 'BAR', 'LDL', 'STL']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE).stdout.decode(errors='replace')
    fn, hist, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE).stdout.decode().strip()
            fn = re.sub(r'\(anonymous namespace\)::', '', fn).split('(')[0][:70]
            hist[fn] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and fn:
            op = m.group(1)
            total[fn] += 1
            base = op.split('.')[0]
            hist[fn][base] += 1
            if base == 'LDG' and '.256' in op:
                hist[fn]['LDG.256'] += 1
    print('# SASS opcode counts per kernel (static), libsctda_b200.so built for sm_100a; `cuobjdump -sass | tools/sass_hist.py`')
    print('%-72s %6s  %s' % ('kernel', 'instr', 'selected opcodes'))
    for fn, h in hist.items():
        sel = ' '.join('%s=%d' % (k, h[k]) for k in KEYS if h.get(k))
        print('%-72s %6d  %s' % (fn, total[fn], sel))


if __name__ == '__main__':
    main()
