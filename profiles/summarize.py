"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/launches_rNN.txt
  python profiles/summarize.py report   gpurun_out/prof.ncu-rep  > profiles/ncu_rNN.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed']


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(',', ''))
        except ValueError:
            continue
        a = agg.setdefault(r[kn].split('(')[0][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('# per-kernel device time from `ncu --metrics gpu__time_duration.sum` (cold-cache, serialised: compare SHARES)')
    print('%-92s %6s %12s %7s' % ('kernel', 'n', 'total ms', 'share'))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-92s %6d %12.3f %6.1f%%' % (k, n, t / 1e6, 100 * t / tot))


def report(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE).stdout.decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        if name in seen:
            continue
        seen.add(name)
        print('== %s' % name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('   %-72s %12s %s' % (k, r[i], units[i]))


if __name__ == '__main__':
    {'launches': launches, 'report': report}[sys.argv[1]](sys.argv[2])
