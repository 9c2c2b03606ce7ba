"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/launches_rNN.txt
  python profiles/summarize.py report   gpurun_out/prof.ncu-rep  > profiles/ncu_rNN.txt
  python profiles/summarize.py lsu      gpurun_out/prof.ncu-rep UNITS   LSU data-pipe wavefronts per instruction class
                                                                        (source page), per UNITS of work
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


def lsu(path, units):
    """Where the LSU data-pipe wavefronts of the kernel come from: the source page's per-instruction counters
    (global: one data wavefront per quad of lanes for a 256-bit load = L2 theoretical sectors / ... is not it; the
    raw page's pipe totals are printed next to the per-instruction shared-memory wavefronts and request counts)."""
    units = float(units)
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE).stdout.decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, r = rows[0], rows[2]
    get = lambda k: float(r[hdr.index(k)].replace(',', '')) if k in hdr else float('nan')
    sm = get('launch__grid_size')
    print('per unit of work (%d units in this launch):' % units)
    for k, label in (('smsp__inst_executed.sum', 'warp instructions'),
                     ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'LSU data-pipe wavefronts, shared (incl. shuffles)'),
                     ('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', '   of which shared loads'),
                     ('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', '   of which shared stores'),
                     ('l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum', 'local-memory load wavefronts (spills)'),
                     ('l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum', 'local-memory store wavefronts (spills)'),
                     ('lts__t_requests_srcunit_tex_op_read.sum', 'L2 read requests'),
                     ('dram__bytes_read.sum', 'DRAM bytes read'), ('dram__bytes_write.sum', 'DRAM bytes written')):
        v = get(k)
        if 'dram' in k:
            v *= {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}.get(rows[1][hdr.index(k)], 1.0)
        print('   %-62s %12.1f' % (label, v / units))
    k = 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg'
    if k in hdr:
        print('   %-62s %12.1f' % ('LSU data-pipe wavefronts, all (avg per SM x SMs)', get(k) * sm / units))
        print('   %-62s %12.1f' % ('   of which global / local (lgds)', get('SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg') * sm / units))
    print('   %-62s %12.2f %%' % ('l1tex__data_pipe_lsu_wavefronts, pct of peak', get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')))
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], stdout=subprocess.PIPE).stdout.decode(errors='replace')
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    agg = collections.Counter()
    cnt = collections.Counter()
    for r in rows[2:]:
        if len(r) != len(h):
            continue
        op = [o for o in r[ix['Source']].split() if not o.startswith('@')][0]
        base = op.split('.')[0]
        key = base + ('.' + op.split('.')[-1] if base in ('LDS', 'STS') and op.split('.')[-1] in ('64', '128', 'U16') else '')
        if base == 'LDG' and '.256' in op:
            key = 'LDG.256'
        sh = int(r[ix['L1 Wavefronts Shared']] or 0)
        if sh or base in ('LDG', 'STG', 'SHFL', 'LDL', 'STL'):
            agg[key] += sh
            cnt[key] += int(r[ix['Instructions Executed']])
    print('per-instruction classes: executed warp instructions and shared-memory wavefronts per unit')
    for k, v in sorted(cnt.items(), key=lambda x: -x[1]):
        print('   %-10s %12.1f instr %12.1f shared wavefronts' % (k, v / units, agg[k] / units))


if __name__ == '__main__':
    {'launches': launches, 'report': report, 'lsu': lsu}[sys.argv[1]](*sys.argv[2:])
