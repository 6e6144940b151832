"""Summarise an .ncu-rep (raw page + source page) into a short text report: key metrics per kernel, SASS
opcode mix and warp-stall breakdown.  Usage: python scripts/ncu_summary.py file.ncu-rep [kernel-regex]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum.per_second',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum.per_second']


def ncu(args):
    return subprocess.run(['ncu', *args], capture_output=True, text=True).stdout


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = raw[0], raw[1]
    kn = hdr.index('Kernel Name')
    for r in raw[2:]:
        if rx and not re.search(rx, r[kn]):
            continue
        print('==', r[kn][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:78s} {r[i]:>16s} {units[i]}')
    extra = ['--kernel-name', f'regex:{rx}'] if rx else []
    src = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', *extra]))))
    hi = [i for i, r in enumerate(src) if 'Source' in r and 'Instructions Executed' in r]
    if not hi:
        return
    h = src[hi[0]]
    end = hi[1] - 1 if len(hi) > 1 else len(src)
    body = [r for r in src[hi[0] + 1:end] if len(r) == len(h)]
    ci, cs = h.index('Instructions Executed'), h.index('Source')
    tot = sum(num(r[ci]) for r in body)
    op = collections.Counter()
    for r in body:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[cs])
        if m:
            op[m.group(2).split('.')[0]] += num(r[ci])
    print(f'  SASS: {len(body)} instructions, {tot:.0f} warp-level instructions executed (first matching launch)')
    print('  opcode mix: ' + ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in op.most_common(14)))
    st = collections.Counter()
    for r in body:
        for i, c in enumerate(h):
            if c.startswith('stall_'):
                st[c] += num(r[i])
    ts = sum(st.values()) or 1
    print('  warp stalls: ' + ', '.join(f'{k[6:]} {100 * v / ts:.1f}%' for k, v in st.most_common(9)))


if __name__ == '__main__':
    main()
