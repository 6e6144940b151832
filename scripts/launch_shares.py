"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
Usage: python scripts/launch_shares.py gpurun_out/launches_r1.csv"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.reader(lines):
    rows.append(r)
hdr = rows[0]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[kn]).replace('void ', '').replace('eks::', '')
    v = float(r[mv].replace(',', ''))
    unit = r[mu]
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1.0)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(f'{"kernel":60s} {"launches":>9s} {"total us":>12s} {"share":>7s} {"avg us":>10s}')
for k, v in tot.most_common():
    print(f'{k[:60]:60s} {cnt[k]:9d} {v:12.1f} {100 * v / total:6.1f}% {v / cnt[k]:10.2f}')
print(f'{"TOTAL":60s} {sum(cnt.values()):9d} {total:12.1f}')
