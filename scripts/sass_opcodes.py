"""Static SASS opcode listing of the hot kernels of libeks_b200.so (cuobjdump; no GPU needed).
Usage: python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'eks_b200', 'lib')
HOT = [('ensemble.o', r'ensemble_staged_kernelIffLi10ELb1'), ('diag_lag.o', r'lag_stats_kernelIfLi128'),
       ('diag_lag.o', r'diag_lag_opt_kernelIfE'), ('diag_smooth.o', r'diag_smooth_fused_kernelIfE'),
       ('prestage.o', r'med_count_kernelIfE'), ('prestage.o', r'med_final_kernelIfE'), ('prestage.o', r'med_sample_kernelIfE'),
       ('lin_lag.o', r'mlag_stats_kernelIfLi128'), ('lin_lag.o', r'lin_lag_opt_kernelIfLi4'),
       ('generic_runs.o', r'gen_nll_runs_kernelIfLi3ELi6ELb1ELb1'), ('generic_runs.o', r'gen_filter_runs_kernelIfLi3ELi4ELb1ELb0'),
       ('triangulate.o', r'triangulate_mean_kernelIfE'), ('diag.o', r'diag_nll_kernelIfE')]
for obj, rx in HOT:
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(LIB, obj)], capture_output=True, text=True).stdout
    blocks = re.split(r'\n\s*Function : ', out)
    for b in blocks[1:]:
        name = b.split('\n', 1)[0].strip()
        if not re.search(rx, name):
            continue
        ops = collections.Counter()
        for ln in b.split('\n'):
            m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
            if m:
                ops[m.group(2)] += 1
        tot = sum(ops.values())
        demangled = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()[:110]
        print(f'== {demangled}\n   {tot} SASS instructions: ' + ', '.join(f'{k} {v}' for k, v in ops.most_common(18)))
        marks = [k for k in ('UBLKCP', 'SYNCS', 'LDGSTS', 'FFMA2', 'DFMA', 'MUFU', 'HMMA', 'UTMALDG') if ops.get(k)]
        print('   notable: ' + (', '.join(f'{k} x{ops[k]}' for k in marks) or '-'))
        break
