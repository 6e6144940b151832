#!/bin/bash
# Run HERE (no GPU) after `gpurun -- bash scripts/run_profiles.sh <tag>`: turns gpurun_out/ artefacts into the
# tracked summaries under profiles/.
set -u
TAG=${1:-r1}
mkdir -p profiles
cp gpurun_out/launches_$TAG.csv profiles/${TAG}_launches.csv
python scripts/launch_shares.py gpurun_out/launches_$TAG.csv > profiles/${TAG}_launch_shares.txt
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_nll.ncu-rep diag_nll > profiles/${TAG}_ncu_diag_nll_kernel.txt
for k in ensemble_staged select_hist moments_finalize; do
  python scripts/ncu_summary.py gpurun_out/prof_${TAG}_rest.ncu-rep $k > profiles/${TAG}_ncu_$k.txt
done
for k in diag_filter diag_rts; do
  python scripts/ncu_summary.py gpurun_out/prof_${TAG}_smooth.ncu-rep $k > profiles/${TAG}_ncu_$k.txt
done
python scripts/ncu_summary.py gpurun_out/prof_${TAG}_adam.ncu-rep diag_adam > profiles/${TAG}_ncu_diag_adam.txt
for w in c5 c2 reference; do cp gpurun_out/bench_${TAG}_$w.json profiles/${TAG}_bench_$w.json; done
python - <<PY
import csv, io, json, subprocess
raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_${TAG}_nll.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, r = rows[0], rows[1], rows[2]
def val(name):
    i = h.index(name); x = float(r[i].replace(',', '')); unit = u[i]
    return x * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(unit, 1)
dram = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
# the profiled command: 2 sessions x 20 keypoints x 2 channels x 1e6 frames x 4 B, one evaluation, all active
alg = 2 * 20 * 2 * 1_000_000 * 4
json.dump({'diag_nll_kernel': {'dram_bytes': dram, 'algorithmic_bytes': alg,
           'source': 'profiles/${TAG}_ncu_diag_nll_kernel.txt (ncu --set full, bench.py --sessions 2, one launch)'}},
          open('profiles/traffic.json', 'w'), indent=1)
print('traffic ratio', dram / alg)
PY
ls -la profiles
