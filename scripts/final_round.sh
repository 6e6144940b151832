#!/bin/bash
# Round-end evidence run on the GPU box (via gpurun; outputs <= 64 MiB): GPU tests, smoke, the bench lines of the four
# workloads and the CPU reference arm.  Profiles: scripts/run_profiles.sh <tag> full; sanitizers: scripts/sanitize_r2.sh.
# usage: bash scripts/final_round.sh <round-tag>
set -u
TAG=${1:-r2_final}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 500 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err
for W in c2 c3 c4; do
  timeout 400 python bench.py --workload $W --steps 5 --warmup 3 > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err
done
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
python - <<PY
import json
for w in ['c5', 'c2', 'c3', 'c4']:
    d = json.load(open('$OUT/${TAG}_bench_%s.json' % w))
    print(w, '%.3e kf/s' % d['value'], '%.2f ms' % d['ms_per_step'], 'one-touch %.3f' % d['pipeline_one_touch']['frac_of_hbm_peak'],
          'e2e %.3e' % d['e2e']['value'], 'public %.3e' % d['e2e_public']['value'], 'cpu %.3e' % d['cpu_baseline']['value'])
PY
