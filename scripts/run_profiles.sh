#!/bin/bash
# Run on the GPU box (via gpurun): produces the ncu artefacts summarised under profiles/.
# usage: bash scripts/run_profiles.sh <round-tag> [full]
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --sessions 2 --no-e2e --no-cpu --no-public"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
if [ "${2:-}" = "full" ]; then
# (2) full captures, one launch of each kernel of the step (second step = after warm-up)
for K in lag_stats diag_lag_opt ensemble_staged select_hist select_scan diag_filter diag_rts; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_${TAG}_$K $CMD > $OUT/prof_${TAG}_$K.log 2>&1
done
fi
