#!/bin/bash
# Run on the GPU box (via gpurun): produces the ncu artefacts summarised under profiles/.
# usage: bash scripts/run_profiles.sh <round-tag> [full]
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
C5="python bench.py --steps 1 --warmup 1 --sessions 2 --no-e2e --no-cpu --no-public"
C3="python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-public"
C4="python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu --no-public"
KSEL='regex:_ZN3eks'
# (1) every launch of the library with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --kernel-name-base mangled -k $KSEL --log-file $OUT/launches_${TAG}_c5.csv $C5 > $OUT/launches_${TAG}_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --kernel-name-base mangled -k $KSEL --log-file $OUT/launches_${TAG}_c3.csv $C3 > $OUT/launches_${TAG}_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --kernel-name-base mangled -k $KSEL -c 400 --log-file $OUT/launches_${TAG}_c4.csv $C4 > $OUT/launches_${TAG}_c4.log 2>&1
if [ "${2:-}" = "full" ]; then
# (2) full captures, one launch of each kernel of the step (after the warm-up step)
# (the .ncu-rep files are summarised on the box and removed: gpurun brings back at most 64 MiB)
capture() {   # kernel regex, command...
  local K=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_${TAG}_$K "$@" > $OUT/prof_${TAG}_$K.log 2>&1
  python scripts/ncu_summary.py $OUT/prof_${TAG}_$K.ncu-rep > $OUT/${TAG}_ncu_$K.txt 2>&1
  rm -f $OUT/prof_${TAG}_$K.ncu-rep
}
for K in lag_stats diag_lag_opt ensemble_staged med_count med_final diag_smooth_fused; do capture $K $C5; done
for K in mlag_stats lin_lag_opt; do capture $K $C3; done
for K in gen_nll_runs triangulate_mean; do capture $K $C4; done
fi
ls -la $OUT | tail -30
