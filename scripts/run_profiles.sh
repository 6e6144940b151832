#!/bin/bash
# Run on the GPU box (via gpurun): produces the ncu artefacts summarised under profiles/.
# usage: bash scripts/run_profiles.sh <round-tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --sessions 2 --no-e2e --no-cpu"
# (1) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/launches_$TAG.log 2>&1
# (2) full capture of the dominant kernel (one evaluation in the middle of the Adam loop)
ncu --set full --clock-control none --import-source on -k regex:diag_nll -s 60 -c 1 -o $OUT/prof_${TAG}_nll $CMD > $OUT/prof_${TAG}_nll.log 2>&1
# (3) the other kernels of the step
ncu --set full --clock-control none --import-source on -k regex:"ensemble_staged|select_hist|moments_finalize" -c 5 -o $OUT/prof_${TAG}_rest $CMD > $OUT/prof_${TAG}_rest.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"diag_filter|diag_rts" -c 2 -o $OUT/prof_${TAG}_smooth $CMD > $OUT/prof_${TAG}_smooth.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"diag_adam" -s 60 -c 1 -o $OUT/prof_${TAG}_adam $CMD > $OUT/prof_${TAG}_adam.log 2>&1
# (4) the real numbers (not under a profiler)
python bench.py --steps 3 --warmup 3 > $OUT/bench_${TAG}_c5.json 2> $OUT/bench_${TAG}_c5.err
python bench.py --workload c2 --steps 5 --warmup 3 > $OUT/bench_${TAG}_c2.json 2> $OUT/bench_${TAG}_c2.err
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
tail -c 600 $OUT/bench_${TAG}_c5.json; echo; tail -c 400 $OUT/bench_${TAG}_c2.json
