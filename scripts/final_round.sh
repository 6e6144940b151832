#!/bin/bash
# Round-end evidence run on the GPU box: tests, smoke, bench lines, multi-camera benches, launch lists.
# usage: bash scripts/final_round.sh <round-tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_${TAG}_c5.json 2> $OUT/bench_${TAG}_c5.err
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 > $OUT/bench_${TAG}_c2.json 2> $OUT/bench_${TAG}_c2.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
timeout 300 python scripts/multicam_pipeline_bench.py 1000000 3 > $OUT/bench_${TAG}_multicam_linear.json 2> $OUT/bench_${TAG}_multicam_linear.err
timeout 300 python scripts/multicam_bench.py both > $OUT/bench_${TAG}_multicam_device.json 2> $OUT/bench_${TAG}_multicam_device.err
# launch lists (time only; cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --sessions 2 --no-e2e --no-cpu > $OUT/launches_${TAG}.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_${TAG}_multicam.csv \
    python scripts/multicam_pipeline_bench.py 200000 1 > $OUT/launches_${TAG}_multicam.log 2>&1
tail -c 700 $OUT/bench_${TAG}_c5.json; echo; tail -c 300 $OUT/bench_${TAG}_multicam_linear.json; echo
