#!/bin/bash
# compute-sanitizer memcheck + racecheck over GPU tests that exercise the kernels added in round 2: lag statistics +
# persistent Adam kernel (singlecam), fused time-segmented final pass, one-pass bracketed median, geometric
# initialisation, lag-statistics optimiser of the linear models.  Output: gpurun_out/sanitize_r2_<tool>.log
set -u
mkdir -p gpurun_out
SEL='test_forgetting_regimes_match_oracle_fp64 or test_blocks_long_sequences or test_spans_and_offsets or (fused_smoother_equals_exact_scan) or fused_smoother_slow_forgetting or (bracketed_median_is_exact and 131072) or median_with_spans or median_all_nan or (geometric_init_on_device and 1000) or (lag_statistics_optimiser and 2500)'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_r2_$tool.log 2>&1
  echo "$tool exit=$?" >> gpurun_out/sanitize_r2_$tool.log
  tail -4 gpurun_out/sanitize_r2_$tool.log
done
