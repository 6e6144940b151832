#!/bin/bash
# compute-sanitizer memcheck + racecheck over GPU tests that exercise the kernels added after the first sanitizer
# pass: pupil run-parallel optimiser, linear tabulated-gain path, verification/reduction kernel, multi-camera
# pre-stage (centring, PCA moments, latent init), Mahalanobis inflation.  Output: gpurun_out/sanitize_new_<tool>.log
set -u
mkdir -p gpurun_out
SEL='test_reference_style_random_input or (run_parallel_generic_linear and 777) or (prestage_matches_oracle_fp64 and 777) or variance_inflation_fixed_loading or test_sessions_batch_equals_single'
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_new_$tool.log 2>&1
  echo "$tool exit=$?" >> gpurun_out/sanitize_new_$tool.log
  tail -4 gpurun_out/sanitize_new_$tool.log
done
