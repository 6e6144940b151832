#!/bin/bash
# Run on the GPU box: compute-sanitizer memcheck + racecheck over a small set of GPU tests that exercise
# every kernel (ragged lengths, blocks, multicam).  Output: gpurun_out/sanitize_<tool>.log
set -u
mkdir -p gpurun_out
SEL='ragged_lengths_and_alignment and (33 or 2001 or 8193) or blocks_share_s or nll_grad_entry or ensemble_edge or test_multicam_linear_fp64 or test_multicam_nonlinear'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?" >> gpurun_out/sanitize_$tool.log
  tail -4 gpurun_out/sanitize_$tool.log
done
