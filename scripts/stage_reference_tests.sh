#!/bin/bash
# Stage the reference's own unit-test files next to the oracle's build outputs (oracle/_ref/ is git-ignored and is
# NOT part of the repo's history; it travels to the GPU box with the snapshot) so that
# tests/test_reference_unit_tests.py can run them there against the eks -> eks_b200 alias.
set -eu
SRC=${1:-/root/reference/tests}
DST="$(dirname "$0")/../oracle/_ref/reference_tests"
mkdir -p "$DST"
for f in test_core.py test_singlecam_smoother.py test_ibl_pupil_smoother.py test_multicam_smoother.py \
         test_marker_array.py test_utils.py test_stats.py test_ibl_paw_multicam_smoother.py; do
  cp "$SRC/$f" "$DST/$f"
done
echo "staged $(ls "$DST" | wc -l) files into $DST"
