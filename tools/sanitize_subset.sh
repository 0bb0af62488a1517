#!/bin/bash
# compute-sanitizer passes over a small subset of the GPU parity tests: memcheck, racecheck, synccheck.
# Usage (on the GPU box): tools/sanitize_subset.sh <out-dir>
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SUBSET='tests/test_gpu_parity.py::test_tiny_inputs tests/test_gpu_parity.py::test_fork_heavy_queries tests/test_gpu_parity.py::test_projection_matches_oracle_fold_order tests/test_gpu_parity.py::test_merge_topk_equals_single_forest tests/test_gpu_parity.py::test_reference_spec_conduit_shape tests/test_gpu_parity.py::test_repeated_builds_replay_the_graph_and_follow_new_data'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
      python -m pytest -x -q -m gpu $SUBSET > "$OUT/sanitizer_$tool.log" 2>&1
  echo "$tool exit=$?" >> "$OUT/sanitizer_summary.txt"
  tail -3 "$OUT/sanitizer_$tool.log" >> "$OUT/sanitizer_summary.txt"
done
