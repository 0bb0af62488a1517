#!/bin/bash
# One compute-sanitizer pass (ONE tool per gpurun call: B200_PROFILING.md) over a small subset of the GPU parity tests
# (SURVEY.md section 5).  Usage (on the GPU box): tools/sanitize_subset.sh <memcheck|racecheck|synccheck> <out-dir>
TOOL=${1:-memcheck}
OUT=${2:-gpurun_out}
mkdir -p "$OUT"
SUBSET='tests/test_gpu_parity.py::test_tiny_inputs tests/test_gpu_parity.py::test_fork_heavy_queries tests/test_gpu_parity.py::test_projection_matches_oracle_fold_order tests/test_gpu_parity.py::test_merge_topk_equals_single_forest tests/test_gpu_parity.py::test_reference_spec_conduit_shape tests/test_gpu_parity.py::test_repeated_builds_replay_the_graph_and_follow_new_data'
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 77 --print-limit 20 \
    python -m pytest -x -q -m gpu $SUBSET > "$OUT/sanitizer_$TOOL.log" 2>&1
echo "$TOOL exit=$?" >> "$OUT/sanitizer_summary.txt"
tail -3 "$OUT/sanitizer_$TOOL.log" >> "$OUT/sanitizer_summary.txt"
