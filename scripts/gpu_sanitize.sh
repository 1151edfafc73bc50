#!/bin/bash
# compute-sanitizer over the small parity tests: memcheck (out-of-bounds / misaligned accesses of every kernel the tests reach)
# and racecheck (shared-memory hazards of the window-row kernel, the read scorer and the tile kernels).
# Usage: scripts/gpu_sanitize.sh <tag> [race-only]
tag=${1:-san}; out=gpurun_out/$tag; mkdir -p $out
if [ "$2" != race-only ]; then
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_rows.py tests/test_gpu_sparse.py tests/test_gpu_reads.py "tests/test_gpu_parity.py::test_one_context_through_every_mode" \
    "tests/test_gpu_parity.py::test_batches_of_only_empty_items" -q -x -k "not large_random" > $out/memcheck.log 2>&1
  echo "memcheck exit $?" >> $out/memcheck.log; grep -E "ERROR SUMMARY|passed|failed|memcheck exit|Invalid|out of bounds" $out/memcheck.log | head -20
fi
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_rows.py tests/test_gpu_reads.py "tests/test_gpu_parity.py::test_one_context_through_every_mode" -q -x > $out/racecheck.log 2>&1
echo "racecheck exit $?" >> $out/racecheck.log; grep -E "RACECHECK SUMMARY|passed|failed|racecheck exit" $out/racecheck.log | head -20
grep -E "Warning|Error" $out/racecheck.log | cut -c1-160 | sort | uniq -c | sort -rn | head -12
