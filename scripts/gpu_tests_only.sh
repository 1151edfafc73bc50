#!/bin/bash
# The whole GPU suite and smoke(), nothing else (what the driver runs at round end).  Usage: scripts/gpu_tests_only.sh <tag>
tag=${1:-tests}; out=gpurun_out/$tag; mkdir -p $out
rm -f gpurun_out/harness_table.md
timeout 1500 python -m pytest tests -x -q -m gpu --durations=12 > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
tail -22 $out/pytest.log; tail -2 $out/smoke.log
