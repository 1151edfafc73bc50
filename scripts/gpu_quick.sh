#!/bin/bash
# Quick GPU check: window-row stress tests, golden parity, kernel timings.  Usage: scripts/gpu_quick.sh <tag> [ncu]
tag=${1:-quick}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_rows.py -x -q > $out/rows.log 2>&1; echo "exit $?" >> $out/rows.log
tail -15 $out/rows.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $out/parity.log 2>&1; echo "exit $?" >> $out/parity.log
tail -4 $out/parity.log
TLS_TRACE=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace.log 2>&1; tail -4 $out/trace.log
if [ "$2" = ncu ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowScanKernel -s 1 -c 1 \
    -o $out/prof_rows python scripts/trace_bench.py 0.25 > $out/ncu_full.log 2>&1
  tail -2 $out/ncu_full.log
fi
