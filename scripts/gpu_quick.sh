#!/bin/bash
# Quick GPU check: window-row stress tests (dense + sparse units), golden parity, kernel timings.
# Usage: scripts/gpu_quick.sh <tag> [ncu]
tag=${1:-quick}; out=gpurun_out/$tag; mkdir -p $out
for t in test_gpu_sparse test_gpu_rows test_gpu_parity test_gpu_overlap test_gpu_large; do
  timeout 600 python -m pytest tests/$t.py -x -q > $out/$t.log 2>&1; echo "exit $?" >> $out/$t.log
  echo "== $t"; tail -12 $out/$t.log
done
TLS_TRACE=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace.log 2>&1; tail -6 $out/trace.log
TLS_COMPACT=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace_compact.log 2>&1; tail -3 $out/trace_compact.log
TLS_ROW_PARTS=1 TLS_COMPACT=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace_compact_p1.log 2>&1; tail -3 $out/trace_compact_p1.log
TLS_ROW_PARTS=2 TLS_COMPACT=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace_compact_p2.log 2>&1; tail -3 $out/trace_compact_p2.log
if [ "$2" = ncu ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel' -s 24 -c 28 --csv \
    --log-file $out/launches.csv python scripts/trace_bench.py 1.0 > $out/trace_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowScanKernel -s 1 -c 1 \
    -o $out/prof_rows python scripts/trace_bench.py 0.25 > $out/ncu_full.log 2>&1
  tail -2 $out/ncu_full.log
fi
timeout 300 python scripts/trace_reads.py > $out/trace_reads.log 2>&1; tail -3 $out/trace_reads.log
