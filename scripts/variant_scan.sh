#!/bin/bash
# Kernel tuning sweep on the headline batch: scan-kernel / device time of every library build under teloscope_b200/variants/
# and of the in-tree build (scripts/trace_bench.py, compact records, ONE scan launch).  Usage: scripts/variant_scan.sh <tag>
out=gpurun_out/${1:-sweep}; mkdir -p $out
for lib in "" teloscope_b200/variants/lib_*.so ""; do
  name=${lib:-default}
  echo "== $name" | tee -a $out/sweep.log
  TLS_ROW_PARTS=1 TLS_COMPACT=1 TLS_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/trace_bench.py 1.0 2>&1 | tail -2 | tee -a $out/sweep.log
done
