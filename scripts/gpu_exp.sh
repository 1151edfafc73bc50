#!/bin/bash
# One experiment call: block / ITS parity tests, step trace with the ITS counters, launch lists of the in-tree build and of the
# variants under teloscope_b200/variants/.  Usage: scripts/gpu_exp.sh <tag> [prof-regex]
tag=${1:-exp}; out=gpurun_out/$tag; mkdir -p $out
for t in test_gpu_sparse test_gpu_large test_gpu_parity test_gpu_rows; do
  timeout 600 python -m pytest tests/$t.py -x -q > $out/$t.log 2>&1; echo "exit $?" >> $out/$t.log
  echo "== $t"; tail -5 $out/$t.log
done
TLS_TRACE=1 TLS_COMPACT=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace_tls.log 2>&1; grep -E "ITS chains|iter 3" $out/trace_tls.log | tail -3
for lib in "" teloscope_b200/variants/lib_*.so; do
  name=$(basename ${lib:-default} .so)
  echo "== $name"
  TLS_LIB=${lib:+$PWD/$lib} TLS_COMPACT=1 timeout 300 python scripts/trace_bench.py 1.0 2>&1 | tail -2
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel' -s 22 -c 24 --csv \
    --log-file $out/launches_$name.csv env TLS_LIB=${lib:+$PWD/$lib} TLS_COMPACT=1 python scripts/trace_bench.py 1.0 > $out/trace_$name.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$out/launches_$name.csv")) if len(r)>5 and r[0].isdigit()]
seen=set()
for r in rows[-24:]:
    k=r[4][:40]
    if k in seen: continue
    seen.add(k); print("   ", k, r[-1])
PY
done
if [ -n "$2" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s 3 -c 3 \
    -o $out/prof_down env TLS_COMPACT=1 python scripts/trace_bench.py 1.0 > $out/ncu_down.log 2>&1
fi
