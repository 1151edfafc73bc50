#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of one step and one
# --set full capture of the lean scan kernel.  Usage: scripts/gpu_session.sh <tag> [full|quick]
tag=${1:-run}; mode=${2:-full}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
if [ "$mode" = full ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
  timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel' -s 24 -c 28 --csv \
  --log-file $out/launches.csv python scripts/trace_bench.py 1.0 > $out/trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowScanKernel -s 1 -c 1 \
  -o $out/prof_rows python scripts/trace_bench.py 0.25 > $out/ncu_full.log 2>&1
ls -la $out
tail -3 $out/pytest.log $out/smoke.log 2>/dev/null; cat $out/bench.json; tail -3 $out/bench.err; tail -5 $out/trace.log
