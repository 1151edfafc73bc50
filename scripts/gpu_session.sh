#!/bin/bash
# One gpurun call: the bench line, the ncu launch list of one step and --set full captures of the window-row scan kernel
# and the read scorer.  Usage: scripts/gpu_session.sh <tag> [tests]
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
if [ "$2" = tests ]; then
  timeout 2400 python -m pytest tests -m gpu -q --durations=12 > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
  timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log
fi
timeout 1500 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?" >> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel' -s 36 -c 18 --csv \
  --log-file $out/launches.csv env TLS_PACKED=1 python scripts/trace_bench.py 1.0 > $out/trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowScanKernel -s 1 -c 1 \
  -o $out/prof_rows env TLS_PACKED=1 TLS_ROW_PARTS=1 python scripts/trace_bench.py 0.25 > $out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:readScanKernel -s 1 -c 1 \
  -o $out/prof_reads python scripts/trace_reads.py > $out/ncu_reads.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gatherKernel|gatherDenseKernel|terminalKernel|itsConnKernel|itsWalkKernel' -s 5 -c 5 \
  -o $out/prof_blocks env TLS_PACKED=1 python scripts/trace_bench.py 1.0 > $out/ncu_blocks.log 2>&1
ls -la $out
tail -3 $out/pytest.log $out/smoke.log 2>/dev/null; cat $out/bench.json; tail -3 $out/bench.err; tail -5 $out/trace.log
TLS_TRACE=1 TLS_PACKED=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace_tls.log 2>&1; grep -E "timeline|submit:|row kernel enqueued" $out/trace_tls.log | tail -3 | cut -c1-300
TLS_PIPE_TRACE=1 timeout 600 python scripts/pipeline_bench.py 1.0 --repeat 3 > $out/pipeline_n1.log 2>&1; tail -8 $out/pipeline_n1.log
