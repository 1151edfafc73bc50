#!/bin/bash
# Quick GPU check of the block kernels: parity tests that reach them, a traced step (timeline of the collect phase), a bench line.
# Usage: scripts/gpu_check.sh <tag>
tag=${1:-check}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_large.py tests/test_gpu_parity.py tests/test_gpu_overlap.py tests/test_gpu_rows.py tests/test_gpu_reads.py tests/test_gpu_configs.py -q -x > $out/pytest.log 2>&1; tail -5 $out/pytest.log
TLS_TRACE=1 TLS_PACKED=1 timeout 300 python scripts/trace_bench.py 1.0 > $out/trace_tls.log 2>&1; grep -E "timeline|submit:|gather_blocks" $out/trace_tls.log | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -3 $out/bench.err; python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "scan", d["kernel_only"]["scan_kernels_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"], "launches", d["gpu_launches"])
PY
