out=gpurun_out/e7; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_harness.py -x > $out/pytest.log 2>&1; tail -5 $out/pytest.log
bash scripts/gpu_exp.sh e7 2>&1 | grep -v "^\[tls\] submit" | tail -30
TLS_TRACE=1 TLS_COMPACT=1 timeout 300 python scripts/trace_bench.py 1.0 2>&1 | grep "submit:" | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "scan", d["kernel_only"]["scan_kernels_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
PY
