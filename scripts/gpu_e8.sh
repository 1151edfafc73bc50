out=gpurun_out/e8; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_large.py tests/test_gpu_parity.py tests/test_gpu_overlap.py -q -x > $out/pytest.log 2>&1; tail -15 $out/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -3 $out/bench.err; python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "scan", d["kernel_only"]["scan_kernels_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"], "launches", d["gpu_launches"])
PY
