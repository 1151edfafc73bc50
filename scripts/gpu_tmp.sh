out=gpurun_out/e13; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_reads.py tests/test_gpu_bam.py tests/test_gpu_configs.py tests/test_gpu_parity.py -q -x 2>&1 | tail -4
timeout 300 python scripts/trace_reads.py 2>&1 | tail -3
