import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench, teloscope_b200 as tb
seq, offsets, n_telo = bench.build_reads("cuda:0", 242, int(2e9))
sc = tb.Scanner(tb.ScanConfig(min_block_len=42))
for i in range(4):
    t0 = time.perf_counter()
    tk = sc.submit(None, offsets, None, tb.MODE_READS, device_ptr=seq.data_ptr())
    t1 = time.perf_counter()
    res = sc.collect(tk, copy=False)
    t2 = time.perf_counter()
    print(f"iter {i}: submit {1e3*(t1-t0):.2f} ms collect {1e3*(t2-t1):.2f} ms scan {res.scan_kernel_ms:.2f} device {res.device_ms:.2f} passed {int(np.count_nonzero(res.passed))}", flush=True)
    res.release()
