import sys, time, os
import numpy as np, torch
sys.path.insert(0, ".")
import bench, teloscope_b200 as tb
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
chroms = bench.plan_assembly(42, scale)
seq, offsets, abs_pos = bench.build_assembly(chroms, "cuda:0", 42)
print(f"bases {int(offsets[-1])} contigs {len(offsets) - 1}", flush=True)
sc = tb.Scanner(tb.ScanConfig(canonical="TTAGGG", patterns=bench.PATTERNS, out_acgt=True, out_matches=True, compact=bool(os.environ.get("TLS_COMPACT")), packed=bool(os.environ.get("TLS_PACKED"))))
for i in range(4):
    t0 = time.perf_counter()
    tk = sc.submit(None, offsets, abs_pos, tb.MODE_GENOME, device_ptr=seq.data_ptr())
    t1 = time.perf_counter()
    r = tb._Result(); 
    tb._check(sc._L.tls_collect(sc._h, tk, __import__("ctypes").byref(r)))
    t2 = time.perf_counter()
    out = tb.BatchResult(r)
    t3 = time.perf_counter()
    tb._check(sc._L.tls_release(sc._h, tk))
    print(f"iter {i}: submit {1e3*(t1-t0):.1f} ms  collect {1e3*(t2-t1):.1f} ms  python-copy {1e3*(t3-t2):.1f} ms  scan {out.scan_kernel_ms:.2f} device {out.device_ms:.2f}", flush=True)
