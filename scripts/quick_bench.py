"""Ad-hoc: time the fused scan on a device-resident random sequence (not the bench contract)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import teloscope_b200 as tb

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mbp * 1_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
codes = torch.randint(0, 4, (n + 4096,), device="cuda", dtype=torch.uint8, generator=g)
lut = torch.tensor(list(b"ACGT"), device="cuda", dtype=torch.uint8)
seq = lut[codes.long()] if n <= 300_000_000 else torch.cat([lut[codes[i:i+100_000_000].long()] for i in range(0, n + 4096, 100_000_000)])
del codes
nitems = 8
offs = np.linspace(0, n, nitems + 1).astype(np.uint64)
for acgt in (True, False):
    cfg = tb.ScanConfig(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], out_acgt=acgt, out_matches=True)
    sc = tb.Scanner(cfg)
    for it in range(3):
        t0 = time.time()
        r = sc.collect(sc.submit(None, offs, None, tb.MODE_GENOME, device_ptr=seq.data_ptr()))
        dt = time.time() - t0
        print(f"acgt={acgt} iter={it} scan_ms={r.scan_kernel_ms:.3f} device_ms={r.device_ms:.3f} wall_ms={dt*1e3:.1f} "
              f"Gbp/s(scan)={n/r.scan_kernel_ms/1e6:.1f} matches={r.n_matches} launches={r.n_kernel_launches}", flush=True)
    sc.close()
