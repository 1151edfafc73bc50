#!/usr/bin/env python3
"""Stage timing of `teloscope-b200 --fastq-subset` on synthetic HiFi reads: python scripts/fastq_bench.py [Gbp]"""
import os, subprocess, sys, tempfile, time, shutil
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, bench
gbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
seq, offsets, _ = bench.build_reads("cuda", 542, int(gbp * 1e9))
host = seq.cpu().numpy(); del seq
tmp = Path(tempfile.mkdtemp(prefix="telo_fq_", dir="/dev/shm"))
try:
    fq = tmp / "reads.fq"
    qual = b"I" * int(np.max(np.diff(offsets)))
    with open(fq, "wb") as fh:
        for i in range(len(offsets) - 1):
            a, b = int(offsets[i]), int(offsets[i + 1])
            fh.write(b"@read%d\n" % i); fh.write(host[a:b].tobytes()); fh.write(b"\n+\n"); fh.write(qual[: b - a]); fh.write(b"\n")
    for i in range(3):
        d = tmp / "o"; shutil.rmtree(d, ignore_errors=True); d.mkdir()
        t0 = time.perf_counter()
        p = subprocess.run([str(ROOT / "teloscope_b200" / "teloscope-b200"), "--fastq-subset", "-f", str(fq), "-o", str(d)],
                           capture_output=True, env=dict(os.environ, TLS_PIPE_TRACE="1"))
        print(f"run {i}: {time.perf_counter()-t0:.2f} s rc={p.returncode}"); print(p.stderr.decode())
finally:
    shutil.rmtree(tmp, ignore_errors=True)
