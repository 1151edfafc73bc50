#!/usr/bin/env python3
"""File -> files timing of the host pipeline binary (and, with --ref, of the unmodified reference) on the
synthetic human-shaped assembly of bench.py written as FASTA to /dev/shm.
    python scripts/pipeline_bench.py [scale] [--ref] [--repeat N] [--gpus N]"""
import shutil, subprocess, sys, tempfile, time, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench

scale = float(next((a for a in sys.argv[1:] if not a.startswith("-")), "0.0625"))
repeat = int(sys.argv[sys.argv.index("--repeat") + 1]) if "--repeat" in sys.argv else 2
import torch
chroms = bench.plan_assembly(1234, scale)
seq, offsets, abs_pos = bench.build_assembly(chroms, "cuda" if torch.cuda.is_available() else "cpu", 1234)
tmp = Path(tempfile.mkdtemp(prefix="telo_pipe_", dir="/dev/shm"))
try:
    fa = tmp / "sample.fa"
    t0 = time.perf_counter()
    bases = bench.write_fasta(chroms, seq.cpu().numpy(), offsets, fa)
    print(f"FASTA: {bases/1e6:.0f} Mbp, {fa.stat().st_size/1e6:.0f} MB written in {time.perf_counter()-t0:.1f} s", flush=True)
    del seq
    env = dict(os.environ, TLS_PIPE_TRACE="1")
    gpus = sys.argv[sys.argv.index("--gpus") + 1] if "--gpus" in sys.argv else "1"
    runs = [("teloscope-b200", [str(ROOT / "teloscope_b200" / "teloscope-b200"), "--gpus", gpus])]
    if "--ref" in sys.argv:
        runs.append(("reference", [str(ROOT / "oracle" / "_ref" / "teloscope"), "-j", str(os.cpu_count())]))
    outs = {}
    for name, cmd in runs:
        for i in range(repeat):
            out = tmp / f"out_{name}"
            shutil.rmtree(out, ignore_errors=True); out.mkdir()
            t0 = time.perf_counter()
            p = subprocess.run([*cmd, "-f", str(fa), *bench.REF_FLAGS, "-o", str(out)], capture_output=True, env=env)
            dt = time.perf_counter() - t0
            print(f"{name} run {i}: rc={p.returncode} {dt:.2f} s  {bases/dt/1e9:.3f} Gbp/s", flush=True)
            if name != "reference":
                print("\n".join(l for l in p.stderr.decode().splitlines() if l.startswith("[pipe]") or (l.startswith("[tls]") and "--tls" in sys.argv)), flush=True)
        outs[name] = out
    if len(outs) == 2:
        same = True
        for f in sorted(outs["reference"].iterdir()):
            a, b = f.read_bytes(), (outs["teloscope-b200"] / f.name).read_bytes()
            if a != b:
                same = False; print("DIFFERENT:", f.name, len(a), len(b))
        print("all output files byte-identical to the reference's:", same)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
