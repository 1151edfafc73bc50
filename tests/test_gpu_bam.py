"""--bam-subset of teloscope-b200 (host/bam_subset.cpp: parallel BGZF inflate, SEQ decode, read scorer on the device, parallel
BGZF deflate) against the UNMODIFIED reference binary on the same BAM files: the subset BAM byte for byte (stdout stream and
-o file), stderr, exit code.  Inputs: seeded reads (tests/test_gpu_reads.py shapes: telomeric ends around the -l threshold,
arrays across the kernel's row edges, N bytes, reads of 100 kb, reads without SEQ) as unmapped BAM records; small inflate
segments so that records straddle segment edges; the reference's default -l 42 and explicit block parameters."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
BIN = ROOT / "teloscope_b200" / "teloscope-b200"
REF = ROOT / "oracle" / "_ref" / "teloscope"


def test_bam_subset_equals_reference_binary(tmp_path):
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    if not REF.exists():
        pytest.skip("oracle/_ref/teloscope not built (make -C oracle ref needs /root/reference)")
    import bench
    from test_gpu_reads import make_reads
    rng = np.random.default_rng(77)
    reads = [r.rstrip(b"\r").upper().replace(b"\r", b"N") for r in make_reads(rng, 400, b"TTAGGG")]
    reads[7] = b""                                               # records without SEQ are counted and dropped
    reads[200] = b""
    host = np.frombuffer(b"".join(reads), dtype=np.uint8)
    offsets = np.cumsum([0] + [len(r) for r in reads]).astype(np.uint64)
    bam = tmp_path / "reads.bam"
    bench.write_unmapped_bam(bam, host, offsets, 4)
    for flags in ([], ["-x", "0", "-l", "18", "-y", "0.8", "-k", "10", "-d", "10"], ["-l", "100"]):
        outs = {}
        for tag, exe, env in (("ref", REF, {}), ("ours", BIN, {}), ("ours_small_segments", BIN, {"TLS_BAM_SEGMENT": "300000"})):
            d = tmp_path / f"{tag}_{len(flags)}"
            d.mkdir()
            p = subprocess.run([str(exe), "--bam-subset", "-f", str(bam), *flags], capture_output=True, timeout=600, env=dict(os.environ, **env))
            q = subprocess.run([str(exe), "--bam-subset", "-f", str(bam), *flags, "-o", str(d)], capture_output=True, timeout=600, env=dict(os.environ, **env))
            assert p.returncode == 0 and q.returncode == 0, (tag, p.stderr, q.stderr)
            assert q.stdout == b"" and (d / "reads_telomeric.bam").read_bytes() == p.stdout, tag
            outs[tag] = (p.stdout, p.stderr, q.stderr.replace(str(d).encode(), b"<out>"))
        assert outs["ours"] == outs["ref"], flags
        assert outs["ours_small_segments"] == outs["ref"], flags
        assert b"skipped 2 records without SEQ" in outs["ref"][1] and len(outs["ref"][0]) > 100_000
