"""BASELINE.json configs at scale, inside `pytest -m gpu`:
  * configs[2] (human-shaped assembly, 124 6-mer patterns, -r -g -e -m -i) at 0.1 scale (0.3 Gbp) as a FASTA file through
    the product binary and through the UNMODIFIED reference binary (oracle/_ref/teloscope, built by oracle/Makefile):
    stdout, stderr and all 11 output files byte for byte;
  * configs[3] (plant-shaped assembly, 7-mers: -c CCCTAAA and 4 variants) at 0.05 scale (125 Mbp) through the C ABI against
    the CPU oracle, every contig: windows, both -m lists with flags, terminal and interstitial blocks."""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
BIN = ROOT / "teloscope_b200" / "teloscope-b200"
REF = ROOT / "oracle" / "_ref" / "teloscope"


def _need_gpu():
    import teloscope_b200 as tb
    if tb.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    return tb


def test_config3_file_to_files_equals_reference_binary():
    _need_gpu()
    import bench
    if not REF.exists():
        pytest.skip("oracle/_ref/teloscope not built (make -C oracle ref needs /root/reference)")
    assert BIN.exists(), "build first: make -C teloscope_b200/csrc"
    chroms = bench.plan_assembly(11, 0.1)
    seq, offsets, _ = bench.build_assembly(chroms, "cuda:0", 11)
    host = seq.cpu().numpy()
    del seq
    tmp = Path(tempfile.mkdtemp(prefix="telo_c3_", dir="/dev/shm" if Path("/dev/shm").exists() else None))
    try:
        fa = tmp / "assembly.fa"
        bases = bench.write_fasta(chroms, host, offsets, fa)
        assert bases > 250_000_000
        outs = {}
        for tag, exe, extra in (("ours", BIN, []), ("ref", REF, ["-j", "16"])):
            d = tmp / tag
            d.mkdir()
            p = subprocess.run([str(exe), "-f", str(fa), *bench.REF_FLAGS, *extra, "-o", str(d)], capture_output=True, timeout=900)
            assert p.returncode == 0, (tag, p.stderr[-2000:])
            outs[tag] = (p.stdout, p.stderr, {f.name: f for f in d.iterdir()})
        assert outs["ours"][0] == outs["ref"][0], "stdout"
        assert outs["ours"][1] == outs["ref"][1], "stderr"
        assert sorted(outs["ours"][2]) == sorted(outs["ref"][2]) and len(outs["ref"][2]) == 11
        for name, f in outs["ref"][2].items():
            r = subprocess.run(["cmp", "-s", str(f), str(outs["ours"][2][name])])
            assert r.returncode == 0, name
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_config4_plant_7mers_against_oracle():
    tb = _need_gpu()
    import bench
    chroms = bench.plan_assembly(142, 0.05, [f"chr{i}" for i in range(1, 13)], [2500.0 / 12] * 12, bench.PLANT_PATTERNS, "TTTAGGG")
    seq, offsets, abs_pos = bench.build_assembly(chroms, "cuda:0", 142)
    n_bases = int(offsets[-1])
    assert n_bases > 100_000_000
    host = seq[:n_bases].cpu().numpy()
    cfg = ol.Config(canonical="CCCTAAA", patterns=bench.PLANT_PATTERNS)
    orc = ol.Oracle(cfg)
    sc = tb.Scanner(tb.ScanConfig(canonical="CCCTAAA", patterns=bench.PLANT_PATTERNS, out_acgt=True, out_matches=True))
    try:
        res = sc.collect(sc.submit(None, offsets, abs_pos, tb.MODE_GENOME, device_ptr=seq.data_ptr()))
        n_its = n_term = 0
        for i in range(len(offsets) - 1):
            a, b = int(offsets[i]), int(offsets[i + 1])
            want = orc.scan_segment(host[a:b].tobytes().upper(), int(abs_pos[i]), False)
            got = res.item(i)
            for f in ("acgt", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
                assert np.array_equal(got["windows"][f], want["windows"][f]), (i, f)
            assert got["n_canonical"] == len(want["canonical"]), i
            assert res.n_fwd[i] == len(want["fwd"]) and res.n_rev[i] == len(want["rev"]), i
            for k in ("canonical", "noncanonical"):
                for f in ("pos", "len", "fwd", "can"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, k, f)
            for k in ("terminal", "its"):
                assert len(got[k]) == len(want[k]), (i, k)
                for f in ("start", "len", "counts", "fwd_count", "can_count", "total_cov", "fwd_cov", "can_cov", "label", "valid_or"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, k, f)
            n_its += len(want["its"]); n_term += len(want["terminal"])
        assert n_term >= 20 and n_its >= 50, (n_term, n_its)
    finally:
        sc.close()
