"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol
the header declares, the pattern table producer equals the oracle's (row P of SURVEY §8a), and
the library refuses to run without a B200 (no CPU fallback)."""
import numpy as np
import ctypes as C
import itertools
import re
from pathlib import Path

import pytest

import oracle_lib as ol
import teloscope_b200 as tb

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "teloscope_b200.h").read_text()
    declared = set(re.findall(r"\b(tls_[a-z_]+)\s*\(", header))
    assert declared == set(tb.EXPORTS), declared ^ set(tb.EXPORTS)
    lib = C.CDLL(str(tb.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    L = tb.load_library()
    if L.tls_device_count() > 0:
        pytest.skip("a B200 is visible; the refusal path is exercised on the CPU box")
    with pytest.raises(tb.TlsError, match="no CUDA device|sm_100"):
        tb.Scanner(tb.ScanConfig())


PATTERN_SETS = [
    (["CCCTAA", "TTAGGG"], "CCCTAA"), (["TTAGGG"], "CCCTAA"), (["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], "CCCTAA"),
    (["CCCTAAA", "TTTAGGG"], "CCCTAAA"), (["TTAGGG", "TTTAGGG"], "CCCTAA"), (["TTAGG", "TTAGGG"], "CCCTAA"),
    (["NNAGGG"], "CCCTAA"), (["TTAGGR", "YCCTAA"], "CCCTAA"), (["TTAGGGTTAGGG"], "CCCTAA"), (["TTAA"], "CCCTAA"),
    (["ACGT", "TTAGGG"], "CCCTAA"), (["CCCTAAA", "CCCTAAG", "CCCTAAC", "CCCTAAT", "CCCTAGA"], "CCCTAAA"),
    (["TTAGGGZ"], "CCCTAA"), (["AAAAAA"], "AAAAAA"), (["TTAGGGTTAGGGTTAGGGTTAGGGTTAGGG"], "CCCTAA"),
]


@pytest.mark.parametrize("x", [0, 1, 2])
def test_pattern_table_equals_oracle(x):
    for raw, can_fwd in PATTERN_SETS:
        if x == 2 and max(len(r) for r in raw) > 12:
            continue
        want = ol.expand_patterns(raw, x, can_fwd)
        got = tb.expand_patterns(raw, x, can_fwd)
        can_rev = ol.revcomp(can_fwd)
        assert [(p, f) for p, f, _ in got] == want, (raw, x)
        assert [c for _, _, c in got] == [p in (can_fwd, can_rev) for p, _ in want]


def _hamming_ball(seed, x):
    out = {seed}
    for _ in range(x):
        nxt = set(out)
        for s in out:
            for i, ch in enumerate(s):
                for b in "ACGT":
                    if b != ch:
                        nxt.add(s[:i] + b + s[i + 1:])
        out = nxt
    return out


def test_expand_patterns_under_asan(tmp_path):
    """-x 2 grows the variant vector while reading from it (use-after-free in the distance-2 expansion):
    the producer runs under AddressSanitizer and must equal a brute-force enumeration of the Hamming balls."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = tmp_path / "asan_patterns"
    src = ROOT / "teloscope_b200" / "csrc"
    subprocess.run(["g++", "-std=c++17", "-g", "-O0", "-fsanitize=address", f"-I{src}", str(ROOT / "tests" / "asan_patterns_main.cpp"),
                    str(src / "patterns.cpp"), "-o", str(exe)], check=True)
    for x, seeds, can in [(2, ["CCCTAA"], "CCCTAA"), (2, ["TTAGGG", "TCAGGG"], "CCCTAA"), (2, ["CCCTAAA"], "CCCTAAA"), (1, ["TTAGGR"], "CCCTAA")]:
        r = subprocess.run([str(exe), str(x), can] + seeds, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        got = [ln.split()[0] for ln in r.stdout.splitlines()]
        want = set()
        for s in seeds:
            for combo in ol.iupac_expand(s) if hasattr(ol, "iupac_expand") else [s]:
                for v in _hamming_ball(combo, x):
                    want.add(v)
                    want.add(ol.revcomp(v))
        if all(set(s) <= set("ACGT") for s in seeds):
            assert got == sorted(want), (seeds, x, len(got), len(want))
        assert [(p, f) for p, f in (ln.split()[:2] for ln in r.stdout.splitlines())] == \
            [(p, str(int(f))) for p, f in ol.expand_patterns(seeds, x, can)]


def test_golden_case_pattern_counts(golden):
    cases, _ = golden
    seen = set()
    for c in cases:
        cfg, _ = ol.config_from_args(c["args"], read_mode=c["kind"] == "fastq")
        key = (tuple(cfg.raw_patterns()), cfg.edit_distance, cfg.canonical_fwd)
        if key in seen:
            continue
        seen.add(key)
        n = len(tb.expand_patterns(cfg.raw_patterns(), cfg.edit_distance, cfg.canonical_fwd))
        assert f"Scanning {n} telomeric variants" in c["stderr"]


def test_validity_formula():
    """The lean kernel's span validity test (scan_fast.cu ValidAcc): exact on all 256 byte values in
    every byte lane, and over random multi-word spans."""
    valid = set(b"ACGTacgt")
    M = 0xFFFFFFFF

    def bad(words):
        a1, a2, any_ = 0, M, 0
        for w in words:
            m = w & 0x06060606
            y = (m * 6 + 0x40404040) & M
            z = (w * 16) & M
            a1 |= y ^ w
            a2 &= z ^ w
            any_ |= w
        return ((a1 & 0xD0D0D0D0) | (~a2 & 0x10101010) | (any_ & 0x08080808)) & M

    for pos in range(4):
        for b in range(256):
            for other in (0x41, 0x74, 0x47, 0x63):
                by = [other] * 4
                by[pos] = b
                assert (bad([int.from_bytes(bytes(by), "little")]) != 0) == (b not in valid), (pos, b)
    rng = np.random.default_rng(3)
    pool = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    for _ in range(20000):
        by = rng.choice(pool, size=12)
        if rng.random() < 0.5:
            by[rng.integers(0, 12)] = rng.integers(0, 256)
        words = [int.from_bytes(bytes(by[i:i + 4].tolist()), "little") for i in range(0, 12, 4)]
        assert (bad(words) != 0) == any(int(x) not in valid for x in by)


def test_pipeline_binary_has_no_cpu_path(tmp_path):
    """teloscope-b200 is built with the library, parses its flags like the reference (messages, exit codes) and fails
    loudly when there is no B200: nothing is scanned on the host."""
    import subprocess
    exe = ROOT / "teloscope_b200" / "teloscope-b200"
    assert exe.exists(), "build first: make -C teloscope_b200/csrc"
    fa = tmp_path / "a.fa"
    fa.write_text(">s extra words\nACGTNNNNacgt\nTTAGGG\n")
    p = subprocess.run([str(exe), "-f", str(fa), "-w", "10", "-s", "20", "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 1 and b"Error: Step size (20) cannot be larger than window size (10)." in p.stderr
    p = subprocess.run([str(exe), "-f", str(fa), "-x", "3", "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 1 and b"Error: Edit distance (-x/--edit-distance) must be in the range [0,2]." in p.stderr
    p = subprocess.run([str(exe), "--bam-subset", str(fa)], capture_output=True)
    assert p.returncode == 1 and b"Error: BAM subset failed: input is not BGZF-compressed BAM." in p.stderr
    p = subprocess.run([str(exe), "-f", str(fa), "--plot-report", "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 2                    # front ends this binary does not have: said so, status 2
    if tb.load_library().tls_device_count() > 0:
        pytest.skip("a B200 is visible; the refusal path is exercised on the CPU box")
    p = subprocess.run([str(exe), "-f", str(fa), "-c", "ccctaa", "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 1
    err = p.stderr.decode()
    assert "Setting canonical pattern: CCCTAA and its reverse complement: TTAGGG" in err
    assert "Scanning 38 telomeric variants (includes reverse complements)." in err
    assert "tls_create failed" in err and ("no CUDA device" in err or "sm_100" in err)
    assert not list(tmp_path.glob("a.fa_*")), "no output file may appear without a scan"


def test_gfa_reader_writer_equals_reference(tmp_path):
    """GFA front end of teloscope-b200 (host/gfa_pipeline.cpp) without the scan: the graph read and written back as GFA 1.2
    (--gfa-normalize, no device) must equal what the UNMODIFIED reference writes for the same input once the synthetic
    telomere_ lines are taken out — for every GFA fixture of the reference (GFA1 with and without header / paths / jumps /
    tags, '*' sequences, a GFA2 segment file).  Needs oracle/_ref (make -C oracle ref)."""
    import subprocess
    ref = ROOT / "oracle" / "_ref" / "teloscope"
    fixtures = sorted((ROOT / "oracle" / "_ref" / "harness" / "testFiles").glob("*.gfa"))
    exe = ROOT / "teloscope_b200" / "teloscope-b200"
    if not ref.exists() or not fixtures or not exe.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    assert len(fixtures) >= 15
    for f in fixtures:
        out = tmp_path / f.stem
        out.mkdir()
        subprocess.run([str(ref), "-f", str(f), "-j", "1", "-x", "0", "-l", "60", "-o", str(out)], check=True, capture_output=True)
        want = [l for l in (out / (f.name + ".telo.annotated.gfa")).read_text().splitlines() if "telomere_" not in l]
        mine = tmp_path / (f.stem + ".norm.gfa")
        subprocess.run([str(exe), "--gfa-normalize", str(f), str(mine)], check=True, capture_output=True)
        assert mine.read_text().splitlines() == want, f.name
