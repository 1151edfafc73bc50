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


def test_packed_records_widen_exactly():
    """tls_window8 / tls_window_exc / tls_match4 (include/teloscope_b200.h) as the binding widens them (BatchResult._build):
    fabricated records, no device.  Field packing, the 255 = 'see the exception list' rule and item-relative positions."""
    rng = np.random.default_rng(8)
    n_win, L = 500, 6
    acgt = rng.integers(0, 1001, size=(n_win, 4)).astype(np.uint32)
    acgt[:4] = [[1023, 0, 0, 0], [0, 1023, 0, 0], [0, 0, 1023, 0], [0, 0, 0, 1023]]
    allm = rng.integers(0, 254, size=n_win).astype(np.uint32)
    can = (allm * rng.random(n_win)).astype(np.uint32)
    fwd = (allm * rng.random(n_win)).astype(np.uint32)
    big = rng.choice(n_win, size=9, replace=False)
    allm[big] = rng.integers(255, 1000, size=9); can[big] = allm[big] // 2; fwd[big] = allm[big] - 3
    can[big[0]] = 255; can[big[1]] = 254                      # the boundary values inside exception windows
    w = np.zeros(n_win, dtype=tb.WINDOW8_DT)
    w["w"][:, 0] = acgt[:, 0] | (acgt[:, 1] << 10) | (acgt[:, 2] << 20) | ((acgt[:, 3] & 3) << 30)
    w["w"][:, 1] = (acgt[:, 3] >> 2) | (np.minimum(can, 255) << 8) | (np.minimum(allm, 255) << 16) | (np.minimum(fwd, 255) << 24)
    exc = np.zeros(len(big), dtype=tb.WINDOW_EXC_DT)
    order = rng.permutation(len(big))                          # the list is unordered
    exc["window"], exc["canonical"], exc["all"], exc["forward"] = big[order], can[big][order], allm[big][order], fwd[big][order]
    win_off = np.array([0, 200, 200, n_win], dtype=np.uint64)  # three items, the second empty
    can_off = np.array([0, 3, 3, 5], dtype=np.uint64)
    m4 = np.array([5, 70 | (1 << 31), 0x7FFFFFFF, 0, 9 | (1 << 31)], dtype=np.uint32)
    non_off = np.zeros(4, dtype=np.uint64)
    abs_pos = np.array([1000, 5, 1 << 40], dtype=np.uint64)
    r = tb._Result()
    r.n_items, r.mode, r.pattern_len = 3, tb.MODE_GENOME, L
    keep = [w, exc, win_off, can_off, m4, non_off]
    r.window_off, r.windows8 = win_off.ctypes.data, w.ctypes.data
    r.window_exc, r.n_window_exc = exc.ctypes.data, len(exc)
    r.canonical_off, r.canonical4, r.noncanonical_off = can_off.ctypes.data, m4.ctypes.data, non_off.ctypes.data
    r.terminal_off = r.its_off = non_off.ctypes.data
    out = tb.BatchResult(r, copy=True, abs_pos=abs_pos)
    assert np.array_equal(out.windows["acgt"], acgt)
    assert np.array_equal(out.windows["can_cov"], can * L) and np.array_equal(out.windows["noncan_cov"], (allm - can) * L)
    assert np.array_equal(out.windows["fwd_cov"], fwd * L) and np.array_equal(out.windows["rev_cov"], (allm - fwd) * L)
    assert out.canonical["pos"].tolist() == [1005, 1070, 1000 + 0x7FFFFFFF, 1 << 40, (1 << 40) + 9]
    assert out.canonical["fwd"].tolist() == [0, 1, 0, 0, 1] and set(out.canonical["len"].tolist()) == {L} and set(out.canonical["can"].tolist()) == {1}
    assert len(out.noncanonical) == 0 and len(out.item(1)["windows"]) == 0 and len(out.item(2)["windows"]) == 300
    lazy = tb.BatchResult(r, copy=False)                       # views: nothing is built before the first use
    assert lazy._raw is not None and lazy.windows["w"].shape == (n_win, 2) and lazy._raw is None
    assert len(lazy.window_exc) == len(big)
    del keep


def _load_test_module(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, ROOT / "tests" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_harness_setup_is_green_for_the_reference_itself(tmp_path, monkeypatch):
    """The GPU suite runs the reference's acceptance harness against teloscope-b200 (tests/test_gpu_harness.py) with the
    validator behind a sharding wrapper.  Here the same tree, wrapper and test functions run against the UNMODIFIED reference
    binary on the CPU: the harness as we drive it must be green for the reference (and the wrapper must not lose a case)."""
    import shutil
    import stat
    ref = ROOT / "oracle" / "_ref" / "teloscope"
    h = _load_test_module("test_gpu_harness")
    if not ref.exists() or not h.HARNESS.exists() or not h.VALIDATE.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    tree = tmp_path / "tree"
    for d in ("validateFiles", "testFiles", "scripts"):
        shutil.copytree(h.HARNESS / d, tree / d)
    (tree / ".github" / "workflows").mkdir(parents=True)
    shutil.copy(h.HARNESS / ".github" / "workflows" / "val.sh", tree / ".github" / "workflows" / "val.sh")
    bin_dir = tree / "build" / "bin"
    bin_dir.mkdir(parents=True)
    for name, text in (("teloscope", f"#!/bin/sh\nexec {ref} \"$@\"\n"), ("teloscope-validate", h.VALIDATE_WRAPPER)):
        (bin_dir / name).write_text(text)
        (bin_dir / name).chmod((bin_dir / name).stat().st_mode | stat.S_IEXEC)
    shutil.copy(h.VALIDATE, bin_dir / "teloscope-validate.real")
    monkeypatch.setattr(h, "ROOT", tmp_path)                       # the pass / fail table and logs go to the scratch directory
    h.test_validate_files(tree)
    h.test_val_sh(tree)
    h.test_shell_scripts(tree, "test_gaps_bed.sh")
    h.test_shell_scripts(tree, "test_n50.sh")
    h.test_bam_subset_script(tree)
    table = (tmp_path / "gpurun_out" / "harness_table.md").read_text()
    assert table.count("| PASS |") >= 400 and table.count("| FAIL |") == 0


def test_golden_pipeline_test_is_green_for_the_reference_itself(golden, tmp_path, monkeypatch):
    """tests/test_gpu_pipeline.py::test_pipeline_equals_reference_outputs (every supported golden FASTA case through the binary,
    six at a time: stdout, stderr, every file) driven against the unmodified reference binary on the CPU."""
    ref = ROOT / "oracle" / "_ref" / "teloscope"
    if not ref.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    tp = _load_test_module("test_gpu_pipeline")

    class OneDevice:
        def tls_device_count(self):
            return 1
    monkeypatch.setattr(tb, "load_library", lambda: OneDevice())
    monkeypatch.setattr(tp, "BIN", ref)
    tp.test_pipeline_equals_reference_outputs(golden, tmp_path)
