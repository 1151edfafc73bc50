"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol
the header declares, the pattern table producer equals the oracle's (row P of SURVEY §8a), and
the library refuses to run without a B200 (no CPU fallback)."""
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
