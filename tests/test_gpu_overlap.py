"""Overlapping windows (-s < -w; src/teloscope.cpp:409-434,485,512-529,600-637) on the
device against the CPU oracle: window tables (A/C/G/T and the four covered-bp counters of every
overlapping window, short tail windows included), match lists (every match of the contig exactly
once, no loss at window ends), blocks.  Seeded contigs with telomere caps, islands, soft-masked
stretches, bytes outside ACGTacgt and lengths around multiples of step and window.  Bit-exact,
through the C ABI.  Also: the configurations the device refuses are refused at submit time with the
reason, and only for genome-mode batches."""
import numpy as np
import pytest

import oracle_lib as ol
import product_adapter as pa
from test_gpu_rows import make_contig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    return teloscope_b200


CASES = [
    # window, step, terminal limit, canonical, patterns
    (500, 250, 3000, "TTAGGG", None),
    (1000, 100, 5000, "TTAGGG", ["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"]),
    (100, 30, 800, "TTAGGG", None),
    (2000, 500, 20000, "TTAGGG", None),
    (64, 6, 500, "TTAGGG", None),              # step == L
    (64, 58, 500, "TTAGGG", None),             # window - step == L
    (1000, 999 - 6, 4000, "CCCTAAA", None),    # 7-mers, overlap == L
    (333, 111, 50000, "TTAGG", None),          # 5-mers, contigs shorter than the terminal limit
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_overlapping_windows_against_oracle(tb, case):
    w, s, T, canonical, patterns = CASES[case]
    rng = np.random.default_rng(4000 + case)
    cfg = ol.Config(canonical=canonical, patterns=patterns, window_size=w, step=s, terminal_limit=T)
    lens = [int(x) for x in rng.integers(1, 40000, size=10)] + [w, w - 1, w + 1, s, s + 1, s - 1, 7 * s, 7 * s + w, 3 * w + s - 1, 2, 1]
    unit = cfg.canonical_fwd if cfg.canonical_fwd.count("G") >= cfg.canonical_fwd.count("C") else cfg.canonical_rev
    seqs = [make_contig(rng, n, unit, i % 2 == 1) for i, n in enumerate(lens)]
    seqs.append((unit.encode() * (3 * w))[: 2 * w + 17])          # telomeric end to end
    seqs.append(b"")
    abs_pos = np.arange(len(seqs), dtype=np.uint64) * 777_777
    orc = ol.Oracle(cfg)
    sc = tb.Scanner(pa.scan_config(cfg))
    try:
        res = sc.scan_batch(seqs, abs_pos, tb.MODE_GENOME)
        for i, (q, a) in enumerate(zip(seqs, abs_pos)):
            want = orc.scan_segment(q.upper(), int(a), False)
            got = res.item(i)
            assert len(got["windows"]) == len(want["windows"]) == (len(q) + s - 1) // s, (i, len(q))
            for f in ("acgt", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
                assert np.array_equal(got["windows"][f], want["windows"][f]), (i, len(q), f)
            assert got["n_canonical"] == len(want["canonical"]), (i, len(q))
            assert res.n_fwd[i] == len(want["fwd"]) and res.n_rev[i] == len(want["rev"]), (i, len(q))
            for k in ("canonical", "noncanonical"):
                for f in ("pos", "len", "fwd", "can"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, len(q), k, f)
            for k in ("terminal", "its"):
                assert len(got[k]) == len(want[k]), (i, len(q), k)
                for f in ("start", "len", "counts", "fwd_count", "can_count", "total_cov", "fwd_cov", "can_cov", "label", "valid_or"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, len(q), k, f)
    finally:
        sc.close()


def test_unsupported_overlap_is_refused_for_genome_batches_only(tb):
    """step < L, window - step < L (the reference's uint32 scan start wraps, src/teloscope.cpp:413-415)
    and mixed pattern lengths: genome-mode batches fail loudly, tips / read batches (no windows) run."""
    seq = [b"CCCTAA" * 100 + b"ACGT" * 500 + b"TTAGGG" * 100]
    for kw in (dict(window_size=64, step=5), dict(window_size=1000, step=999),
               dict(window_size=500, step=250, patterns=["TTAGGG", "TTTAGGG"])):
        cfg = ol.Config(**kw)
        sc = tb.Scanner(pa.scan_config(cfg))
        try:
            with pytest.raises(tb.TlsError, match="overlapping windows"):
                sc.scan_batch(seq, None, tb.MODE_GENOME)
            tips = sc.scan_batch(seq, None, tb.MODE_TIPS).item(0)
            want = ol.Oracle(cfg).scan_segment(seq[0], 0, True)
            assert np.array_equal(tips["terminal"]["start"], want["terminal"]["start"])
            assert np.array_equal(tips["terminal"]["len"], want["terminal"]["len"])
        finally:
            sc.close()


def test_compact_results_equal_full_results(tb):
    """TLS_OUT_COMPACT (tls_window16 / tls_match8): same values as the 32-byte / 16-byte records, for step == window
    (window-row kernel) and for overlapping windows; refused when a 16-bit window counter could overflow."""
    rng = np.random.default_rng(99)
    seqs = [make_contig(rng, int(n), "TTAGGG", i % 2 == 1) for i, n in enumerate(rng.integers(1, 60000, size=12))]
    seqs += [(b"TTAGGG" * 3000)[:12345], b""]
    abs_pos = np.arange(len(seqs), dtype=np.uint64) * 10_000_019 + (1 << 33)       # positions beyond 32 bits
    for kw in (dict(), dict(window_size=500, step=250), dict(window_size=1000, terminal_limit=3000,
                                                               patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"])):
        cfg = ol.Config(**kw)
        out = []
        for compact in (False, True):
            sc_cfg = pa.scan_config(cfg)
            sc_cfg.compact = compact
            sc = tb.Scanner(sc_cfg)
            try:
                res = sc.scan_batch(seqs, abs_pos, tb.MODE_GENOME)
                out.append(res)
            finally:
                sc.close()
        full, comp = out
        assert comp.windows.dtype == tb.WINDOW16_DT and full.windows.dtype == tb.WINDOW_DT
        for f in ("acgt", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
            assert np.array_equal(full.windows[f], comp.windows[f]), f
        for k in ("canonical", "noncanonical"):
            a, b = getattr(full, k), getattr(comp, k)
            assert len(a) and np.array_equal(getattr(full, k + "_off"), getattr(comp, k + "_off"))
            for f in ("pos", "len", "fwd", "can"):
                assert np.array_equal(a[f], b[f]), (k, f)
        assert full.terminal.tobytes() == comp.terminal.tobytes() and full.its.tobytes() == comp.its.tobytes()
    big = pa.scan_config(ol.Config(window_size=20000, step=20000))
    big.compact = True
    with pytest.raises(tb.TlsError, match="TLS_OUT_COMPACT"):
        tb.Scanner(big)
