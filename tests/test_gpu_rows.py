"""Window-row kernel (teloscope_b200/csrc/scan_rows.cu) against the CPU oracle on seeded inputs that
hit its corners: window sizes at both ends of its range, contigs whose length is not a multiple of
the window, every 16-byte alignment of the contig start, soft-masked stretches, bytes outside
ACGTacgt (IUPAC letters, '\\r'), telomere caps dense enough to overflow the per-row queue and slot,
terminal limits inside the contig (all three "terminal" window classes, src/teloscope.cpp:403-406),
pattern lengths 5, 6 and 7.  Bit-exact comparisons through the C ABI."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import product_adapter as pa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    os.environ["TLS_ROW_MIN_WINDOWS"] = "1"          # route every eligible contig through the window-row kernel
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    yield teloscope_b200
    os.environ.pop("TLS_ROW_MIN_WINDOWS", None)


def make_contig(rng, n, unit, iupac):
    body = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    rc = ol.revcomp(unit).encode()
    cap = int(rng.integers(0, 4000))
    if cap and cap < n // 3:
        rep = np.frombuffer((rc * (cap // len(rc) + 1))[:cap], dtype=np.uint8)
        body[:cap] = rep
        rep = np.frombuffer((unit.encode() * (cap // len(unit) + 1))[:cap], dtype=np.uint8)
        body[n - cap:] = rep
    for _ in range(int(rng.integers(0, 6))):                     # interstitial islands with a variant every few repeats
        reps = int(rng.integers(4, 40))
        isl = bytearray((unit if rng.integers(0, 2) else rc.decode()).encode() * reps)
        for k in range(0, len(isl), 17):
            isl[k] = b"ACGT"[int(rng.integers(0, 4))]
        a = int(rng.integers(0, max(1, n - len(isl))))
        body[a:a + len(isl)] = np.frombuffer(bytes(isl[:n - a]), dtype=np.uint8)
    for _ in range(int(rng.integers(0, 4))):                     # soft-masked stretches
        ln = int(rng.integers(1, 3000)); a = int(rng.integers(0, max(1, n - ln)))
        body[a:a + ln] |= 0x20
    if iupac:
        k = int(rng.integers(1, 40))
        pos = rng.integers(0, n, size=k)
        body[pos] = rng.choice(np.frombuffer(b"RYKMSWrykm\rBDHVEQ@`", dtype=np.uint8), size=k)
    return body.tobytes()


CASES = [
    # window, terminal limit, canonical, patterns, min contig len, max contig len, iupac
    (1000, 50000, "TTAGGG", None, 1, 70000, False),
    (1000, 5000, "TTAGGG", ["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], 500, 90000, True),
    (600, 3000, "TTAGGG", ["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], 100, 40000, True),
    (1002, 7000, "CCCTAAA", None, 900, 60000, True),
    (777, 2500, "TTAGG", None, 300, 30000, True),
    (999, 1000, "TTAGGG", None, 2000, 50000, False),
    (1000, 4000, "TTTTTT", None, 500, 30000, True),      # homopolymer caps: a match at every position (dense-row path)
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_rows_against_oracle(tb, case):
    w, T, canonical, patterns, lo, hi, iupac = CASES[case]
    rng = np.random.default_rng(1000 + case)
    cfg = ol.Config(canonical=canonical, patterns=patterns, window_size=w, step=w, terminal_limit=T)
    lens = [int(x) for x in rng.integers(lo, hi, size=14)] + [w, w - 1, w + 1, 32 * w, 32 * w + 1, 33 * w - 1, 5, 1]
    unit = cfg.canonical_fwd if cfg.canonical_fwd.count("G") >= cfg.canonical_fwd.count("C") else cfg.canonical_rev
    seqs = [make_contig(rng, n, unit, iupac and (i % 3 != 0)) for i, n in enumerate(lens)]
    # one contig that is telomeric end to end: more matches per row than the queue and the slot hold
    seqs.append((unit.encode() * (6 * w))[: 5 * w + 123])
    abs_pos = np.arange(len(seqs), dtype=np.uint64) * 1_000_003
    orc = ol.Oracle(cfg)
    sc = tb.Scanner(pa.scan_config(cfg))
    try:
        res = sc.scan_batch(seqs, abs_pos, tb.MODE_GENOME)
        for i, (s, a) in enumerate(zip(seqs, abs_pos)):
            want = orc.scan_segment(s.upper(), int(a), False)
            got = res.item(i)
            assert got["n_canonical"] == len(want["canonical"]), (i, len(s))
            assert res.n_fwd[i] == len(want["fwd"]) and res.n_rev[i] == len(want["rev"]), (i, len(s))
            for k in ("canonical", "noncanonical"):
                for f in ("pos", "len", "fwd", "can"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, len(s), k, f)
            for f in ("acgt", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
                assert np.array_equal(got["windows"][f], want["windows"][f]), (i, len(s), f)
            for k in ("terminal", "its"):
                assert len(got[k]) == len(want[k]), (i, len(s), k)
                for f in ("start", "len", "counts", "fwd_count", "can_count", "total_cov", "fwd_cov", "can_cov", "label", "valid_or"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, len(s), k, f)
    finally:
        sc.close()


def test_rows_equal_general_kernel(tb):
    """Same batch through the window-row kernel and through the general kernel (TLS_FORCE_GENERAL):
    identical windows, match lists and blocks."""
    rng = np.random.default_rng(77)
    cfg = ol.Config(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], terminal_limit=20000)
    seqs = [make_contig(rng, int(n), "TTAGGG", True) for n in rng.integers(50_000, 400_000, size=12)]
    out = []
    for force in (False, True):
        if force:
            os.environ["TLS_FORCE_GENERAL"] = "1"
        try:
            sc = tb.Scanner(pa.scan_config(cfg))
            try:
                res = sc.scan_batch(seqs, None, tb.MODE_GENOME)
                out.append([res.item(i) for i in range(len(seqs))])
            finally:
                sc.close()
        finally:
            os.environ.pop("TLS_FORCE_GENERAL", None)
    for a, b in zip(*out):
        for k in ("windows", "canonical", "noncanonical", "terminal", "its"):
            assert a[k].tobytes() == b[k].tobytes(), k


def test_rows_without_acgt_counts(tb):
    """-r / -m / -i without -g / -e: the COUNT = false instantiation of the window-row kernel (no A/C/G/T counting);
    the window's A/C/G/T fields stay zero, everything else equals the oracle."""
    rng = np.random.default_rng(4242)
    cfg = ol.Config(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], terminal_limit=6000)
    cfg.count_acgt = False
    seqs = [make_contig(rng, int(n), "TTAGGG", i % 2 == 0) for i, n in enumerate(rng.integers(900, 80000, size=10))]
    orc = ol.Oracle(cfg)
    sc = tb.Scanner(pa.scan_config(cfg))
    try:
        res = sc.scan_batch(seqs, None, tb.MODE_GENOME)
        for i, s in enumerate(seqs):
            want = orc.scan_segment(s.upper(), 0, False)
            got = res.item(i)
            assert not got["windows"]["acgt"].any() and not want["windows"]["acgt"].any()
            for f in ("can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
                assert np.array_equal(got["windows"][f], want["windows"][f]), (i, f)
            for k in ("canonical", "noncanonical"):
                assert np.array_equal(got[k]["pos"], want[k]["pos"]), (i, k)
            for k in ("terminal", "its"):
                assert np.array_equal(got[k]["start"], want[k]["start"]) and np.array_equal(got[k]["len"], want[k]["len"]), (i, k)
    finally:
        sc.close()
