"""Sparse work units of the window-row kernel (teloscope_b200/csrc/scan_rows.cu, block_kernels.cu): away from the terminal
zones only canonical matches become records, the terminal walk goes back to the sequence when a run is still open at the
edge of the dense part, and interstitial blocks are rebuilt from the sequence around chains of canonical matches that the
zero-lane masks cannot prove apart.  Seeded contigs that hit those paths, bit-exact against the CPU oracle:
telomere caps longer than the dense prefix / suffix, contigs that are telomeric end to end, interstitial arrays across
window, unit and dense / sparse edges, arrays of thousands of repeats, match distances from 0 to 400 (1 to 13 zero lanes
needed for a proof), canonical homopolymers (slot overflow inside sparse rows), IUPAC bytes inside arrays."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import product_adapter as pa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    os.environ["TLS_ROW_MIN_WINDOWS"] = "1"
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    yield teloscope_b200
    os.environ.pop("TLS_ROW_MIN_WINDOWS", None)


def _arr(unit: bytes, n: int) -> np.ndarray:
    return np.frombuffer((unit * (n // len(unit) + 1))[:n], dtype=np.uint8)


def build(rng, n, unit, w, T, cap_p, cap_q, arrays, iupac, variants):
    """Random contig with a p cap of cap_p bases, a q cap of cap_q, and interstitial arrays (start, bases, strand)."""
    body = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    rc = ol.revcomp(unit.decode()).encode()
    if cap_p:
        body[:cap_p] = _arr(rc, cap_p)
    if cap_q:
        body[n - cap_q:] = _arr(unit, cap_q)
    for a, ln, fwd in arrays:
        a = max(0, min(a, n - ln))
        isl = _arr(unit if fwd else rc, ln).copy()
        for k in range(5, ln, 23):                                   # a variant repeat every few units
            isl[k] = variants[int(rng.integers(0, len(variants)))]
        body[a:a + ln] = isl
    if iupac:
        pos = rng.integers(0, n, size=iupac)
        body[pos] = rng.choice(np.frombuffer(b"RYKMSWrykmBDHV\r", dtype=np.uint8), size=iupac)
    lower = rng.integers(0, n, size=8)
    for a in lower:
        body[a:a + 2000] |= 0x20
    return body.tobytes()


def check(tb, cfg, seqs, packed=False):
    orc = ol.Oracle(cfg)
    scfg = pa.scan_config(cfg)
    scfg.packed = packed                       # TLS_OUT_PACKED: 8-byte windows (+ exception list), 4-byte matches, widened by the wrapper
    sc = tb.Scanner(scfg)
    try:
        res = sc.scan_batch(seqs, None, tb.MODE_GENOME)
        n_its = n_term = 0
        for i, s in enumerate(seqs):
            want = orc.scan_segment(s.upper(), 0, False)
            got = res.item(i)
            assert got["n_canonical"] == len(want["canonical"]), (i, len(s))
            assert res.n_fwd[i] == len(want["fwd"]) and res.n_rev[i] == len(want["rev"]), (i, len(s))
            for k in ("canonical", "noncanonical"):
                for f in ("pos", "len", "fwd", "can"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, len(s), k, f)
            for f in ("acgt", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
                assert np.array_equal(got["windows"][f], want["windows"][f]), (i, len(s), f)
            for k in ("terminal", "its"):
                assert len(got[k]) == len(want[k]), (i, len(s), k, got[k]["start"][:8], want[k]["start"][:8])
                for f in ("start", "len", "counts", "fwd_count", "can_count", "total_cov", "fwd_cov", "can_cov", "label", "valid_or"):
                    assert np.array_equal(got[k][f], want[k][f]), (i, len(s), k, f)
            n_its += len(want["its"]); n_term += len(want["terminal"])
        return n_term, n_its
    finally:
        sc.close()


# window, terminal limit, max match dist (-k), canonical unit
CASES = [
    (1000, 5000, 50, b"TTAGGG"),
    (1000, 5000, 10, b"TTAGGG"),
    (1000, 20000, 200, b"TTAGGG"),
    (640, 3000, 0, b"TTAGGG"),
    (1002, 8000, 400, b"TTAGGG"),
    (1000, 5000, 50, b"TTTAGGG"),
    (900, 2000, 31, b"TTAGG"),
    (1000, 5000, 33, b"TTAGGG"),
    (1000, 5000, 6, b"TTAGGG"),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_sparse_units_against_oracle(tb, case):
    w, T, k, unit = CASES[case]
    rng = np.random.default_rng(9000 + case)
    pats = None
    if unit == b"TTAGGG":
        pats = ["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"]
    cfg = ol.Config(canonical=unit.decode(), patterns=pats, window_size=w, step=w, terminal_limit=T, max_match_dist=k)
    span = 16 * w
    lo = (T // span + 1) * span                                          # first base of the sparse part
    variants = b"ACGT"
    seqs = []
    # caps shorter than / equal to / longer than the dense prefix; arrays at every kind of edge
    for n, cap_p, cap_q in [(20 * span + 777, 300, 300), (14 * span, lo + 5000, 2 * span + 31), (9 * span + 1, lo, lo - 1),
                            (11 * span - 1, 0, 0), (30 * span + 5, 6 * span, 3), (7 * span, 0, lo + span + 17)]:
        hi = ((n - T) // span) * span if n > T else 0
        arrays = [(lo - 200, 700, True), (lo + span - 100, 400, False), (lo + 2 * span + w - 90, 180, True),
                  (hi - 300, 900, False), (lo + 3 * span + 5000, 12000, True), (lo + 5 * span, 6 * 5, True),
                  (lo + 5 * span + 2000, 6 * 4 + 2, False), (lo + 6 * span + 333, 48, True), (n // 2, 3000, False)]
        seqs.append(build(rng, n, unit, w, T, cap_p, cap_q, arrays, iupac=0 if len(seqs) % 2 else 25, variants=variants))
    # telomeric end to end (the p walk crosses the whole sparse part on the sequence), and a contig without any sparse unit
    seqs.append(_arr(ol.revcomp(unit.decode()).encode(), 5 * span + 123).tobytes())
    seqs.append(build(rng, lo + span - 1, unit, w, T, 200, 200, [(lo - 500, 300, True)], 5, variants))
    n_term, n_its = check(tb, cfg, seqs)
    if k >= 6:
        assert n_term >= 6 and n_its >= 6, (n_term, n_its)


def test_sparse_units_canonical_homopolymer(tb):
    """A canonical pattern that matches at every position of a long stretch inside sparse rows: more canonical records than
    the row's slot holds (overflow chunks from a sparse unit), chains of thousands of canonical matches."""
    rng = np.random.default_rng(31337)
    w, T = 1000, 4000
    cfg = ol.Config(canonical="TTTTTT", window_size=w, step=w, terminal_limit=T)
    span = 16 * w
    seqs = []
    for n in (12 * span + 99, 8 * span):
        body = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
        for a, ln, ch in [(2 * span + 100, 5000, b"T"), (4 * span - 40, 90, b"A"), (5 * span + 7, 2 * w + 500, b"a"), (6 * span + 12345, 37, b"T")]:
            body[a:a + ln] = ch[0]
        seqs.append(body.tobytes())
    n_term, n_its = check(tb, cfg, seqs)
    assert n_its >= 4


def test_packed_windows_with_exceptions(tb):
    """TLS_OUT_PACKED keeps 8 bits per match count: windows that are one homopolymer hold ~1000 matches and must come back
    exactly through the exception list, from the window-row kernel (long contigs) and from the pack kernel behind the other
    scan kernels (contigs shorter than a window); ordinary windows beside them stay exact too."""
    rng = np.random.default_rng(271828)
    w, T = 1000, 4000
    cfg = ol.Config(canonical="TTTTTT", window_size=w, step=w, terminal_limit=T)
    span = 16 * w
    seqs = []
    for n in (6 * span + 321, 3 * span):
        body = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
        for a, ln, ch in [(span + 100, 5000, b"T"), (2 * span - 300, 254 + 5, b"A"), (2 * span + 1000, 255 + 5, b"A"), (2 * span + 2000, 256 + 5, b"t"),
                          (n - 2500, 2500, b"A")]:
            body[a:a + ln] = ch[0]
        seqs.append(body.tobytes())
    seqs.append(b"T" * 700)
    seqs.append(b"A" * 300 + bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=150)))
    sc = tb.Scanner(tb.ScanConfig(canonical="TTTTTT", window_size=w, step=w, terminal_limit=T, out_acgt=True, out_matches=True, packed=True))
    try:
        offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum([len(x) for x in seqs], out=offsets[1:])
        raw = sc.collect(sc.submit(np.frombuffer(b"".join(seqs), dtype=np.uint8), offsets, None, tb.MODE_GENOME), copy=False)   # the records as they cross the C ABI
        counts = (raw.windows["w"][:, 1] >> 16) & 255
        assert len(raw.window_exc) >= 8 and int((counts == 255).sum()) == len(raw.window_exc)
        assert sorted(raw.window_exc["window"].tolist()) == np.nonzero(counts == 255)[0].tolist()
        assert int(raw.window_exc["all"].min()) >= 255
        raw.release()
    finally:
        sc.close()
    check(tb, cfg, seqs, packed=True)


def test_gather_behind_the_scan_keeps_or_redoes(tb):
    """From the second batch of a context on, the gather is enqueued behind the scan with the buffers the previous batch left
    (csrc/api.cu enqueueScan) and kept only if the batch's totals fit them.  One context sees a small batch, a batch with
    several times the matches (the speculative gather must be redone after the buffers grew), the small one again (kept);
    every result equals the one of a fresh context, for the three record forms."""
    rng = np.random.default_rng(1618)
    small = [build(rng, 150_000, b"TTAGGG", 1000, 50000, 3000, 2500, [(70_000, 120, True)], iupac=3, variants=b"ACGT")]
    big = [build(rng, n, b"TTAGGG", 1000, 50000, 14000, 9000, [(int(a), 200, bool(a & 1)) for a in rng.integers(60000, n - 60000, size=40)],
                 iupac=5, variants=b"ACGT") for n in (900_000, 400_000, 132_001)]
    for kw in ({}, {"compact": True}, {"packed": True}):
        mk = lambda: tb.Scanner(tb.ScanConfig(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], **kw))
        fresh = []
        for batch in (small, big):
            sc = mk()
            try:
                fresh.append(sc.scan_batch(batch, None, tb.MODE_GENOME))
            finally:
                sc.close()
        sc = mk()
        try:
            for batch, want in ((small, fresh[0]), (big, fresh[1]), (small, fresh[0]), (big, fresh[1])):
                got = sc.scan_batch(batch, None, tb.MODE_GENOME)
                for name in ("windows", "canonical", "noncanonical", "terminal", "its", "window_off", "canonical_off", "noncanonical_off", "n_fwd", "n_rev"):
                    assert getattr(got, name).tobytes() == getattr(want, name).tobytes(), (kw, name)
        finally:
            sc.close()


def test_sparse_units_large_random(tb):
    """A few Mbp of random sequence with the headline pattern set: the common path (no chain survives the zero-lane proof)."""
    rng = np.random.default_rng(4711)
    cfg = ol.Config(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"])
    seqs = []
    for n in (3_000_000, 1_234_567, 200_000, 132_001):
        seqs.append(build(rng, n, b"TTAGGG", 1000, 50000, int(rng.integers(2000, 15000)), int(rng.integers(2000, 15000)),
                          [(int(a), int(l), bool(f)) for a, l, f in zip(rng.integers(60000, n - 60000, size=30), rng.integers(24, 240, size=30), rng.integers(0, 2, size=30))],
                          iupac=12, variants=b"ACGT"))
    check(tb, cfg, seqs)
