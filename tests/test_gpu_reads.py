"""Read scorer (teloscope_b200/csrc/scan_reads.cu: ReadTelomereFilter::matches fused, one warp per read) against the CPU
oracle and against the general scan + block kernels (TLS_NO_READ_KERNEL=1) on seeded reads that hit its corners: telomeric
prefixes / suffixes around the -l threshold on both strands, arrays in the middle of a read, arrays that straddle the
1000-base rows of the kernel, two short dense runs that only pass merged (Phase 2, -d), variant repeats, bytes outside
ACGT, a trailing '\\r', reads shorter than the pattern, reads of 100 kb, other -l / -k / -d / -y / pattern sets."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import product_adapter as pa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    return teloscope_b200


def make_reads(rng, n_reads, unit: bytes):
    rc = ol.revcomp(unit.decode()).encode()
    reads = []
    for i in range(n_reads):
        kind = i % 16
        n = int(max(1, rng.normal(15000, 3000))) if kind else int(rng.integers(1, 60))
        if kind == 15:
            n = int(rng.integers(90_000, 130_000))
        body = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)

        def stamp(a, ln, u, vary=0):
            a = max(0, min(a, n - 1)); ln = min(ln, n - a)
            arr = np.frombuffer((u * (ln // len(u) + 2))[:ln], dtype=np.uint8).copy()
            for k in range(3, ln, vary or ln + 1):
                arr[k] = b"ACGT"[int(rng.integers(0, 4))]
            body[a:a + ln] = arr

        if kind in (1, 2):                         # prefix / suffix around the threshold, either strand
            ln = int(rng.integers(30, 60))
            stamp(0 if kind == 1 else n - ln, ln, unit if rng.integers(0, 2) else rc)
        elif kind == 3:
            stamp(int(rng.integers(0, n)), int(rng.integers(36, 5000)), rc if rng.integers(0, 2) else unit, vary=int(rng.integers(7, 60)))
        elif kind == 4:                            # across a row edge of the kernel (rows own 1000 starts)
            stamp(1000 * int(rng.integers(1, max(2, n // 1000))) - int(rng.integers(1, 40)), int(rng.integers(20, 90)), unit)
        elif kind == 5:                            # two short dense runs, close enough to merge or not
            a = int(rng.integers(0, max(1, n - 2000)))
            stamp(a, 12, unit); stamp(a + int(rng.integers(60, 700)), 13, unit)
        elif kind == 6:
            stamp(int(rng.integers(0, n)), int(rng.integers(100, 3000)), unit, vary=int(rng.integers(5, 12)))
        elif kind == 7:
            stamp(0, n, unit)                      # telomeric end to end
        elif kind == 8:
            stamp(int(rng.integers(0, n)), 48, unit); body[rng.integers(0, n, size=20)] = ord("N")
        seq = body.tobytes()
        if kind == 9:
            seq = seq.lower()
        if kind in (10, 1):
            seq += b"\r"
        reads.append(seq)
    return reads


CONFIGS = [
    dict(),                                                      # --fastq-subset defaults: -l 42
    dict(min_block_len=30),
    dict(min_block_len=100, max_block_dist=50),
    dict(max_match_dist=31, min_block_density=0.8),
    dict(canonical="TTTAGGG", min_block_len=42),
    dict(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], min_block_len=42, min_block_counts=1),
    dict(min_block_len=6, min_block_counts=1, min_block_density=0.1),
]


@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_reads_against_oracle_and_general_path(tb, ci):
    kw = dict(min_block_len=42)
    kw.update(CONFIGS[ci])
    cfg = ol.Config(**kw)
    rng = np.random.default_rng(5000 + ci)
    unit = cfg.canonical_fwd.encode() if cfg.canonical_fwd.count("G") >= cfg.canonical_fwd.count("C") else cfg.canonical_rev.encode()
    reads = make_reads(rng, 640, unit)
    orc = ol.Oracle(cfg)
    want = [orc.read_matches(r) for r in reads]
    got = {}
    for general in (False, True):
        if general:
            os.environ["TLS_NO_READ_KERNEL"] = "1"
        try:
            sc = tb.Scanner(pa.scan_config(cfg))
            try:
                res = sc.scan_batch(reads, None, tb.MODE_READS)
                got[general] = [bool(x) for x in res.passed]
                if not general:
                    assert res.n_kernel_launches == 1          # the fused kernel, nothing else
                # a second batch on the same context (buffers reused, counters re-zeroed), fewer reads
                res2 = sc.scan_batch(reads[5:200], None, tb.MODE_READS)
                assert [bool(x) for x in res2.passed] == want[5:200]
            finally:
                sc.close()
        finally:
            os.environ.pop("TLS_NO_READ_KERNEL", None)
    bad = [i for i, (a, b) in enumerate(zip(got[False], want)) if a != b]
    assert not bad, (bad[:10], [len(reads[i]) for i in bad[:10]])
    assert got[True] == want
    assert 40 < sum(want) < 640
