"""Parity of the CUDA path (through the C ABI) against the CPU oracle and against the files the
unmodified reference binary wrote.  Bit-exact: everything here is integer / byte / index work;
the float window tracks are derived from the integer counters with the reference's expressions
and compared as the text the reference prints."""
import numpy as np
import pytest

import golden_cmp
import oracle_lib as ol
import product_adapter as pa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    return teloscope_b200


def _same_blocks(a, b):
    assert len(a) == len(b)
    for name in ("start", "len", "counts", "fwd_count", "rev_count", "can_count", "noncan_count",
                 "total_cov", "fwd_cov", "can_cov", "valid_or", "label"):
        assert np.array_equal(a[name], b[name]), name


def device_supports(cfg, genome: bool = True) -> bool:
    """step == window always (and windows do not exist outside genome mode); overlapping windows for one
    pattern length L with step >= L and window - step >= L (what tls_submit accepts, see buildDevConfig
    in csrc/api.cu)."""
    if cfg.step == cfg.window_size or not genome:
        return True
    lens = {len(p) for p, _ in ol.expand_patterns(cfg.raw_patterns(), cfg.edit_distance, cfg.canonical_fwd)}
    L = max(lens)
    return len(lens) == 1 and cfg.step >= L and cfg.window_size - cfg.step >= L


def test_golden_fasta_cases(tb, golden):
    """Every FASTA case the device supports (all with step == window, and the overlapping-window cases):
    product == reference files, and product == oracle record by record (block order included)."""
    cases, fixtures = golden
    done = overlapping = 0
    for c in cases:
        if c["kind"] != "fasta":
            continue
        cfg, flags = ol.config_from_args(c["args"])
        if not device_supports(cfg, flags["genome"]):
            continue
        overlapping += flags["genome"] and cfg.step != cfg.window_size
        data = fixtures[c["fixture"]].tobytes()
        paths = ol.parse_fasta(data)
        tips = not flags["genome"]
        scanner = tb.Scanner(pa.scan_config(cfg, flags))
        try:
            prod = pa.BatchedProduct(scanner, cfg, paths, tips)
            golden_cmp.check_fasta_case(c, data, prod.scan_fn, cfg, flags)
            # record-level comparison with the oracle
            orc = ol.Oracle(cfg)
            i = 0
            for _, comps, _ in paths:
                for kind, a, payload in comps:
                    if kind != "seg":
                        continue
                    want = orc.scan_segment(payload.upper(), a, tips)
                    got = prod.item_as_oracle(i, a)
                    _same_blocks(got["terminal"], want["terminal"])
                    _same_blocks(got["its"], want["its"])
                    if not tips:
                        assert got["n_canonical"] == len(want["canonical"])
                        for k in ("canonical", "noncanonical"):
                            assert np.array_equal(got[k][["pos", "len", "fwd", "can"]], want[k][["pos", "len", "fwd", "can"]]), k
                        w1, w2 = got["windows"], want["windows"]
                        assert len(w1) == len(w2)
                        for k in ("start", "size", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
                            assert np.array_equal(w1[k], w2[k]), k
                        if cfg.count_acgt:
                            assert np.array_equal(w1["acgt"], w2["acgt"])
                    assert prod.res.n_fwd[i] == len(want["fwd"]) and prod.res.n_rev[i] == len(want["rev"])
                    i += 1
        finally:
            scanner.close()
        done += 1
    assert done >= 170
    assert overlapping >= 3          # -w 500 -s 250 (the reference's own .tst flag set, with the genome-wide outputs), -w 100 -s 30, chr33 -w 2000 -s 500


def test_golden_fastq_cases(tb, golden):
    cases, fixtures = golden
    done = 0
    for c in cases:
        if c["kind"] != "fastq":
            continue
        cfg, flags = ol.config_from_args(c["args"], read_mode=True)
        reads = ol.parse_fastq(fixtures[c["fixture"]].tobytes())
        scanner = tb.Scanner(pa.scan_config(cfg))
        try:
            res = scanner.scan_batch([s for _, s in reads], None, tb.MODE_READS)
            kept = [name for (name, _), ok in zip(reads, res.passed) if ok]
            assert kept == c["kept"], (c["fixture"], c["args"])
        finally:
            scanner.close()
        done += 1
    assert done >= 15


def test_batches_of_only_empty_items(tb):
    """A batch whose items are all empty launches no kernel: as the first submission of a fresh context and
    after a non-empty one (stale per-item counters), every mode returns empty results like the reference's SegmentData."""
    for mode in (tb.MODE_GENOME, tb.MODE_TIPS, tb.MODE_READS):
        sc = tb.Scanner(tb.ScanConfig())
        try:
            for rep in range(2):
                res = sc.scan_batch([b"", b"", b""], None, mode)
                assert res.n_matches == 0
                for i in range(3):
                    it = res.item(i)
                    if mode == tb.MODE_READS:
                        assert not res.passed[i]
                        continue
                    assert len(it["terminal"]) == 0 and len(it["its"]) == 0
                    assert res.n_fwd[i] == 0 and res.n_rev[i] == 0
                    if mode == tb.MODE_GENOME:
                        assert it["n_canonical"] == 0 and len(it["windows"]) == 0
                # a non-empty batch in between leaves per-item counters behind
                full = sc.scan_batch([b"TTAGGG" * 100, b"CCCTAA" * 80 + b"ACGT" * 500, b"ACGTTGCA" * 40], None, mode)
                assert full.n_matches > 0
        finally:
            sc.close()


def test_sharded_job_equals_single_batch(tb):
    """One job over two contexts (two GPUs when the box has them, the same GPU twice otherwise): items partitioned by
    teloscope_b200.shard.partition_items, each share scanned through its own ctx, results put back in submission order —
    identical to the single-batch scan, item by item (what ShardedJob does over gloo, tests/test_shard_gloo.py)."""
    from teloscope_b200.shard import partition_items
    from test_gpu_rows import make_contig
    rng = np.random.default_rng(2024)
    cfg = ol.Config(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], terminal_limit=20000)
    seqs = [make_contig(rng, int(n), "TTAGGG", i % 3 == 0) for i, n in enumerate(rng.integers(2_000, 600_000, size=40))]
    abs_pos = np.cumsum([0] + [len(s) + 100 for s in seqs[:-1]]).astype(np.uint64)
    ndev = tb.load_library().tls_device_count()
    one = tb.Scanner(pa.scan_config(cfg))
    try:
        ref = one.scan_batch(seqs, abs_pos, tb.MODE_GENOME)
        want = [ref.item(i) for i in range(len(seqs))]
    finally:
        one.close()
    parts = partition_items([len(s) for s in seqs], 2)
    got = [None] * len(seqs)
    scanners = [tb.Scanner(pa.scan_config(cfg), device=r % ndev) for r in range(2)]
    try:
        for r, sc in enumerate(scanners):
            res = sc.scan_batch([seqs[i] for i in parts[r]], abs_pos[parts[r]], tb.MODE_GENOME)
            for k, i in enumerate(parts[r]):
                got[int(i)] = res.item(k)
    finally:
        for sc in scanners:
            sc.close()
    for a, b in zip(got, want):
        for k in ("windows", "canonical", "noncanonical", "terminal", "its"):
            assert a[k].tobytes() == b[k].tobytes(), k
        assert a["n_canonical"] == b["n_canonical"]


def test_one_context_through_every_mode(tb):
    """One context sees genome, read, tips and genome batches in turn (the per-batch tables are views into one block in the
    genome / tips path and buffers of their own in the fused read path; the gather runs behind the scan from the second
    batch on): every result equals the one a fresh context gives."""
    rng = np.random.default_rng(20251)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def contig(n, cap):
        body = rng.choice(acgt, size=n)
        body[:cap] = np.frombuffer((b"CCCTAA" * (cap // 6 + 1))[:cap], dtype=np.uint8)
        body[n - cap:] = np.frombuffer((b"TTAGGG" * (cap // 6 + 1))[:cap], dtype=np.uint8)
        return body.tobytes()

    genome = [contig(260_000, 3000), contig(9_000, 600), contig(131_072, 2000), b"", contig(700, 60)]
    reads = [contig(15_000, 400), bytes(rng.choice(acgt, size=12_000)), contig(9_000, 90), b"ACGT" * 50]
    jobs = [(genome, tb.MODE_GENOME), (reads, tb.MODE_READS), (genome, tb.MODE_TIPS), (genome[::-1], tb.MODE_GENOME), (reads[::-1], tb.MODE_READS),
            (genome, tb.MODE_GENOME)]
    names = ("windows", "canonical", "noncanonical", "terminal", "its", "window_off", "canonical_off", "noncanonical_off", "n_fwd", "n_rev", "passed")
    for kw in ({}, {"packed": True}):
        cfg = tb.ScanConfig(patterns=["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"], min_block_len=42, **kw)
        want = []
        for seqs, mode in jobs:
            sc = tb.Scanner(cfg)
            try:
                want.append(sc.scan_batch(seqs, None, mode))
            finally:
                sc.close()
        sc = tb.Scanner(cfg)
        try:
            for (seqs, mode), w in zip(jobs, want):
                got = sc.scan_batch(seqs, None, mode)
                for name in names:
                    assert getattr(got, name).tobytes() == getattr(w, name).tobytes(), (kw, mode, name)
        finally:
            sc.close()
        assert bool(want[1].passed[0]) and not bool(want[1].passed[1]) and len(want[0].terminal) >= 4
