"""A batch large enough (1.06 Gbp, 66 K units) for the paths small parity cases do not reach: the window-row kernel in
several launches with the window counters of each finished unit range copied while the next range is scanned
(TLS_ROW_PARTS, csrc/api.cu enqueueScan) or in one launch with stream wait-value progress ranges, the compact and packed result records on the same batch, and the ring staging of pageable
host input on four threads.  All variants must return identical results; a sample of contigs is checked against the oracle."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _scan(tb, cfg, seq, offsets, abs_pos, device_ptr=None, env=None):
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        sc = tb.Scanner(cfg)               # TLS_ROW_PARTS is read by tls_create
        try:
            t = sc.submit(None if device_ptr else seq, offsets, abs_pos, tb.MODE_GENOME, device_ptr=device_ptr)
            return sc.collect(t)
        finally:
            sc.close()
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


def test_split_scan_compact_and_pageable_input_agree():
    import torch
    import bench
    import teloscope_b200 as tb
    if tb.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    chroms = bench.plan_assembly(7, 0.35)
    seq, offsets, abs_pos = bench.build_assembly(chroms, "cuda:0", 7)
    n_bases = int(offsets[-1])
    assert n_bases > 1_000_000_000
    host = seq[:n_bases].cpu().numpy()                                  # pageable
    mk = lambda compact, packed=False: tb.ScanConfig(canonical="TTAGGG", patterns=bench.PATTERNS, out_acgt=True, out_matches=True,
                                                     compact=compact, packed=packed)
    one = _scan(tb, mk(False), None, offsets, abs_pos, device_ptr=seq.data_ptr(), env={"TLS_ROW_PARTS": "1"})
    split = _scan(tb, mk(False), None, offsets, abs_pos, device_ptr=seq.data_ptr(), env={"TLS_ROW_PARTS": "4"})
    comp = _scan(tb, mk(True), None, offsets, abs_pos, device_ptr=seq.data_ptr())
    pack = _scan(tb, mk(False, True), None, offsets, abs_pos, device_ptr=seq.data_ptr())          # TLS_OUT_PACKED: 8-byte windows, 4-byte matches
    launches = _scan(tb, mk(False, True), None, offsets, abs_pos, device_ptr=seq.data_ptr(), env={"TLS_ROW_PARTS": "3"})
    paged = _scan(tb, mk(False), host, offsets, abs_pos)
    assert len(one.windows) == (one.window_off[-1]) > 1_000_000
    for other in (split, paged):
        for name in ("windows", "canonical", "noncanonical", "terminal", "its", "window_off", "canonical_off", "n_fwd", "n_rev"):
            assert getattr(one, name).tobytes() == getattr(other, name).tobytes(), name
    for f in ("acgt", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
        assert np.array_equal(one.windows[f], comp.windows[f]), f
    for name in ("canonical", "noncanonical"):
        for f in ("pos", "len", "fwd", "can"):
            assert np.array_equal(getattr(one, name)[f], getattr(comp, name)[f]), (name, f)
    assert one.terminal.tobytes() == comp.terminal.tobytes() and one.its.tobytes() == comp.its.tobytes()
    for other in (pack, launches):
        assert one.windows.tobytes() == other.windows.tobytes()
        for name in ("canonical", "noncanonical", "terminal", "its"):
            assert getattr(one, name).tobytes() == getattr(other, name).tobytes(), name
    # a few contigs against the oracle: the first, the last, one at every quarter of the batch (part boundaries)
    orc = ol.Oracle(ol.Config(patterns=bench.PATTERNS))
    n = len(offsets) - 1
    lens = np.diff(offsets.astype(np.int64))
    picks = sorted({0, n - 1} | {int(np.argmin(np.abs(np.cumsum(lens) - q * n_bases))) for q in (0.25, 0.5, 0.75)})
    for i in picks:
        if lens[i] > 40_000_000:
            continue                                                     # keep the CPU side short
        a, b = int(offsets[i]), int(offsets[i + 1])
        want = orc.scan_segment(host[a:b].tobytes().upper(), int(abs_pos[i]), False)
        got = one.item(i)
        for f in ("acgt", "can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
            assert np.array_equal(got["windows"][f], want["windows"][f]), (i, f)
        assert got["n_canonical"] == len(want["canonical"])
        assert one.n_fwd[i] == len(want["fwd"]) and one.n_rev[i] == len(want["rev"]), i
        for k in ("canonical", "noncanonical"):
            for f in ("pos", "len", "fwd", "can"):
                assert np.array_equal(got[k][f], want[k][f]), (i, k, f)
        for k in ("terminal", "its"):                                    # blocks: every field the reference prints or classifies by
            assert len(got[k]) == len(want[k]), (i, k)
            for f in ("start", "len", "counts", "fwd_count", "can_count", "total_cov", "fwd_cov", "can_cov", "label", "valid_or"):
                assert np.array_equal(got[k][f], want[k][f]), (i, k, f)
