"""Compare scan results (oracle or product) with the files the reference binary wrote."""
from __future__ import annotations

import numpy as np

import oracle_lib as ol


def expected_blocks(lines, ncols=10):
    return ["\t".join(l.split("\t")[:ncols]) for l in lines]


def check_fasta_case(case, fixture_bytes: bytes, scan_fn, cfg=None, flags=None):
    """scan_fn(seq_upper, abs_pos, tips_only) -> dict like Oracle.scan_segment.  Asserts that
    everything the reference wrote for this case is reproduced."""
    if cfg is None:
        cfg, flags = ol.config_from_args(case["args"])
    tips = not flags["genome"]
    paths = ol.parse_fasta(fixture_bytes)
    walked = ol.walk_paths(scan_fn, paths, cfg, tips)
    files = case["files"]

    # gaps.bed — host splitter (src/teloscope.cpp:761-765)
    got = [f"{p['header']}\t{a}\t{a + n}" for p in walked for a, n in p["gaps"]]
    assert got == files["gaps.bed"]

    # terminal blocks: with -n every block is written; otherwise only scaffold-terminal ones
    got = []
    for p in walked:
        T = cfg.terminal_limit
        for line, b in zip(ol.block_lines(p["header"], p["terminal"], p["size"]),
                           sorted(p["terminal"].tolist(), key=lambda t: t[0])):
            start, end = b[0], b[0] + b[1]
            scaffold = start < T or end > ((p["size"] - T) & 0xFFFFFFFFFFFFFFFF)
            if scaffold or flags["n"]:
                got.append(line)
    assert got == expected_blocks(files["terminal_telomeres.bed"])

    if flags["i"]:
        got = [l for p in walked for l in ol.block_lines(p["header"], p["its"], p["size"])]
        assert got == expected_blocks(files["interstitial_telomeres.bed"])

    if flags["m"]:
        for key, fname in (("canonical", "canonical_matches.bed"), ("noncanonical", "noncanonical_matches.bed")):
            got = [f"{p['header']}\t{int(m['pos'])}\t{int(m['pos']) + int(m['len'])}" for p in walked for m in p[key]]
            exp = ["\t".join(l.split("\t")[:3]) for l in files[fname]]
            assert got == exp, fname

    if flags["genome"]:
        # report: its / canonical / windows columns (src/teloscope.cpp:832-840)
        rows = [l for l in files["report.tsv"][1:1 + len(walked)]]
        for p, row in zip(walked, rows):
            cols = row.split("\t")
            assert cols[1] == p["header"]
            assert [int(c) for c in cols[7:10]] == [len(p["its"]), len(p["canonical"]), len(p["windows"])], row

    def track(fname, values):
        exp = files[fname][1:]
        got = [f"{p['header']}\t{int(w['start'])}\t{int(w['start']) + int(w['size'])}\t{ol.fmt_float(v)}"
               for p, vals in values for w, v in zip(p["windows"], vals)]
        assert got == exp, fname

    if flags["r"]:
        tr = [(p, [ol.window_tracks(w) for w in p["windows"]]) for p in walked]
        track("window_repeat_density.bedgraph", [(p, [t["density"] for t in ts]) for p, ts in tr])
        track("window_canonical_ratio.bedgraph", [(p, [t["canonical_ratio"] for t in ts]) for p, ts in tr])
        track("window_strand_ratio.bedgraph", [(p, [t["strand_ratio"] for t in ts]) for p, ts in tr])
    if flags["g"]:
        track("window_gc.bedgraph", [(p, p["windows"]["gc"]) for p in walked])
    if flags["e"]:
        track("window_entropy.bedgraph", [(p, p["windows"]["entropy"]) for p in walked])
    return walked
