"""Adapters that make teloscope_b200 (the CUDA product, called through its C ABI) look like
Oracle.scan_segment so the same golden comparisons run on both."""
from __future__ import annotations

import numpy as np

import oracle_lib as ol
import teloscope_b200 as tb


def scan_config(cfg: ol.Config, flags: dict | None = None) -> tb.ScanConfig:
    return tb.ScanConfig(canonical=cfg.canonical, patterns=cfg.patterns, edit_distance=cfg.edit_distance,
                         window_size=cfg.window_size, step=cfg.step, terminal_limit=cfg.terminal_limit,
                         max_match_dist=cfg.max_match_dist, min_block_len=cfg.min_block_len,
                         max_block_dist=cfg.max_block_dist, min_block_counts=cfg.min_block_counts,
                         min_block_density=cfg.min_block_density, out_acgt=cfg.count_acgt,
                         out_matches=True)


def to_oracle_windows(win: np.ndarray, n: int, abs_pos: int, cfg: ol.Config) -> np.ndarray:
    out = np.zeros(len(win), dtype=ol.WINDOW_DT)
    for i, w in enumerate(win):
        start = i * cfg.step
        size = min(cfg.window_size, n - start)
        out[i]["start"] = abs_pos + start
        out[i]["size"] = size
        out[i]["acgt"] = w["acgt"]
        for k in ("can_cov", "noncan_cov", "fwd_cov", "rev_cov"):
            out[i][k] = w[k]
        out[i]["gc"], out[i]["entropy"] = ol.window_floats(w["acgt"], size)
    return out


def to_oracle_blocks(b: np.ndarray) -> np.ndarray:
    out = np.zeros(len(b), dtype=ol.BLOCK_DT)
    for name in ol.BLOCK_DT.names:
        out[name] = b[name]
    return out


def to_oracle_matches(m: np.ndarray) -> np.ndarray:
    out = np.zeros(len(m), dtype=ol.MATCH_DT)
    for name in ol.MATCH_DT.names:
        out[name] = m[name]
    return out


class BatchedProduct:
    """Scans all contigs of a parsed FASTA in ONE batch through the C ABI (raw, un-uppercased
    bytes — the kernel folds case) and then serves per-contig results."""

    def __init__(self, scanner: tb.Scanner, cfg: ol.Config, paths, tips_only: bool):
        self.cfg = cfg
        seqs, abs_pos, self.index = [], [], {}
        for pi, (_, comps, _) in enumerate(paths):
            for kind, a, payload in comps:
                if kind == "seg":
                    self.index[(pi, a)] = len(seqs)
                    seqs.append(payload)
                    abs_pos.append(a)
        self.seqs = seqs
        self.res = scanner.scan_batch(seqs, np.array(abs_pos, dtype=np.uint64),
                                      tb.MODE_TIPS if tips_only else tb.MODE_GENOME)
        self._next = 0

    def item_as_oracle(self, i: int, abs_pos: int) -> dict:
        it = self.res.item(i)
        return {"terminal": to_oracle_blocks(it["terminal"]), "its": to_oracle_blocks(it["its"]),
                "canonical": to_oracle_matches(it["canonical"]), "noncanonical": to_oracle_matches(it["noncanonical"]),
                "windows": to_oracle_windows(it["windows"], len(self.seqs[i]), abs_pos, self.cfg),
                "n_canonical": it["n_canonical"]}

    def scan_fn(self, seq_upper: bytes, abs_pos: int, tips_only: bool) -> dict:
        # walk_paths visits contigs in the same order they were batched
        i = self._next
        self._next += 1
        assert len(self.seqs[i]) == len(seq_upper)
        return self.item_as_oracle(i, abs_pos)
