"""ctypes binding of the TEST-ONLY CPU oracle (oracle/oracle.cpp) plus small host-side helpers
that restate the reference's callers (FASTA splitting, path walk, writers) so that oracle and
product results can be compared with files written by the unmodified reference binary.

Nothing under teloscope_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "_build" / "libtelo_oracle.so"

MATCH_DT = np.dtype([("pos", "<u8"), ("len", "<u2"), ("fwd", "u1"), ("can", "u1"), ("pad", "<u4")])
BLOCK_DT = np.dtype([("start", "<u8"), ("len", "<u4"), ("counts", "<u4"), ("fwd_count", "<u4"),
                     ("rev_count", "<u4"), ("can_count", "<u4"), ("noncan_count", "<u4"),
                     ("total_cov", "<u4"), ("fwd_cov", "<u4"), ("can_cov", "<u4"),
                     ("valid_or", "u1"), ("label", "S1"), ("pad", "<u2")])
WINDOW_DT = np.dtype([("start", "<u8"), ("size", "<u4"), ("acgt", "<u4", (4,)), ("can_cov", "<u4"),
                      ("noncan_cov", "<u4"), ("fwd_cov", "<u4"), ("rev_cov", "<u4"),
                      ("gc", "<f4"), ("entropy", "<f4"), ("pad", "<u4")])
assert MATCH_DT.itemsize == 16 and BLOCK_DT.itemsize == 48 and WINDOW_DT.itemsize == 56


class OrcParams(C.Structure):
    _fields_ = [("window_size", C.c_uint32), ("step", C.c_uint32), ("terminal_limit", C.c_uint32),
                ("max_match_dist", C.c_uint16), ("min_block_len", C.c_uint16),
                ("max_block_dist", C.c_uint16), ("min_block_counts", C.c_uint16),
                ("min_block_density", C.c_float), ("count_acgt", C.c_uint8), ("pad", C.c_uint8 * 3)]


def revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


@dataclass
class Config:
    """Mirror of the scan-relevant part of UserInputTeloscope (include/input.h:15-64) after
    main() has processed the flags (src/main.cpp:287-296, 617-643)."""
    canonical: str = "TTAGGG"
    patterns: list[str] | None = None       # -p seeds (upper-cased); None = canonical pair
    edit_distance: int = 1
    window_size: int = 1000
    step: int = 1000
    terminal_limit: int = 50000
    max_match_dist: int = 50
    min_block_len: int = 300
    max_block_dist: int = 500
    min_block_counts: int = 2
    min_block_density: float = 0.5
    count_acgt: bool = True
    canonical_fwd: str = field(init=False)
    canonical_rev: str = field(init=False)

    def __post_init__(self):
        c = self.canonical.upper()
        rc = revcomp(c)
        self.canonical_fwd, self.canonical_rev = (c, rc) if c <= rc else (rc, c)

    def raw_patterns(self) -> list[str]:
        if self.patterns:
            return [p.upper() for p in self.patterns]
        return [self.canonical_fwd, self.canonical_rev]

    def params(self) -> OrcParams:
        return OrcParams(self.window_size, self.step, self.terminal_limit, self.max_match_dist,
                         self.min_block_len, self.max_block_dist, self.min_block_counts,
                         float(np.float32(self.min_block_density)), 1 if self.count_acgt else 0)


def build_oracle() -> None:
    if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (ROOT / "oracle" / "oracle.cpp").stat().st_mtime:
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "oracle"], check=True,
                       stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.orc_expand_patterns.restype = C.c_int
        L.orc_expand_patterns.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(OrcParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_scan_segment.restype = C.c_void_p
        L.orc_scan_segment.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_result_free.argtypes = [C.c_void_p]
        for fn in (L.orc_result_matches, L.orc_result_blocks):
            fn.restype = C.c_size_t
            fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.orc_result_windows.restype = C.c_size_t
        L.orc_result_windows.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.orc_read_matches.restype = C.c_int
        L.orc_read_matches.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.orc_scan_many.restype = C.c_uint64
        L.orc_scan_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_window_floats.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib = L
    return _lib


def window_floats(acgt, size: int) -> tuple[float, float]:
    a = (C.c_uint32 * 4)(*[int(x) for x in acgt])
    gc, ent = C.c_float(), C.c_float()
    lib().orc_window_floats(a, int(size), C.byref(gc), C.byref(ent))
    return gc.value, ent.value


def expand_patterns(raw: list[str], x: int, canonical_fwd: str) -> list[tuple[str, bool]]:
    buf = C.create_string_buffer(1 << 22)
    n = lib().orc_expand_patterns(",".join(raw).encode(), x, canonical_fwd.encode(), buf, len(buf))
    if n < 0:
        raise ValueError(f"orc_expand_patterns failed: {n}")
    out = []
    for line in buf.value.decode().splitlines():
        p, f = line.split("\t")
        out.append((p, f == "1"))
    assert len(out) == n
    return out


def _copy(ptr, n, dt):
    if n == 0:
        return np.zeros(0, dtype=dt)
    raw = C.string_at(ptr, n * dt.itemsize)
    return np.frombuffer(raw, dtype=dt).copy()


class Oracle:
    """One pattern table + parameter set (a `Teloscope` instance of the reference)."""

    MATCH_KINDS = ("canonical", "noncanonical", "fwd", "rev", "all")

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.table = expand_patterns(cfg.raw_patterns(), cfg.edit_distance, cfg.canonical_fwd)
        pats = b"".join(p.encode() + b"\0" for p, _ in self.table)
        fwd = bytes(1 if f else 0 for _, f in self.table)
        params = cfg.params()
        self._h = lib().orc_create(pats, fwd, len(self.table), cfg.canonical_fwd.encode(),
                                   cfg.canonical_rev.encode(), C.byref(params))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def scan_segment(self, seq: bytes, abs_pos: int = 0, tips_only: bool = False) -> dict:
        L = lib()
        r = L.orc_scan_segment(self._h, seq, len(seq), abs_pos, 1 if tips_only else 0)
        try:
            out = {}
            p = C.c_void_p()
            for k, name in enumerate(self.MATCH_KINDS):
                n = L.orc_result_matches(r, k, C.byref(p))
                out[name] = _copy(p, n, MATCH_DT)
            for k, name in enumerate(("terminal", "its")):
                n = L.orc_result_blocks(r, k, C.byref(p))
                out[name] = _copy(p, n, BLOCK_DT)
            n = L.orc_result_windows(r, C.byref(p))
            out["windows"] = _copy(p, n, WINDOW_DT)
            return out
        finally:
            L.orc_result_free(r)

    def read_matches(self, seq: bytes) -> bool:
        return bool(lib().orc_read_matches(self._h, seq, len(seq)))

    def scan_many(self, buf: np.ndarray, offsets: np.ndarray, tips_only: bool, threads: int) -> int:
        assert buf.dtype == np.uint8 and offsets.dtype == np.uint64
        return int(lib().orc_scan_many(self._h, buf.ctypes.data, offsets.ctypes.data,
                                       len(offsets) - 1, 1 if tips_only else 0, threads))


# -------------------------------------------------------------------------------------------------
# Host-side restatements used only to compare against files written by the reference binary.
# -------------------------------------------------------------------------------------------------

def parse_fasta(data: bytes):
    """gfalibs FASTA loader semantics (gfalibs/src/input-gfa.cpp:41-72, gfalibs/src/gfa.cpp:88-239):
    header = text up to the first space; only '\\n' is removed from the sequence; maximal runs of
    N/n/X/x become gaps, everything else stays in the contig.  Returns a list of paths:
    (header, components, path_size) with components ('seg', abs_start, bytes) / ('gap', abs_start, len).
    """
    assert data[:1] == b">"
    paths = []
    for rec in data[1:].split(b">"):
        nl = rec.find(b"\n")
        head_line = rec if nl < 0 else rec[:nl]
        body = b"" if nl < 0 else rec[nl + 1:]
        sp = head_line.find(b" ")
        header = (head_line if sp < 0 else head_line[:sp]).replace(b"\r", b"")  # walkPath erases '\r'
        seq = body.replace(b"\n", b"")
        comps = []
        is_gap = np.isin(np.frombuffer(seq, dtype=np.uint8), np.frombuffer(b"NnXx", dtype=np.uint8))
        if len(seq):
            edges = np.flatnonzero(np.diff(is_gap.astype(np.int8))) + 1
            starts = np.concatenate(([0], edges))
            ends = np.concatenate((edges, [len(seq)]))
            for a, b in zip(starts.tolist(), ends.tolist()):
                if is_gap[a]:
                    comps.append(("gap", a, b - a))
                else:
                    comps.append(("seg", a, seq[a:b]))
        paths.append((header.decode(), comps, len(seq)))
    return paths


def fmt_float(x) -> str:
    """std::ostream << float with default precision (6 significant digits, %g style)."""
    return "%g" % float(np.float32(x))


def window_tracks(w) -> dict:
    """Per-window float tracks exactly as src/teloscope.cpp:788-803 computes them (float32)."""
    f32 = np.float32
    total = int(w["fwd_cov"]) + int(w["rev_cov"])
    density = f32(total) / f32(int(w["size"]))
    canon = f32(int(w["can_cov"])) / f32(int(w["can_cov"]) + int(w["noncan_cov"])) if total > 0 else f32(-1.0)
    strand = f32(int(w["fwd_cov"])) / f32(int(w["fwd_cov"]) + int(w["rev_cov"])) if total > 0 else f32(-1.0)
    return {"density": density, "canonical_ratio": canon, "strand_ratio": strand}


def walk_paths(scan_fn, paths, cfg: Config, tips_only: bool):
    """Teloscope::walkPath (src/input.cpp:942-1041) minus classification: per path, scan every
    contig with its absPos and concatenate.  scan_fn(seq_upper, abs_pos, tips_only) -> dict as
    Oracle.scan_segment returns."""
    out = []
    for header, comps, size in paths:
        agg = {"header": header, "size": size, "gaps": [], "terminal": [], "its": [],
               "canonical": [], "noncanonical": [], "windows": []}
        for kind, a, payload in comps:
            if kind == "gap":
                agg["gaps"].append((a, payload))
                continue
            seq = payload.upper()
            r = scan_fn(seq, a, tips_only)
            for k in ("terminal", "its", "canonical", "noncanonical", "windows"):
                agg[k].append(r[k])
        for k, dt in (("terminal", BLOCK_DT), ("its", BLOCK_DT), ("canonical", MATCH_DT),
                      ("noncanonical", MATCH_DT), ("windows", WINDOW_DT)):
            agg[k] = np.concatenate(agg[k]) if agg[k] else np.zeros(0, dtype=dt)
        out.append(agg)
    return out


def block_lines(header: str, blocks, path_size: int) -> list[str]:
    """First ten BED columns of src/teloscope.cpp:722-731 / :747-756 for every block, sorted by
    start like labelTerminalBlocks does (:282-285)."""
    rows = []
    for b in sorted(blocks.tolist(), key=lambda t: t[0]):
        start, ln = b[0], b[1]
        rows.append("\t".join(str(v) for v in (header, start, start + ln, ln, b[11].decode(), b[3], b[4],
                                               b[5], b[6], path_size)))
    return rows


def config_from_args(args: list[str], read_mode: bool = False) -> tuple[Config, dict]:
    """Scan-relevant subset of the reference's getopt loop (src/main.cpp:255-506).
    Returns (Config, flags) where flags holds the output toggles and mode."""
    kw, flags = {}, {"r": False, "g": False, "e": False, "m": False, "i": False, "n": False, "l_set": False}
    i = 0
    while i < len(args):
        a = args[i]
        val = args[i + 1] if i + 1 < len(args) else None
        if a == "-c": kw["canonical"] = val.upper(); i += 2
        elif a == "-p": kw["patterns"] = [p.upper() for p in val.split(",") if p]; i += 2
        elif a == "-x": kw["edit_distance"] = int(val); i += 2
        elif a == "-w": kw["window_size"] = int(val); i += 2
        elif a == "-s": kw["step"] = int(val); i += 2
        elif a == "-t": kw["terminal_limit"] = int(val); i += 2
        elif a == "-k": kw["max_match_dist"] = int(val); i += 2
        elif a == "-l": kw["min_block_len"] = int(val); flags["l_set"] = True; i += 2
        elif a == "-d": kw["max_block_dist"] = int(val); i += 2
        elif a == "-y": kw["min_block_density"] = float(np.float32(float(val))); i += 2
        elif a == "-j": i += 2
        elif a in ("-r", "-g", "-e", "-m", "-i", "-n"): flags[a[1]] = True; i += 1
        elif a == "-u": i += 1
        else: raise ValueError(f"unhandled flag {a}")
    if read_mode and not flags["l_set"]:
        kw["min_block_len"] = 42                                  # src/read-filter.cpp:14-16
    flags["genome"] = any(flags[k] for k in "rgemi")               # ultraFastMode = false
    cfg = Config(**kw)
    cfg.count_acgt = flags["g"] or flags["e"]
    return cfg, flags


def parse_fastq(data: bytes) -> list[tuple[str, bytes]]:
    """readFastqRecord (src/input.cpp:113-138): blank (or lone '\\r') lines before a header are
    skipped; the sequence line is taken verbatim (a trailing '\\r' is stripped later by
    ReadTelomereFilter::matches).  Returns (name, raw sequence line)."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    out, i = [], 0
    while i < len(lines):
        if len(lines[i].rstrip(b"\r")) == 0:
            i += 1
            continue
        name = lines[i][1:].rstrip(b"\r").split()[0].decode() if len(lines[i]) > 1 else ""
        out.append((name, lines[i + 1]))
        i += 4
    return out
