"""teloscope_b200 — B200-native telomeric motif scan (drop-in for Teloscope's scanSegment path).

This package is a thin ctypes binding over the C ABI in include/teloscope_b200.h
(libteloscope_b200.so: hand-written CUDA for sm_100a).  There is no CPU path: importing works
anywhere (so the symbol table can be checked), but creating a Scanner without the built library
or without a B200 raises.

Reference interface mirrored here (all `path:line` relative to vgl-hub/teloscope):
  * ScanConfig            <- UserInputTeloscope scan fields, include/input.h:31-48
  * expand_patterns()     <- expandPatternsWithOrientation, src/tools.cpp:201-283
  * Scanner.scan_segment  <- Teloscope::scanSegment, include/teloscope.h:260
  * Scanner.read_matches  <- ReadTelomereFilter::matches, src/read-filter.cpp:37-45
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

__all__ = ["ScanConfig", "Scanner", "BatchResult", "expand_patterns", "load_library", "LIB_PATH",
           "MODE_GENOME", "MODE_TIPS", "MODE_READS", "MATCH_DT", "BLOCK_DT", "WINDOW_DT", "TlsError"]

# TLS_LIB: another build of the library (kernel tuning experiments); the product is the in-tree default
LIB_PATH = Path(os.environ.get("TLS_LIB") or (Path(__file__).resolve().parent / "libteloscope_b200.so"))
ABI_VERSION = 4
MODE_GENOME, MODE_TIPS, MODE_READS = 0, 1, 2
OUT_ACGT, OUT_MATCHES, OUT_ITS, OUT_COMPACT, OUT_PACKED = 1, 2, 4, 8, 16

MATCH_DT = np.dtype([("pos", "<u8"), ("len", "<u2"), ("fwd", "u1"), ("can", "u1"), ("pad", "<u4")])
BLOCK_DT = np.dtype([("start", "<u8"), ("len", "<u4"), ("counts", "<u4"), ("fwd_count", "<u4"),
                     ("rev_count", "<u4"), ("can_count", "<u4"), ("noncan_count", "<u4"),
                     ("total_cov", "<u4"), ("fwd_cov", "<u4"), ("can_cov", "<u4"),
                     ("valid_or", "u1"), ("label", "S1"), ("pad", "<u2")])
WINDOW_DT = np.dtype([("acgt", "<u4", (4,)), ("can_cov", "<u4"), ("noncan_cov", "<u4"),
                      ("fwd_cov", "<u4"), ("rev_cov", "<u4")])
# compact result forms (ScanConfig.compact / TLS_OUT_COMPACT): tls_window16, tls_match8 (pos relative to the item)
WINDOW16_DT = np.dtype([("acgt", "<u2", (4,)), ("can_cov", "<u2"), ("noncan_cov", "<u2"),
                        ("fwd_cov", "<u2"), ("rev_cov", "<u2")])
MATCH8_DT = np.dtype([("pos", "<u4"), ("len", "<u2"), ("fwd", "u1"), ("can", "u1")])
# packed result forms (ScanConfig.packed / TLS_OUT_PACKED): tls_window8 (10-bit base counts, 8-bit match counts, the few
# windows with 255 or more matches listed beside), tls_match4 (u32)
WINDOW8_DT = np.dtype([("w", "<u4", (2,))])
WINDOW_EXC_DT = np.dtype([("window", "<u4"), ("canonical", "<u4"), ("all", "<u4"), ("forward", "<u4")])
assert MATCH_DT.itemsize == 16 and BLOCK_DT.itemsize == 48 and WINDOW_DT.itemsize == 32
assert WINDOW16_DT.itemsize == 16 and MATCH8_DT.itemsize == 8


class TlsError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_patterns", C.c_uint32),
                ("patterns", C.POINTER(C.c_char_p)), ("is_forward", C.POINTER(C.c_uint8)),
                ("is_canonical", C.POINTER(C.c_uint8)),
                ("window_size", C.c_uint32), ("step", C.c_uint32), ("terminal_limit", C.c_uint32),
                ("max_match_dist", C.c_uint16), ("min_block_len", C.c_uint16),
                ("max_block_dist", C.c_uint16), ("min_block_counts", C.c_uint16),
                ("min_block_density", C.c_float), ("outputs", C.c_uint32)]


class _Batch(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("offsets", C.c_void_p), ("abs_pos", C.c_void_p),
                ("n_items", C.c_uint32), ("mode", C.c_uint32), ("seq_on_device", C.c_uint32),
                ("pad", C.c_uint32)]


class _Result(C.Structure):
    _fields_ = [("n_items", C.c_uint32), ("mode", C.c_uint32),
                ("window_off", C.c_void_p), ("windows", C.c_void_p),
                ("terminal_off", C.c_void_p), ("terminal", C.c_void_p),
                ("its_off", C.c_void_p), ("its", C.c_void_p),
                ("canonical_off", C.c_void_p), ("canonical", C.c_void_p),
                ("noncanonical_off", C.c_void_p), ("noncanonical", C.c_void_p),
                ("n_canonical", C.c_void_p), ("n_fwd", C.c_void_p), ("n_rev", C.c_void_p),
                ("pass_", C.c_void_p),
                ("scan_kernel_ms", C.c_float), ("device_ms", C.c_float),
                ("n_bases_scanned", C.c_uint64), ("n_matches", C.c_uint64),
                ("n_kernel_launches", C.c_uint32), ("lean_kernel_ms", C.c_float),
                ("windows16", C.c_void_p), ("canonical8", C.c_void_p), ("noncanonical8", C.c_void_p),
                ("windows8", C.c_void_p), ("canonical4", C.c_void_p), ("noncanonical4", C.c_void_p),
                ("pattern_len", C.c_uint32), ("n_window_exc", C.c_uint32), ("window_exc", C.c_void_p)]


EXPORTS = ("tls_expand_patterns", "tls_device_count", "tls_create", "tls_destroy", "tls_submit",
           "tls_collect", "tls_release", "tls_last_error")

_lib = None


def load_library():
    """dlopen libteloscope_b200.so (built in-tree by `make -C teloscope_b200/csrc` or
    __graft_entry__.build()).  Raises if it is missing — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TlsError(f"{LIB_PATH} not built; run `make -C {LIB_PATH.parent / 'csrc'}` "
                       "(this package has no CPU implementation)")
    L = C.CDLL(str(LIB_PATH))
    L.tls_last_error.restype = C.c_char_p
    L.tls_expand_patterns.restype = C.c_int
    L.tls_expand_patterns.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t,
                                      C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_size_t)]
    L.tls_device_count.restype = C.c_int
    L.tls_create.restype = C.c_int
    L.tls_create.argtypes = [C.POINTER(_Config), C.c_int, C.POINTER(C.c_void_p)]
    L.tls_destroy.argtypes = [C.c_void_p]
    L.tls_submit.restype = C.c_int
    L.tls_submit.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(C.c_int)]
    L.tls_collect.restype = C.c_int
    L.tls_collect.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Result)]
    L.tls_release.restype = C.c_int
    L.tls_release.argtypes = [C.c_void_p, C.c_int]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise TlsError(f"tls error {rc}: {load_library().tls_last_error().decode()}")


def revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def expand_patterns(raw: list[str], edit_distance: int, canonical_fwd: str):
    """-> list of (pattern, is_forward, is_canonical) exactly as the reference's
    expandPatternsWithOrientation + Teloscope ctor produce them."""
    L = load_library()
    n, need = C.c_uint32(), C.c_size_t()
    csv = ",".join(raw).encode()
    _check(L.tls_expand_patterns(csv, edit_distance, canonical_fwd.encode(), None, 0, None, None, 0,
                                 C.byref(n), C.byref(need)))
    text = C.create_string_buffer(max(need.value, 1))
    fwd = (C.c_uint8 * max(n.value, 1))()
    can = (C.c_uint8 * max(n.value, 1))()
    _check(L.tls_expand_patterns(csv, edit_distance, canonical_fwd.encode(), text, need.value, fwd, can,
                                 n.value, C.byref(n), C.byref(need)))
    pats = text.raw[:need.value].split(b"\0")[:n.value]
    return [(p.decode(), bool(fwd[i]), bool(can[i])) for i, p in enumerate(pats)]


@dataclass
class ScanConfig:
    """Scan-relevant fields of UserInputTeloscope after main() processed the flags
    (include/input.h:31-48, src/main.cpp:287-296 for -c, :617-643 for the pattern table)."""
    canonical: str = "TTAGGG"                 # -c (either strand; the lexicographically smaller is Fwd)
    patterns: list[str] | None = None         # -p seeds; None = canonical pair
    edit_distance: int = 1                    # -x
    window_size: int = 1000                   # -w
    step: int = 1000                          # -s
    terminal_limit: int = 50000               # -t
    max_match_dist: int = 50                  # -k
    min_block_len: int = 300                  # -l (42 in read modes unless given)
    max_block_dist: int = 500                 # -d
    min_block_counts: int = 2
    min_block_density: float = 0.5            # -y
    out_acgt: bool = True                     # -g or -e
    out_matches: bool = True                  # -m
    compact: bool = False                     # TLS_OUT_COMPACT: 16-byte window and 8-byte match records
    packed: bool = False                      # TLS_OUT_PACKED: 8-byte window and 4-byte match records (one pattern length, -w <= 1023)
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


class BatchResult:
    """Per-item views of one collected batch.  With copy=True (default) the arrays are numpy copies
    independent of the ctx; with copy=False they alias ctx-owned pinned memory and stay valid until
    release() — the zero-copy path a host pipeline would use."""

    def __init__(self, r: _Result, copy: bool = True, on_release=None, abs_pos=None):
        self.n_items, self.mode = r.n_items, r.mode
        self._on_release = on_release
        self.scan_kernel_ms = float(r.scan_kernel_ms)
        self.device_ms = float(r.device_ms)
        self.n_bases_scanned = int(r.n_bases_scanned)
        self.n_matches = int(r.n_matches)
        self.n_kernel_launches = int(r.n_kernel_launches)
        self.lean_kernel_ms = float(r.lean_kernel_ms)
        self.pattern_len = int(r.pattern_len)
        # the array views: at once for copies; on first use for views of ctx-owned memory (a streaming caller that only
        # wants some of them, or none before release(), pays for none)
        self._raw = (r, copy, abs_pos)
        if copy:
            self._build()

    def __getattr__(self, name):                  # reached only when the attribute is not there yet
        if name.startswith("_") or self.__dict__.get("_raw") is None:
            raise AttributeError(name)
        self._build()
        try:
            return self.__dict__[name]
        except KeyError:
            raise AttributeError(name) from None

    def _build(self):
        r, copy, abs_pos = self._raw
        self._raw = None
        n = r.n_items

        def arr(ptr, count, dt):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dt)
            nbytes = int(count) * np.dtype(dt).itemsize
            view = np.frombuffer((C.c_char * nbytes).from_address(ptr), dtype=dt)
            return view.copy() if copy else view

        def off(ptr):
            return arr(ptr, n + 1, "<u8") if ptr else np.zeros(n + 1, dtype="<u8")

        self.window_off = off(r.window_off)
        n_win = self.window_off[-1] if r.mode == MODE_GENOME else 0
        if r.windows8:
            w8 = arr(r.windows8, n_win, WINDOW8_DT)
            self.window_exc = arr(r.window_exc, r.n_window_exc, WINDOW_EXC_DT)
            if copy:                      # widened to the full record: covered bp = count x pattern length
                w, L = w8["w"].astype(np.uint32), self.pattern_len
                full = np.zeros(len(w8), dtype=WINDOW_DT)
                full["acgt"][:, 0] = w[:, 0] & 1023; full["acgt"][:, 1] = (w[:, 0] >> 10) & 1023; full["acgt"][:, 2] = (w[:, 0] >> 20) & 1023
                full["acgt"][:, 3] = (w[:, 0] >> 30) | ((w[:, 1] & 255) << 2)
                can, allm, fwd = (w[:, 1] >> 8) & 255, (w[:, 1] >> 16) & 255, w[:, 1] >> 24
                if len(self.window_exc):  # 255 = "255 or more": the exact counts are listed
                    x = self.window_exc
                    can[x["window"]] = x["canonical"]; allm[x["window"]] = x["all"]; fwd[x["window"]] = x["forward"]
                full["can_cov"] = can * L; full["noncan_cov"] = (allm - can) * L
                full["fwd_cov"] = fwd * L; full["rev_cov"] = (allm - fwd) * L
                self.windows = full
            else:
                self.windows = w8
        else:
            self.windows = arr(r.windows16, n_win, WINDOW16_DT) if r.windows16 else arr(r.windows, n_win, WINDOW_DT)
        self.terminal_off = off(r.terminal_off)
        self.terminal = arr(r.terminal, self.terminal_off[-1], BLOCK_DT)
        self.its_off = off(r.its_off)
        self.its = arr(r.its, self.its_off[-1], BLOCK_DT)
        self.canonical_off = off(r.canonical_off)
        self.noncanonical_off = off(r.noncanonical_off)

        def matches(ptr16, ptr8, offsets, ptr4=None, canonical=False):
            """16-byte records as they are; 8-byte / 4-byte records (item-relative positions) are widened to MATCH_DT with
            absolute positions when copy=True, and left as MATCH8_DT / u32 views otherwise."""
            if ptr4:
                m4 = arr(ptr4, offsets[-1], "<u4")
                if not copy:
                    return m4
                out = np.zeros(len(m4), dtype=MATCH_DT)
                out["len"] = self.pattern_len; out["fwd"] = m4 >> 31; out["can"] = 1 if canonical else 0
                base = np.zeros(n, dtype=np.uint64) if abs_pos is None else np.asarray(abs_pos, dtype=np.uint64)
                out["pos"] = (m4 & 0x7FFFFFFF).astype(np.uint64) + np.repeat(base, np.diff(offsets).astype(np.int64))
                return out
            if not ptr8:
                return arr(ptr16, offsets[-1], MATCH_DT)
            m8 = arr(ptr8, offsets[-1], MATCH8_DT)
            if not copy:
                return m8
            out = np.zeros(len(m8), dtype=MATCH_DT)
            for f in ("len", "fwd", "can"):
                out[f] = m8[f]
            base = np.zeros(n, dtype=np.uint64) if abs_pos is None else np.asarray(abs_pos, dtype=np.uint64)
            counts = np.diff(offsets).astype(np.int64)
            out["pos"] = m8["pos"].astype(np.uint64) + np.repeat(base, counts)
            return out

        self.canonical = matches(r.canonical, r.canonical8, self.canonical_off, r.canonical4, True)
        self.noncanonical = matches(r.noncanonical, r.noncanonical8, self.noncanonical_off, r.noncanonical4, False)
        self.n_canonical = arr(r.n_canonical, n, "<u4")
        self.n_fwd = arr(r.n_fwd, n, "<u4")
        self.n_rev = arr(r.n_rev, n, "<u4")
        self.passed = arr(r.pass_, n, "u1").astype(bool)

    def release(self):
        """Give the ticket's buffers back to the ctx (only needed after collect(copy=False))."""
        self._raw = None                          # views not built yet would alias memory that is being handed back
        if self._on_release is not None:
            self._on_release()
            self._on_release = None

    def item(self, i: int) -> dict:
        """What SegmentData holds for item i (include/teloscope.h:139-148)."""
        s = lambda a, o: a[int(o[i]):int(o[i + 1])]
        return {"windows": s(self.windows, self.window_off), "terminal": s(self.terminal, self.terminal_off),
                "its": s(self.its, self.its_off), "canonical": s(self.canonical, self.canonical_off),
                "noncanonical": s(self.noncanonical, self.noncanonical_off),
                "n_canonical": int(self.n_canonical[i]) if len(self.n_canonical) else 0}


class Scanner:
    """One pattern table + parameter set bound to one GPU (a `Teloscope` instance)."""

    def __init__(self, cfg: ScanConfig, device: int = 0):
        L = load_library()
        self.cfg = cfg
        self.table = expand_patterns(cfg.raw_patterns(), cfg.edit_distance, cfg.canonical_fwd)
        n = len(self.table)
        pats = (C.c_char_p * n)(*[p.encode() for p, _, _ in self.table])
        fwd = (C.c_uint8 * n)(*[int(f) for _, f, _ in self.table])
        can = (C.c_uint8 * n)(*[int(c) for _, _, c in self.table])
        outputs = (OUT_ACGT if cfg.out_acgt else 0) | (OUT_MATCHES if cfg.out_matches else 0) | OUT_ITS | \
            (OUT_COMPACT if cfg.compact else 0) | (OUT_PACKED if cfg.packed else 0)
        c = _Config(ABI_VERSION, n, pats, fwd, can, cfg.window_size, cfg.step, cfg.terminal_limit,
                    cfg.max_match_dist, cfg.min_block_len, cfg.max_block_dist, cfg.min_block_counts,
                    float(np.float32(cfg.min_block_density)), outputs)
        h = C.c_void_p()
        _check(L.tls_create(C.byref(c), device, C.byref(h)))
        self._h, self._L = h, L
        self._keep = {}
        self._abs_pos = {}               # ticket -> abs_pos of the batch (compact match records are item-relative)

    def close(self):
        if getattr(self, "_h", None):
            self._L.tls_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batch interface (the C ABI, one call per batch) ------------------------------------
    def submit(self, seq, offsets: np.ndarray, abs_pos: np.ndarray | None = None, mode: int = MODE_GENOME,
               device_ptr: int | None = None) -> int:
        """seq: bytes / numpy uint8 array on the host, or pass device_ptr (a CUDA address on this
        scanner's device, e.g. torch_tensor.data_ptr()) for a device-resident batch."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        keep = [offsets]
        b = _Batch()
        if device_ptr is not None:
            b.seq, b.seq_on_device = device_ptr, 1
        else:
            if isinstance(seq, (bytes, bytearray)):
                seq = np.frombuffer(seq, dtype=np.uint8)
            seq = np.ascontiguousarray(seq, dtype=np.uint8)
            keep.append(seq)
            b.seq, b.seq_on_device = seq.ctypes.data, 0
        b.offsets = offsets.ctypes.data
        if abs_pos is not None:
            abs_pos = np.ascontiguousarray(abs_pos, dtype=np.uint64)
            keep.append(abs_pos)
            b.abs_pos = abs_pos.ctypes.data
        b.n_items, b.mode = n, mode
        t = C.c_int()
        _check(self._L.tls_submit(self._h, C.byref(b), C.byref(t)))
        self._keep[t.value] = keep
        self._abs_pos[t.value] = abs_pos
        return t.value

    def collect(self, ticket: int, copy: bool = True) -> BatchResult:
        r = _Result()
        _check(self._L.tls_collect(self._h, ticket, C.byref(r)))

        def rel():
            if self._h:
                _check(self._L.tls_release(self._h, ticket))
            self._keep.pop(ticket, None)

        out = BatchResult(r, copy=copy, on_release=None if copy else rel, abs_pos=self._abs_pos.pop(ticket, None))
        if copy:
            rel()
        return out

    def scan_batch(self, seqs: list[bytes], abs_pos=None, mode: int = MODE_GENOME) -> BatchResult:
        offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
        buf = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
        return self.collect(self.submit(buf, offsets, abs_pos, mode))

    # ---- the reference's per-call interface ----------------------------------------------------
    def scan_segment(self, sequence: bytes, abs_pos: int = 0, tips_only: bool = False) -> dict:
        """Teloscope::scanSegment(sequence, absPos, tipsOnly) -> SegmentData (as a dict of arrays)."""
        r = self.scan_batch([sequence], np.array([abs_pos], dtype=np.uint64), MODE_TIPS if tips_only else MODE_GENOME)
        return r.item(0)

    def read_matches(self, sequence: bytes) -> bool:
        """ReadTelomereFilter::matches(sequence)."""
        return bool(self.scan_batch([sequence], None, MODE_READS).passed[0])
