#!/usr/bin/env python3
"""bench.py — genome-wide telomeric motif scan throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale F] [--scaling weak|strong]

One "step" = one pass of the hot path (Teloscope::scanSegment in genome-wide mode, flags
-r -g -e -m -i, -c TTAGGG -p TTAGGG,TCAGGG,TGAGGG,TTGGGG) over one synthetic human-shaped
assembly: 24 chromosomes with GRCh38 lengths (3.09 Gbp), N-gaps (which split them into contigs),
telomere caps + TVR, ITS islands, 10 % soft-masked.  --scaling weak (default): every rank scans its
own assembly; --scaling strong: ONE assembly, its contigs partitioned over the ranks by bases
(teloscope_b200.shard.partition_items).  No data-path collective either way; rank 0 prints ONE JSON line.

  value    Gbp/s with the assembly already resident in HBM (device pointer handed to the C ABI),
           everything else of the job included (kernels, result copies, host-side assembly), one ticket at a
           time (--depth 2: two tickets in flight; measured no faster, the persistent scan CTAs of step i + 1 keep
           the block kernels of step i off the SMs — value_one_ticket is always reported beside it)
  e2e      same call with the sequence in pinned HOST memory: H2D + kernels + D2H in the timed region
  roofline scan kernel only: algorithmic bytes / CUDA-event duration against the measured HBM peak
  pipeline_e2e  file -> files: the same assembly as a FASTA file through teloscope_b200/teloscope-b200 (host ingest, C ABI,
           writers), process start to exit, with the unmodified reference binary timed on the same file beside it
  other_workloads  BASELINE.json configs[3] (plant-shaped, 7-mers) and configs[4] (read scorer), device-resident
  cpu_baseline  the unmodified reference binary (oracle/_ref/teloscope) on the box's host cores on
           a 1/16-scale assembly of the same generator (file parsing and writers included: it has
           no other entry point): median of 3 runs with -j <cores>, one run with -j 1, CPU model,
           compiler and zlib versions; falls back to the oracle port when the
           binary is absent
  --impl reference  the same reference binary on the SAME assembly as the repo arm (same --scale), every
           step one whole run of it
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Gbp/s scanned (genome-wide mode)"
PATTERNS = ["TTAGGG", "TCAGGG", "TGAGGG", "TTGGGG"]
REF_FLAGS = ["-c", "TTAGGG", "-p", ",".join(PATTERNS), "-r", "-g", "-e", "-m", "-i"]
GRCH38_MBP = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28,
              114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23]
CHR_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]


# --------------------------------------------------------------------------------------------------
# synthetic assembly (SURVEY.md §8d, S-human)
# --------------------------------------------------------------------------------------------------
def plan_assembly(seed: int, scale: float, names=None, mbps=None, patterns=None, cap_unit: str = "TTAGGG"):
    """Host-side plan: per chromosome the contig intervals (N-gaps removed) and the patches
    (telomere caps, TVR, ITS) to stamp onto random sequence.  Defaults: the human-shaped assembly
    (SURVEY.md 8d S-human); names / mbps / patterns / cap_unit give the plant-shaped one (S-plant)."""
    rng = np.random.default_rng(seed)
    patterns = patterns or PATTERNS
    variants = patterns + [revcomp(p) for p in patterns]
    cap_fwd, cap_rev = cap_unit.encode(), revcomp(cap_unit).encode()
    chroms = []
    for name, mbp in zip(names or CHR_NAMES, mbps or GRCH38_MBP):
        n = max(200_000, int(mbp * 1e6 * scale))
        gaps = []
        n_gaps = int(rng.integers(5, 31))
        cen = int(rng.integers(1_000_000, 3_000_001) * min(1.0, scale * 4))
        cen_start = int(n * rng.uniform(0.3, 0.6))
        gaps.append((cen_start, max(1000, cen)))
        for _ in range(n_gaps):
            gaps.append((int(rng.integers(100_000, n - 100_000)), int(rng.integers(100, 50_001))))
        gaps.sort()
        contigs, pos = [], 0
        for g, ln in gaps:
            if g <= pos + 1000:
                continue
            contigs.append((pos, g))
            pos = min(n, g + ln)
        if pos < n - 1000:
            contigs.append((pos, n))
        patches = []   # (chromosome position, bytes)
        def tvr(unit, bp):
            out = bytearray()
            while len(out) < bp:
                u = bytearray(unit)
                u[int(rng.integers(0, len(u)))] = b"ACGT"[int(rng.integers(0, 4))]
                out += u
            return bytes(out)
        cap = int(rng.integers(2000, 15001)); tv = int(rng.integers(500, 2001))
        patches.append((0, cap_rev * (cap // len(cap_rev)) + tvr(cap_rev, tv)))
        cap = int(rng.integers(2000, 15001)); tv = int(rng.integers(500, 2001))
        tail = tvr(cap_fwd, tv) + cap_fwd * (cap // len(cap_fwd))
        patches.append((n - len(tail), tail))
        for _ in range(int(rng.integers(20, 61))):
            unit = variants[int(rng.integers(0, len(variants)))].encode()
            reps = int(rng.integers(4, 41))
            patches.append((int(rng.integers(200_000, n - 200_000)), unit * reps))
        chroms.append({"name": name, "len": n, "contigs": contigs, "patches": patches})
    return chroms


def revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def build_assembly(chroms, device, seed: int):
    """-> (seq uint8 CUDA tensor with 4 KiB of slack, offsets[n+1], abs_pos[n]) : the contigs of all
    chromosomes back to back, exactly what the reference's walkPath hands to scanSegment."""
    import torch
    offsets, abs_pos = [0], []
    for c in chroms:
        for a, b in c["contigs"]:
            abs_pos.append(a)
            offsets.append(offsets[-1] + (b - a))
    total = offsets[-1]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    seq = torch.empty(total + 4096, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    step = 1 << 26
    for s in range(0, total + 4096, step):
        e = min(total + 4096, s + step)
        seq[s:e] = lut[torch.randint(0, 4, (e - s,), device=device, generator=g, dtype=torch.int64)]
    # stamp patches: translate chromosome coordinates to buffer coordinates (skip parts in gaps)
    pos_list, val_list = [], []
    k = 0
    for c in chroms:
        starts = np.array([a for a, _ in c["contigs"]]); ends = np.array([b for _, b in c["contigs"]])
        base = np.array(offsets[k:k + len(starts)])
        for p, data in c["patches"]:
            idx = np.arange(p, p + len(data))
            ci = np.searchsorted(starts, idx, side="right") - 1
            ok = (ci >= 0) & (idx < ends[np.clip(ci, 0, len(ends) - 1)])
            pos_list.append((base[ci[ok]] + (idx[ok] - starts[ci[ok]])).astype(np.int64))
            val_list.append(np.frombuffer(data, dtype=np.uint8)[ok])
        k += len(starts)
    pos = torch.from_numpy(np.concatenate(pos_list)).to(device)
    val = torch.from_numpy(np.concatenate(val_list)).to(device)
    seq[pos] = val
    # soft-mask ~10 % in stretches of 1-50 kb
    rng = np.random.default_rng(seed + 1)
    masked, target = 0, total // 10
    while masked < target:
        ln = int(rng.integers(1000, 50001)); a = int(rng.integers(0, max(1, total - ln)))
        seq[a:a + ln] |= 0x20
        masked += ln
    return seq, np.array(offsets, dtype=np.uint64), np.array(abs_pos, dtype=np.uint64)


def write_fasta(chroms, seq_host: np.ndarray, offsets: np.ndarray, path: Path) -> int:
    """The same assembly as a FASTA file with its N-gaps (80-column lines) for the reference."""
    k, bases = 0, 0
    with open(path, "wb") as fh:
        for c in chroms:
            rec = np.full(c["len"], ord("N"), dtype=np.uint8)
            for a, b in c["contigs"]:
                rec[a:b] = seq_host[int(offsets[k]):int(offsets[k + 1])]
                bases += b - a
                k += 1
            fh.write(b">" + c["name"].encode() + b"\n")
            full = (len(rec) // 80) * 80
            if full:
                lines = np.empty((full // 80, 81), dtype=np.uint8)
                lines[:, :80] = rec[:full].reshape(-1, 80)
                lines[:, 80] = ord("\n")
                fh.write(lines.tobytes())
            if full < len(rec):
                fh.write(rec[full:].tobytes() + b"\n")
    return bases


# --------------------------------------------------------------------------------------------------
# other BASELINE.json configs, measured beside the headline one (rank 0, N = 1): device-resident input
# --------------------------------------------------------------------------------------------------
PLANT_PATTERNS = ["CCCTAAA", "CCCTAAG", "CCCTAAC", "CCCTAAT", "CCCTAGA"]


def build_reads(device, seed: int, total_bp: int):
    """S-reads: HiFi-shaped reads, length ~N(15000, 3000) clipped at 1 kb, i.i.d. ACGT,
    2 % with a 120 bp - 5 kb telomeric prefix or suffix.  -> (seq CUDA tensor, offsets, n_telomeric)"""
    import torch
    rng = np.random.default_rng(seed)
    lens = np.maximum(1000, rng.normal(15000, 3000, size=int(total_bp / 15000 * 1.05) + 8).astype(np.int64))
    lens = lens[: int(np.searchsorted(np.cumsum(lens), total_bp)) + 1]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    total = int(offsets[-1])
    g = torch.Generator(device=device); g.manual_seed(seed)
    seq = torch.empty(total + 4096, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    for s0 in range(0, total + 4096, 1 << 26):
        e = min(total + 4096, s0 + (1 << 26))
        seq[s0:e] = lut[torch.randint(0, 4, (e - s0,), device=device, generator=g, dtype=torch.int64)]
    telo = rng.choice(len(lens), size=max(1, len(lens) // 50), replace=False)
    pos_list, val_list = [], []
    for i in telo:
        ln = int(min(rng.integers(120, 5001), lens[i]))
        if rng.integers(0, 2):
            unit, a = b"CCCTAA", int(offsets[i])
        else:
            unit, a = b"TTAGGG", int(offsets[i + 1]) - ln
        pos_list.append(np.arange(a, a + ln, dtype=np.int64))
        val_list.append(np.frombuffer((unit * (ln // 6 + 1))[:ln], dtype=np.uint8))
    seq[torch.from_numpy(np.concatenate(pos_list)).to(device)] = torch.from_numpy(np.concatenate(val_list)).to(device)
    return seq, offsets, len(telo)


def time_workload(tb, cfg, seq, offsets, abs_pos, mode, steps: int, warmup: int, device: int):
    """-> (seconds per step, mean scan-kernel ms, last result summary) with the input resident in HBM."""
    import torch
    sc = tb.Scanner(cfg, device=device)
    try:
        scan_ms, last = [], None
        for k in range(warmup + steps):
            if k == warmup:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            res = sc.collect(sc.submit(None, offsets, abs_pos, mode, device_ptr=seq.data_ptr()), copy=False)
            if k >= warmup:
                scan_ms.append(res.scan_kernel_ms)
            if k == warmup + steps - 1:
                last = {"matches": int(res.n_matches), "launches_per_step": int(res.n_kernel_launches)}
                if mode == tb.MODE_READS:
                    last["reads_passed"] = int(np.count_nonzero(res.passed))
                else:
                    last["windows"] = int(res.window_off[-1]); last["terminal_blocks"] = len(res.terminal); last["its_blocks"] = len(res.its)
            res.release()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps, float(np.mean(scan_ms)), last
    finally:
        sc.close()


def other_workloads(tb, device: int, seed: int, scale: float, peak: float):
    """configs[3] (plant-shaped assembly, 7-mers) and configs[4] (read subset scorer) of BASELINE.json,
    a few steps each; rooflines by the same byte accounting as the headline workload."""
    import torch
    out = []
    dev = f"cuda:{device}"
    # ---- S-plant: 2.5 Gbp in 12 records, -c CCCTAAA + 4 variants, genome-wide ----------------------------
    chroms = plan_assembly(seed + 100, scale, [f"chr{i}" for i in range(1, 13)], [2500.0 / 12] * 12, PLANT_PATTERNS, "TTTAGGG")
    seq, offsets, abs_pos = build_assembly(chroms, dev, seed + 100)
    nb = int(offsets[-1])
    cfg = tb.ScanConfig(canonical="CCCTAAA", patterns=PLANT_PATTERNS, out_acgt=True, out_matches=True)
    per, kms, info = time_workload(tb, cfg, seq, offsets, abs_pos, tb.MODE_GENOME, 3, 2, device)
    algo = nb + 32 * info["windows"]
    out.append({"workload": f"synthetic plant-shaped assembly x{scale:g} (12 records, TTTAGGG caps), -c CCCTAAA -p {','.join(PLANT_PATTERNS)} "
                            f"-r -g -e -m -i ({len(tb.expand_patterns(PLANT_PATTERNS, 1, 'CCCTAAA'))} 7-mer patterns)",
                "bases": nb, "value": nb / per / 1e9, "unit": "Gbp/s", "ms_per_step": per * 1e3, "scan_kernels_ms": kms,
                "roofline_frac": (algo / (kms * 1e-3) / 1e9) / peak, **info})
    del seq
    torch.cuda.empty_cache()
    # ---- S-reads: HiFi-shaped reads through the read scorer (--fastq-subset / --bam-subset path, -l 42) ----------
    total = int(2.0e9 * min(1.0, scale))
    seq, offsets, n_telo = build_reads(dev, seed + 200, total)
    nb = int(offsets[-1])
    cfg = tb.ScanConfig(min_block_len=42)
    per, kms, info = time_workload(tb, cfg, seq, offsets, None, tb.MODE_READS, 3, 2, device)
    out.append({"workload": f"{len(offsets) - 1} synthetic HiFi reads (15 kb mean, {n_telo} with a telomeric end), read subset scorer, -l 42, default motif",
                "bases": nb, "reads": len(offsets) - 1, "value": nb / per / 1e9, "unit": "Gbp/s", "reads_per_s": (len(offsets) - 1) / per,
                "ms_per_step": per * 1e3, "scan_kernels_ms": kms,
                "roofline_frac": ((nb + len(offsets) - 1) / (kms * 1e-3) / 1e9) / peak, **info})
    return out


def pipeline_file_to_files(seed: int, scale: float, with_reference: bool):
    """End to end INCLUDING host I/O: the headline assembly written as a FASTA file (80-column
    lines, N-gaps) to /dev/shm, then `teloscope_b200/teloscope-b200 -f file <flags> -o dir` from process start to
    exit with every output file closed — and, beside it, the unmodified reference binary on the same file with
    all host threads, and a byte comparison of the two output directories."""
    import torch
    exe = ROOT / "teloscope_b200" / "teloscope-b200"
    ref_bin = ROOT / "oracle" / "_ref" / "teloscope"
    if not exe.exists():
        return {"unavailable": "teloscope_b200/teloscope-b200 not built"}
    chroms = plan_assembly(seed, scale)
    seq, offsets, _ = build_assembly(chroms, "cuda", seed)
    seq_host = seq.cpu().numpy()
    del seq
    torch.cuda.empty_cache()
    tmp = Path(tempfile.mkdtemp(prefix="telo_pipe_", dir="/dev/shm" if Path("/dev/shm").exists() else None))
    try:
        fa = tmp / "assembly.fa"
        bases = write_fasta(chroms, seq_host, offsets, fa)
        del seq_host
        out = {"what": "FASTA file (page cache) -> all output files of -r -g -e -m -i, process start to exit",
               "bases": bases, "fasta_bytes": fa.stat().st_size, "unit": "Gbp/s"}
        times = []
        for i in range(3):                                   # first run warms the page cache and the driver
            d = tmp / "ours"
            shutil.rmtree(d, ignore_errors=True); d.mkdir()
            t0 = time.perf_counter()
            subprocess.run([str(exe), "-f", str(fa), *REF_FLAGS, "-o", str(d)], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
            times.append(time.perf_counter() - t0)
        out["seconds"] = min(times[1:]); out["value"] = bases / out["seconds"] / 1e9
        if with_reference and ref_bin.exists():
            cores = os.cpu_count() or 1
            d = tmp / "ref"
            d.mkdir()
            t0 = time.perf_counter()
            subprocess.run([str(ref_bin), "-f", str(fa), *REF_FLAGS, "-j", str(cores), "-o", str(d)], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            same = all((tmp / "ours" / f.name).exists() and f.read_bytes() == (tmp / "ours" / f.name).read_bytes()
                       for f in d.iterdir())
            out["reference_same_file"] = {"seconds": dt, "value": bases / dt / 1e9, "cores": cores,
                                          "outputs_byte_identical": bool(same), "files_compared": len(list(d.iterdir()))}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def fastq_subset_file_to_file(seed: int, total_bp: int, with_reference: bool):
    """BASELINE.json configs[4], file to file: S-reads written as FASTQ to /dev/shm, `teloscope-b200 --fastq-subset`
    (process start to exit, subset to a file) and the unmodified reference on the same file; outputs compared."""
    import torch
    exe = ROOT / "teloscope_b200" / "teloscope-b200"
    ref_bin = ROOT / "oracle" / "_ref" / "teloscope"
    if not exe.exists():
        return {"unavailable": "teloscope_b200/teloscope-b200 not built"}
    seq, offsets, n_telo = build_reads("cuda", seed, total_bp)
    host = seq.cpu().numpy()
    del seq
    torch.cuda.empty_cache()
    tmp = Path(tempfile.mkdtemp(prefix="telo_fq_", dir="/dev/shm" if Path("/dev/shm").exists() else None))
    try:
        fq = tmp / "reads.fq"
        qual = b"I" * int(np.max(np.diff(offsets)))
        with open(fq, "wb") as fh:
            for i in range(len(offsets) - 1):
                a, b = int(offsets[i]), int(offsets[i + 1])
                fh.write(b"@read%d\n" % i); fh.write(host[a:b].tobytes()); fh.write(b"\n+\n"); fh.write(qual[: b - a]); fh.write(b"\n")
        bases = int(offsets[-1])
        out = {"what": "FASTQ file (page cache) -> FASTQ subset file, --fastq-subset -l 42, process start to exit",
               "bases": bases, "reads": len(offsets) - 1, "fastq_bytes": fq.stat().st_size, "unit": "Gbp/s"}
        times = []
        for i in range(3):
            d = tmp / "ours"
            shutil.rmtree(d, ignore_errors=True); d.mkdir()
            t0 = time.perf_counter()
            subprocess.run([str(exe), "--fastq-subset", "-f", str(fq), "-o", str(d)], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
            times.append(time.perf_counter() - t0)
        out["seconds"] = min(times[1:]); out["value"] = bases / out["seconds"] / 1e9
        if with_reference and ref_bin.exists():
            cores = os.cpu_count() or 1
            d = tmp / "ref"
            d.mkdir()
            t0 = time.perf_counter()
            subprocess.run([str(ref_bin), "--fastq-subset", "-f", str(fq), "-j", str(cores), "-o", str(d)], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            a, b = (d / "reads.fq_telomeric.fastq").read_bytes(), (tmp / "ours" / "reads.fq_telomeric.fastq").read_bytes()
            out["reference_same_file"] = {"seconds": dt, "value": bases / dt / 1e9, "cores": cores,
                                          "outputs_byte_identical": a == b, "subset_bytes": len(a)}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def write_unmapped_bam(path: Path, host: np.ndarray, offsets: np.ndarray, threads: int) -> int:
    """S-reads as an unmapped BAM (flag 4, qualities 0xFF): BGZF blocks of 65 280 payload bytes deflated like the
    reference's writer does (raw deflate, default level, memLevel 8), in a thread pool (zlib releases the GIL)."""
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    code = np.full(256, 15, dtype=np.uint8)
    for i, ch in enumerate(b"=ACMGRSVTWYHKDBN"):
        code[ch] = i
        code[ord(chr(ch).lower())] = i
    header = b"BAM\x01" + struct.pack("<i", 0) + struct.pack("<i", 0)
    chunks = [header]
    for i in range(len(offsets) - 1):
        a, b = int(offsets[i]), int(offsets[i + 1])
        n = b - a
        c = code[host[a:b]]
        if n & 1:
            c = np.append(c, 0)
        packed = ((c[0::2] << 4) | c[1::2]).astype(np.uint8).tobytes()
        name = b"read%d\x00" % i
        core = struct.pack("<iiBBHHHiiii", -1, -1, len(name), 0, 4680, 0, 4, n, -1, -1, 0)
        body = core + name + packed + b"\xff" * n
        chunks.append(struct.pack("<i", len(body)) + body)
    payload = b"".join(chunks)
    del chunks

    def block(k):
        data = payload[k:k + 65280]
        co = zlib.compressobj(zlib.Z_DEFAULT_COMPRESSION, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY)
        comp = co.compress(data) + co.flush()
        total = 18 + len(comp) + 8
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", total - 1) + comp +
                struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))

    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool, open(path, "wb") as fh:
        for blk in pool.map(block, range(0, len(payload), 65280)):
            fh.write(blk)
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return len(payload)


def bam_subset_file_to_file(seed: int, total_bp: int, with_reference: bool):
    """BASELINE.json configs[4], BAM half: S-reads as an unmapped BAM in /dev/shm, `teloscope-b200 --bam-subset -o`
    (process start to exit) and the unmodified reference on the same file; the two subset BAMs compared byte for byte."""
    import torch
    exe = ROOT / "teloscope_b200" / "teloscope-b200"
    ref_bin = ROOT / "oracle" / "_ref" / "teloscope"
    if not exe.exists():
        return {"unavailable": "teloscope_b200/teloscope-b200 not built"}
    seq, offsets, n_telo = build_reads("cuda", seed, total_bp)
    host = seq.cpu().numpy()
    del seq
    torch.cuda.empty_cache()
    cores = os.cpu_count() or 1
    tmp = Path(tempfile.mkdtemp(prefix="telo_bam_", dir="/dev/shm" if Path("/dev/shm").exists() else None))
    try:
        bam = tmp / "reads.bam"
        payload = write_unmapped_bam(bam, host, offsets, cores)
        del host
        bases = int(offsets[-1])
        out = {"what": "unmapped BAM (page cache) -> subset BAM, --bam-subset -l 42 -o dir, process start to exit",
               "bases": bases, "reads": len(offsets) - 1, "bam_bytes": bam.stat().st_size, "bam_payload_bytes": payload, "unit": "Gbp/s"}
        times = []
        for i in range(3):
            d = tmp / "ours"
            shutil.rmtree(d, ignore_errors=True); d.mkdir()
            t0 = time.perf_counter()
            subprocess.run([str(exe), "--bam-subset", "-f", str(bam), "-o", str(d)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            times.append(time.perf_counter() - t0)
        out["seconds"] = min(times[1:]); out["value"] = bases / out["seconds"] / 1e9
        if with_reference and ref_bin.exists():
            d = tmp / "ref"
            d.mkdir()
            t0 = time.perf_counter()
            subprocess.run([str(ref_bin), "--bam-subset", "-f", str(bam), "-j", str(cores), "-o", str(d)], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            a, b = (d / "reads_telomeric.bam").read_bytes(), (tmp / "ours" / "reads_telomeric.bam").read_bytes()
            out["reference_same_file"] = {"seconds": dt, "value": bases / dt / 1e9, "cores": cores,
                                          "outputs_byte_identical": a == b, "subset_bytes": len(a)}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# --------------------------------------------------------------------------------------------------
def host_info() -> dict:
    """CPU model, core count, compiler and zlib versions."""
    import zlib
    model = ""
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    try:
        gxx = subprocess.run(["g++", "--version"], capture_output=True, text=True).stdout.splitlines()[0]
    except (OSError, IndexError):
        gxx = "unknown"
    return {"cpu_model": model, "cores": os.cpu_count() or 1, "compiler": gxx + " -O3 (oracle/Makefile)", "zlib": zlib.ZLIB_RUNTIME_VERSION}


def cpu_reference_run(sample_scale: float, steps: int, warmup: int, seed: int = 1234, single_thread_scale: float = 0.0):
    """Times the reference's own CPU implementation on an assembly of the same generator at `sample_scale`.
    Returns (Gbp/s from the MEDIAN step, seconds of that step, description dict)."""
    import torch
    cores = os.cpu_count() or 1
    chroms = plan_assembly(seed, sample_scale)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    seq, offsets, abs_pos = build_assembly(chroms, dev, seed)
    seq_host = seq.cpu().numpy()
    del seq
    ref_bin = ROOT / "oracle" / "_ref" / "teloscope"
    tmp = Path(tempfile.mkdtemp(prefix="telo_ref_", dir="/dev/shm" if Path("/dev/shm").exists() else None))
    try:
        if ref_bin.exists():
            fa = tmp / "sample.fa"
            bases = write_fasta(chroms, seq_host, offsets, fa)
            times = []
            for i in range(warmup + steps):
                out = tmp / f"out{i}"
                out.mkdir()
                t0 = time.perf_counter()
                subprocess.run([str(ref_bin), "-f", str(fa), *REF_FLAGS, "-j", str(cores), "-o", str(out)],
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                dt = time.perf_counter() - t0
                shutil.rmtree(out)
                if i >= warmup:
                    times.append(dt)
            kind = "reference"
            how = (f"oracle/_ref/teloscope (unmodified reference, g++ -O3) -j {cores} on a {sample_scale:g}-scale "
                   f"assembly of the same generator: {bases / 1e6:.0f} Mbp in 24 records, page-cache-warm FASTA in "
                   f"/dev/shm; wall time of the whole run (parse + scan + writers); median of {steps} run(s) after {warmup} warm-up")
            extra = {}
            if single_thread_scale > 0:                       # -j 1 on a smaller sample (one run: it is 10x slower)
                ch1 = plan_assembly(seed + 1, single_thread_scale)
                s1, o1, _ = build_assembly(ch1, dev, seed + 1)
                fa1 = tmp / "sample_j1.fa"
                b1 = write_fasta(ch1, s1.cpu().numpy(), o1, fa1)
                del s1
                out = tmp / "out_j1"
                out.mkdir()
                t0 = time.perf_counter()
                subprocess.run([str(ref_bin), "-f", str(fa1), *REF_FLAGS, "-j", "1", "-o", str(out)],
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                dt = time.perf_counter() - t0
                extra["single_thread"] = {"value": b1 / dt / 1e9, "unit": "Gbp/s", "j": 1, "bases": b1, "seconds": dt,
                                          "sample": f"{single_thread_scale:g}-scale assembly, one run"}
        else:
            sys.path.insert(0, str(ROOT / "tests"))
            import oracle_lib as ol
            orc = ol.Oracle(ol.Config(patterns=PATTERNS))
            bases = int(offsets[-1])
            up = np.where((seq_host >= 97) & (seq_host <= 122), seq_host - 32, seq_host).astype(np.uint8)
            times = []
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                orc.scan_many(up, offsets, False, cores)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
            kind = "port"
            how = (f"oracle port (oracle/oracle.cpp scanSegment) with {cores} threads, one job per contig, on a "
                   f"{sample_scale:g}-scale assembly: {bases / 1e6:.0f} Mbp")
        per_step = float(np.median(times))
        desc = {"kind": kind, "cores": cores, "j": cores, "sample": how, "bases": bases, "seconds_per_run": [round(t, 3) for t in times], **host_info()}
        if kind == "reference":
            desc.update(extra)
        return bases / per_step / 1e9, per_step, desc
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.path = Path(tempfile.mkstemp(prefix="clocks_", suffix=".csv")[1])
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [l.split(", ") for l in self.path.read_text().splitlines() if l.count(",") >= 6]
        self.path.unlink(missing_ok=True)
        if rows:
            sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
            out["sm_mhz"] = float(np.median(sm)) if sm else None
            out["sm_max_mhz"] = float(rows[0][1]) if rows[0][1].replace(".", "").isdigit() else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].strip() == "Active" for r in rows)]
            out["samples"] = len(rows)
        return out


def bind_to_gpu_numa_node(index: int):
    """Run this rank on the CPU cores NVML lists as local to its GPU, before any pinned buffer is allocated: the
    staging and result buffers then live on the GPU's own NUMA node and the host<->device copies of eight ranks do
    not cross the socket interconnect.  Best effort (a box without NVML or with one node is left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="assembly size factor (1.0 = 3.09 Gbp; tests use less)")
    ap.add_argument("--cpu-scale", type=float, default=1.0 / 16, help="size factor of the cpu_baseline sample of the repo arm")
    ap.add_argument("--ref-scale", type=float, default=None, help="--impl reference: size factor of its assembly (default: --scale, the repo arm's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one assembly per rank; strong: one assembly, its contigs partitioned over the ranks")
    ap.add_argument("--depth", type=int, default=1, choices=[1, 2], help="tickets in flight in the timed device-resident loop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-records", action="store_true", help="32-byte window / 16-byte match records instead of the packed forms")
    ap.add_argument("--compact-records", action="store_true", help="16-byte window / 8-byte match records (TLS_OUT_COMPACT) instead of the packed forms")
    ap.add_argument("--no-extra", action="store_true", help="skip the plant-assembly and read-scorer measurements")
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"synthetic human-shaped assembly x{args.scale:g} (24 chromosomes, GRCh38 lengths, N-gaps, telomere caps, "
                f"ITS, 10% soft-masked), -c TTAGGG -p {','.join(PATTERNS)} -r -g -e -m -i (124 patterns, w=s=1000)")

    if args.impl == "reference":
        if rank != 0:
            return
        ref_scale = args.scale if args.ref_scale is None else args.ref_scale
        workload = workload.replace(f"x{args.scale:g}", f"x{ref_scale:g}")
        gbps, per_step, desc = cpu_reference_run(ref_scale, args.steps, args.warmup, seed=args.seed)
        print(json.dumps({"metric": METRIC, "value": gbps, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
                          "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "impl": "reference", "config": {"workload": workload, "scale": ref_scale,
                                                          "what": "one whole run of the reference binary per step (FASTA parse + scan + writers), the same assembly every step"},
                          "cpu_baseline": {"value": gbps, "unit": "Gbp/s", **desc},
                          "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    bind_to_gpu_numa_node(local)
    import torch
    import teloscope_b200 as tb
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: one assembly per rank (weak), or this rank's share of the contigs of ONE assembly (strong) --------
    strong = args.scaling == "strong" and world > 1
    aseed = args.seed if strong else args.seed + rank
    chroms = plan_assembly(aseed, args.scale)
    seq, offsets, abs_pos = build_assembly(chroms, f"cuda:{local}", aseed)
    job_bases = int(offsets[-1])
    if strong:
        from teloscope_b200.shard import partition_items
        lens = np.diff(offsets.astype(np.int64))
        mine = partition_items(lens.tolist(), world)[rank]
        parts = [seq[int(offsets[i]):int(offsets[i + 1])] for i in mine]
        new_off = np.concatenate([[0], np.cumsum(lens[mine])]).astype(np.uint64)
        seq = torch.cat(parts + [torch.full((4096,), 65, dtype=torch.uint8, device=seq.device)])
        abs_pos = abs_pos[mine]
        offsets = new_off
        del parts
        torch.cuda.empty_cache()
    n_bases = int(offsets[-1])
    host = torch.empty(seq.numel(), dtype=torch.uint8, pin_memory=True)
    host.copy_(seq)
    host_np = host.numpy()
    # compact result records (TLS_OUT_COMPACT: 16 bytes per window, 8 per -m match): half the bytes over PCIe
    sc = tb.Scanner(tb.ScanConfig(canonical="TTAGGG", patterns=PATTERNS, out_acgt=True, out_matches=True,
                                  compact=args.compact_records and not args.full_records,
                                  packed=not args.full_records and not args.compact_records), device=local)

    # collect(copy=False): results are read in place from the ctx's pinned buffers (what a C++ host
    # would do through the C ABI); release() hands the ticket back
    def run_steps(k, device_resident, stats=None, depth=1):
        """k steps with `depth` tickets in flight.  depth 1: submit, collect, release, one step after the
        other (kernel timings are then undisturbed).  depth 2 drives the double-buffered C ABI as a
        streaming caller would: step i+1 is submitted (host plan + H2D + scan enqueue) before step i is
        collected, so its H2D overlaps the kernels of step i.  Every step is collected inside the call."""
        pending, last = [], None

        def finish(ticket, keep):
            res = sc.collect(ticket, copy=False)
            if stats is not None:
                stats["scan_ms"].append(res.scan_kernel_ms)
                stats["lean_ms"].append(res.lean_kernel_ms)
                stats["launches"] += res.n_kernel_launches
            if not keep:
                res.release()
            return res

        for _ in range(k):
            if len(pending) == depth:
                finish(pending.pop(0), False)
            if device_resident:
                pending.append(sc.submit(None, offsets, abs_pos, tb.MODE_GENOME, device_ptr=seq.data_ptr()))
            else:
                pending.append(sc.submit(host_np[:n_bases], offsets, abs_pos, tb.MODE_GENOME))
        while pending:
            last = finish(pending.pop(0), not pending)
        return last

    r = run_steps(max(3, args.warmup), True)
    n_windows, n_matches = int(r.window_off[-1]), r.n_matches
    r.release()

    # ---- timed: device-resident input -------------------------------------------------------------------
    clocks = ClockSampler(local) if rank == 0 else None
    # one ticket at a time: kernel timings are undisturbed (this is where the roofline numbers come from) ...
    stats = {"scan_ms": [], "lean_ms": [], "launches": 0}
    barrier()
    t0 = time.perf_counter()
    r = run_steps(args.steps, True, stats)
    barrier()
    elapsed_one = time.perf_counter() - t0
    r.release()
    scan_ms, launches = stats["scan_ms"], stats["launches"]
    # ... and the way a streaming caller drives the double-buffered ABI: step i + 1 is submitted before step i is collected,
    # so the result copies of step i overlap the kernels of step i + 1
    elapsed = elapsed_one
    if args.depth == 2:
        run_steps(2, True, depth=2).release()
        stats2 = {"scan_ms": [], "lean_ms": [], "launches": 0}
        barrier()
        t0 = time.perf_counter()
        r = run_steps(args.steps, True, stats2, depth=2)
        barrier()
        elapsed = time.perf_counter() - t0
        r.release()
        launches = stats2["launches"]
    # ---- timed: end to end from pinned host memory ----------------------------------------------------
    run_steps(3, False, depth=2).release()       # warm-up of the host-buffer path: both tickets' buffers get allocated
    barrier()
    t1 = time.perf_counter()
    r = run_steps(args.steps, False, depth=2)
    barrier()
    elapsed_e2e = time.perf_counter() - t1
    clk = clocks.stop() if clocks else None
    d2h = n_windows * r.windows.dtype.itemsize + (len(r.canonical) + len(r.noncanonical)) * r.canonical.dtype.itemsize + \
        (len(r.terminal) + len(r.its)) * 48 + r.n_items * 32 + 192
    h2d = n_bases + len(offsets) * 40 + (n_bases // 49136 + len(offsets)) * 16

    my_kernel_gbps = n_bases / (float(np.mean(scan_ms)) * 1e-3) / 1e9          # this rank's scan kernels alone
    t = torch.tensor([elapsed, elapsed_e2e, float(n_bases), my_kernel_gbps, elapsed_one], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed, elapsed_e2e, total_bases, kernel_gbps_all, elapsed_one = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3]), float(mx[4])
    else:
        total_bases, kernel_gbps_all = float(n_bases), my_kernel_gbps
    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        # 1 byte per base read + 32 bytes per window + 8 bytes per EMITTED match (canonical and
        # terminal non-canonical, the -m lists) + 32 bytes per block.  The kernel also writes an 8-byte record
        # for every other match (intermediate data for the block kernels): not algorithmic, listed beside it.
        n_emitted = len(r.canonical) + len(r.noncanonical)
        n_blocks = len(r.terminal) + len(r.its)
        algo_bytes = n_bases + 32 * n_windows + 8 * n_emitted + 32 * n_blocks
        moved_bytes = n_bases + 32 * n_windows + 8 * n_matches
        kern_s = float(np.mean(scan_ms)) * 1e-3
        achieved = algo_bytes / kern_s / 1e9
        # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the
        # committed `ncu --set full` capture (taken on a 0.25-scale assembly of the same generator; the kernel
        # streams, so bytes per base carry over), scaled to the bases one launch processes here.
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "rows_kernel_traffic.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            traffic = int(tj["dram_bytes_per_base"] * n_bases)
            traffic_src = f"{tj['dram_bytes_per_base']:.3f} B/base x {n_bases} bases; {tj['source']}"
        line = {
            "metric": METRIC, "value": total_bases * args.steps / elapsed / 1e9, "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": elapsed / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "value_one_ticket": total_bases * args.steps / elapsed_one / 1e9, "ms_per_step_one_ticket": elapsed_one / args.steps * 1e3,
            "config": {"workload": workload, "scale": args.scale, "bases_per_gpu": n_bases, "contigs_per_gpu": len(offsets) - 1,
                       "windows_per_gpu": n_windows, "match_records_per_gpu": n_matches,
                       "sharding": (f"ONE assembly ({job_bases} bases), contigs partitioned over {world} ranks by bases (LPT), no collective" if strong
                                    else "one assembly per GPU, no collective"),
                       "l2": "input (>= 3 GB at scale 1) is larger than the 126 MB L2; no flush needed",
                       "pipelining": (f"value: {args.depth} ticket(s) in flight (the double-buffered C ABI as a streaming caller drives it; "
                                      "value_one_ticket: submit, collect, release one step after the other); e2e: two tickets in flight "
                                      "(H2D of step i+1 overlaps step i)"),
                       "result_records": "full (32-byte windows, 16-byte matches)" if args.full_records else
                                         ("compact (TLS_OUT_COMPACT: 16-byte windows, 8-byte item-relative matches)" if args.compact_records else
                                          "packed (TLS_OUT_PACKED: 8-byte windows of 10-bit base counts and 8-bit match counts, 4-byte item-relative matches)")},
            "e2e": {"value": total_bases * args.steps / elapsed_e2e / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "kernel_only": {"scan_kernels_ms": float(np.mean(scan_ms)), "lean_kernel_ms": float(np.mean(stats["lean_ms"])),
                            "Gbp_per_s": n_bases / kern_s / 1e9, "Gbp_per_s_all_ranks": kernel_gbps_all,
                            "note": "rank 0's kernels; all_ranks = sum over ranks of bases / own scan-kernel time (device-side scaling: "
                                    "the whole-step value at N > 1 is bound by the result copies of all ranks sharing the host link)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "tls::rowScanKernel<3,true> (+ the two small launches that take "
                         "contigs shorter than 8 windows); duration = all three, CUDA events on the ctx stream, mean over "
                         "the timed steps", "algorithmic_bytes": algo_bytes, "bytes_per_base": algo_bytes / n_bases,
                         "achieved_incl_intermediate_match_records": moved_bytes / kern_s / 1e9,
                         "peak_source": peak_src},
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            gbps, _, desc = cpu_reference_run(args.cpu_scale, 3, 1, single_thread_scale=args.cpu_scale / 4)
            line["cpu_baseline"] = {"value": gbps, "unit": "Gbp/s", **desc}
    sc.close()
    if rank == 0:
        if world == 1 and not args.no_extra:
            del seq, host, host_np
            torch.cuda.empty_cache()
            try:
                line["other_workloads"] = other_workloads(tb, local, args.seed, args.scale, peak)
            except Exception as exc:           # the headline line must not be lost to a side measurement
                line["other_workloads"] = {"error": repr(exc)}
            try:
                line["pipeline_e2e"] = pipeline_file_to_files(args.seed, args.scale, not args.no_cpu_baseline)
            except Exception as exc:
                line["pipeline_e2e"] = {"error": repr(exc)}
            try:
                line["fastq_subset_e2e"] = fastq_subset_file_to_file(args.seed + 300, int(1.0e9 * min(1.0, args.scale)), not args.no_cpu_baseline)
            except Exception as exc:
                line["fastq_subset_e2e"] = {"error": repr(exc)}
            try:
                line["bam_subset_e2e"] = bam_subset_file_to_file(args.seed + 400, int(0.5e9 * min(1.0, args.scale)), not args.no_cpu_baseline)
            except Exception as exc:
                line["bam_subset_e2e"] = {"error": repr(exc)}
            line["config5_note"] = ("configs[4] names 100 Gbp of reads: measured here on 1 Gbp (FASTQ) / 0.5 Gbp (BAM) files and on a 2 Gbp device-resident "
                                    "batch (other_workloads); both front ends stream in bounded batches, so 100 Gbp is these rates extrapolated, not measured")
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
