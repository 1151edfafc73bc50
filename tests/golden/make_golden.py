#!/usr/bin/env python3
"""Generate tests/golden/{fixtures.npz,ref_cases.json.gz} from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and oracle/_ref/teloscope, built by
`make -C oracle ref`).  The committed outputs travel to the GPU box; this script does not.

What it records
  * fixtures.npz     — the input sequences of the reference's test FASTA/FASTQ files
                       (/root/reference/testFiles), one uint8 array per file, in our own container.
  * ref_cases.json.gz — for every case (fixture + command line):
      - stdout of the reference binary; for the reference's own legacy `.tst` cases it is first
        asserted equal to the golden block stored in validateFiles/*.tst (so the compiled binary
        is itself pinned to the reference's checked-in goldens);
      - the lines of every BED / bedgraph file the binary wrote;
      - for FASTQ cases the names of the reads the reference kept.
"""
import gzip
import json
import os
import random
import shlex
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
BIN = ROOT / "oracle" / "_ref" / "teloscope"

FULL = ["-r", "-g", "-e", "-m", "-i", "-n"]   # every genome-wide output + all terminal blocks

# Extra genome-wide / tips cases on top of the reference's own .tst list.
EXTRA = [
    ("t2t.fa", FULL), ("t2t.fa", ["-r", "-g", "-e", "-m", "-i"]),
    ("t2t.fa", FULL + ["-w", "500", "-s", "250"]), ("t2t.fa", FULL + ["-w", "100", "-s", "30"]),
    ("t2t.fa", FULL + ["-w", "333", "-s", "333"]), ("t2t.fa", FULL + ["-w", "7", "-s", "7"]),
    ("t2t.fa", FULL + ["-w", "1000", "-s", "999"]), ("t2t.fa", FULL + ["-w", "64", "-s", "5"]),
    ("t2t.fa", FULL + ["-x", "2"]), ("t2t.fa", FULL + ["-x", "0"]),
    ("t2t.fa", FULL + ["-p", "TTAGGG,TCAGGG,TGAGGG,TTGGGG"]),
    ("t2t.fa", FULL + ["-p", "TTAGGG,TTTAGGG"]), ("t2t.fa", FULL + ["-p", "TTAGGG,TTAGG", "-x", "0"]),
    ("t2t.fa", FULL + ["-p", "NNAGGG", "-x", "0"]), ("t2t.fa", FULL + ["-t", "700", "-l", "100"]),
    ("plant.fa", FULL + ["-c", "CCCTAAA"]), ("plant.fa", FULL + ["-c", "CCCTAAA", "-p", "TTTAGGG,CCCTAAA"]),
    ("plant.fa", FULL + ["-c", "CCCTAAA", "-p", "CCCTAAA,CCCTAAG,CCCTAAC,CCCTAAT,CCCTAGA"]),
    ("edit_test.fa", FULL + ["-x", "2"]), ("edit_test.fa", FULL + ["-x", "0"]),
    ("mirror_extend.fa", FULL + ["-t", "300"]), ("boundary_extend_its.fa", FULL + ["-t", "300"]),
    ("boundary_zone_shift.fa", FULL + ["-t", "400"]), ("boundary_zone_shift.fa", FULL + ["-t", "1200"]),
    ("boundary_no_terminal.fa", FULL + ["-t", "500"]),
    ("mirror_inverted_long.fa", FULL + ["-t", "1000"]), ("mirror_fwd_end_long.fa", FULL + ["-t", "1000"]),
    ("mirror_rev_start_long.fa", FULL + ["-t", "1000"]),
    ("density_edge.fa", FULL + ["-y", "0.8"]), ("density_edge.fa", FULL + ["-y", "0.3"]),
    ("density_edge.fa", FULL + ["-l", "500", "-y", "0.4", "-k", "10"]),
    ("short_contig.fa", FULL + ["-w", "50", "-s", "50"]), ("its.fa", FULL + ["-t", "100"]),
    ("multi.fa", FULL + ["-k", "5"]), ("multi.fa", FULL + ["-d", "1"]), ("multi.fa", FULL + ["-l", "100"]),
    ("bTaeGut7_chr33_mat.fa.gz", FULL), ("bTaeGut7_chr33_pat.fa.gz", FULL),
    ("bTaeGut7_chr33_mat.fa.gz", ["-n"]), ("bTaeGut7_chr33_pat.fa.gz", ["-n"]),
    ("bTaeGut7_chr33_mat.fa.gz", FULL + ["-p", "TTAGGG,TCAGGG,TGAGGG,TTGGGG"]),
    ("bTaeGut7_chr33_mat.fa.gz", FULL + ["-w", "2000", "-s", "500", "-t", "20000"]),
]

FASTQ_CASES = [
    ("fastq_subset.fq", ["-x", "0", "-l", "18", "-y", "0.8", "-k", "10", "-d", "10"]),
    ("fastq_subset_large.fq", ["-x", "0", "-l", "18", "-y", "0.8", "-k", "10", "-d", "10"]),
    ("fastq_subset_realistic.fq", ["-x", "0", "-l", "18", "-y", "0.8", "-k", "10", "-d", "10"]),
    ("fastq_subset_crlf.fq", ["-x", "0", "-l", "18", "-y", "0.8", "-k", "10", "-d", "10"]),
    ("fastq_subset_threshold.fq", ["-x", "0", "-y", "0.8", "-k", "10", "-d", "10"]),
    ("fastq_subset_threshold.fq", ["-x", "0", "-y", "0.8", "-k", "10", "-d", "10", "-l", "42"]),
    ("fastq_subset_large.fq", []), ("fastq_subset_realistic.fq", ["-x", "1"]),
]


def run_ref(args, outdir, stdin=None):
    p = subprocess.run([str(BIN), *args, "-o", str(outdir)], cwd=str(REF), capture_output=True, input=stdin)
    return p.returncode, p.stdout.decode(), p.stderr.decode()


def read_files(outdir):
    files = {}
    for f in sorted(Path(outdir).iterdir()):
        suffix = f.name.split(".fa", 1)[-1] if ".fa" in f.name else f.name
        suffix = suffix.replace(".gz_", "_")
        files[suffix.lstrip("_")] = f.read_text().splitlines()
    return files


def tst_cases():
    """(fixture, args, expected_stdout) for the reference's legacy FASTA .tst files
    (format: src/validate.cpp:282-350 — line 1 command, line 2 'embedded', rest stdout)."""
    cases = []
    for tst in sorted((REF / "validateFiles").glob("*.tst")):
        lines = tst.read_text().split("\n")
        if len(lines) < 2 or lines[1].strip() != "embedded":
            continue
        toks = shlex.split(lines[0])
        fixture = Path(toks[0]).name
        args, i = [], 1
        while i < len(toks):
            if toks[i] == "-o":
                i += 2
                continue
            if toks[i] == "-f":
                i += 2
                continue
            args.append(toks[i])
            i += 1
        cases.append((fixture, args, "\n".join(lines[2:])))
    return cases


def main():
    if not BIN.exists():
        sys.exit("build the reference first: make -C oracle ref")
    fixtures, cases = {}, []

    def load_fixture(name):
        if name not in fixtures:
            raw = (REF / "testFiles" / name).read_bytes()
            if name.endswith(".gz"):
                raw = gzip.decompress(raw)
            fixtures[name] = np.frombuffer(raw, dtype=np.uint8)
        return fixtures[name]

    seen = set()
    todo = [(f, a, s, "tst") for f, a, s in tst_cases()] + [(f, a, None, "extra") for f, a in EXTRA]
    for fixture, args, want_stdout, origin in todo:
        key = (fixture, tuple(args))
        if key in seen:
            continue
        seen.add(key)
        load_fixture(fixture)
        with tempfile.TemporaryDirectory() as td:
            rc, out, err = run_ref(["-f", f"testFiles/{fixture}", *args], td)
            assert rc == 0, (fixture, args, err)
            if want_stdout is not None:
                assert out.rstrip("\n") == want_stdout.rstrip("\n"), f"reference binary disagrees with {fixture} {args}"
            cases.append({"kind": "fasta", "fixture": fixture, "args": args, "origin": origin,
                          "stdout": out, "stderr": err, "files": read_files(td)})
    print(f"{len(cases)} FASTA cases ({sum(c['origin'] == 'tst' for c in cases)} pinned to .tst goldens)")

    for fixture, args in FASTQ_CASES:
        load_fixture(fixture)
        with tempfile.TemporaryDirectory() as td:
            p = subprocess.run([str(BIN), "--fastq-subset", *args, f"testFiles/{fixture}"], cwd=str(REF),
                               capture_output=True)
            assert p.returncode == 0, p.stderr
            kept = [l[1:].split()[0] for l in p.stdout.decode().replace("\r", "").splitlines()[0::4] if l]
            cases.append({"kind": "fastq", "fixture": fixture, "args": args, "kept": kept,
                          "stdout_sha": __import__("hashlib").sha256(p.stdout).hexdigest(),
                          "stderr": p.stderr.decode()})
    # the reference's expected FASTQ outputs must equal what the binary printed for its own cases
    exp = REF / "testFiles" / "expected"
    for name, fixture, args in (("fastq_subset_large.fq", "fastq_subset_large.fq", FASTQ_CASES[1][1]),
                                ("fastq_subset_realistic.fq", "fastq_subset_realistic.fq", FASTQ_CASES[2][1])):
        p = subprocess.run([str(BIN), "--fastq-subset", *args, f"testFiles/{fixture}"], cwd=str(REF), capture_output=True)
        assert p.stdout == (exp / name).read_bytes(), name

    # Seeded random reads with the option sets of scripts/test_bam_subset.py:384-404 (same
    # generator recipe, our own seed handling) — known-answer vectors for the read scorer.
    gen = random.Random(23)
    seqs = {}
    for index in range(240):
        length = gen.randrange(18, 250)
        s = "".join(gen.choice("ACGTN") for _ in range(length))
        if index % 3 == 0:
            ins = gen.randrange(len(s) + 1)
            s = s[:ins] + gen.choice(("TTAGGG", "CCCTAA")) * gen.randrange(2, 18) + s[ins:]
        if index % 11 == 0:
            s += "TCAGGG" * 8 + "TTAGGG"
        seqs[f"random_{index:03d}"] = s
    kat = {"short": "TTAGGG" * 6, "default_pass": "TTAGGG" * 7, "long": "CCCTAA" * 15, "fail": "ACGT" * 20,
           "one_repeat": "TTAGGG", "exact_12": "TTAGGG" * 2, "flanked_exact": "ACGT" + "CCCTAA" * 2 + "TGCA",
           "exact_18": "TTAGGG" * 3, "two_thirds": "TTAGGGAAAAAATTAGGG", "plant_pass": "TTTAGGG" * 3,
           "vertebrate_fail": "TTAGGG" * 4, "lower": "ttagggttagggttagggttagggttagggttagggttaggg",
           "withN": "TTAGGG" * 4 + "N" + "TTAGGG" * 4}
    seqs.update(kat)
    fq = "".join(f"@{n}\n{s}\n+\n{'I' * len(s)}\n" for n, s in seqs.items()).encode()
    fixtures["synthetic_reads.fq"] = np.frombuffer(fq, dtype=np.uint8)
    option_sets = [[], ["-x", "0", "-l", "18", "-y", "0.8", "-k", "10", "-d", "10"],
                   ["-x", "1", "-l", "42", "-y", "0.5", "-k", "50", "-d", "50"],
                   ["-x", "0", "-l", "60", "-y", "1", "-k", "10", "-d", "10"],
                   ["-x", "0", "-l", "12", "-y", "1", "-k", "10", "-d", "10"],
                   ["-x", "0", "-l", "18", "-y", "1", "-k", "10", "-d", "10"],
                   ["-x", "0", "-l", "18", "-y", "0.666", "-k", "20", "-d", "10"],
                   ["-x", "0", "-l", "18", "-y", "0.667", "-k", "20", "-d", "10"],
                   ["-c", "CCCTAAA", "-x", "0", "-l", "21", "-y", "1"]]
    for args in option_sets:
        p = subprocess.run([str(BIN), "--fastq-subset", *args], cwd="/tmp", capture_output=True, input=fq)
        assert p.returncode == 0, p.stderr
        kept = [l[1:] for l in p.stdout.decode().splitlines()[0::4]]
        cases.append({"kind": "fastq", "fixture": "synthetic_reads.fq", "args": args, "kept": kept,
                      "stderr": p.stderr.decode()})

    np.savez_compressed(HERE / "fixtures.npz", **fixtures)
    with gzip.open(HERE / "ref_cases.json.gz", "wt", compresslevel=9) as fh:
        json.dump({"reference": "vgl-hub/teloscope@0265c683", "cases": cases}, fh)
    print(f"wrote {len(cases)} cases, {len(fixtures)} fixtures")


if __name__ == "__main__":
    main()
