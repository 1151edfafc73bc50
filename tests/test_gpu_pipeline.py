"""The host pipeline binary (teloscope_b200/teloscope-b200: FASTA ingest -> C ABI -> writers, SURVEY.md 8f
rows 1-2) against what the UNMODIFIED reference binary printed and wrote for the same command lines
(tests/golden/ref_cases.json.gz): stdout, stderr and every output file, byte for byte, for every golden
FASTA case the device supports (the reference's own .tst flag sets included).  .gz fixtures are fed
gzip-compressed, so the zlib path is the one that runs."""
import gzip
import subprocess
from pathlib import Path

import pytest

import oracle_lib as ol
from test_gpu_parity import device_supports

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "teloscope_b200" / "teloscope-b200"


def read_files(outdir: Path) -> dict:
    """Same keys as tests/golden/make_golden.py::read_files: the file name after '<input name>_'."""
    files = {}
    for f in sorted(outdir.iterdir()):
        suffix = f.name.split(".fa", 1)[-1] if ".fa" in f.name else f.name
        suffix = suffix.replace(".gz_", "_")
        files[suffix.lstrip("_")] = f.read_text().splitlines()
    return files


def run_case(tmp: Path, fixture: str, data: bytes, args, env=None):
    indir, outdir = tmp / "in", tmp / "out"
    indir.mkdir(exist_ok=True)
    outdir.mkdir(exist_ok=True)
    for f in outdir.iterdir():
        f.unlink()
    src = indir / fixture
    src.write_bytes(gzip.compress(data, 1) if fixture.endswith(".gz") else data)
    import os
    p = subprocess.run([str(BIN), "-f", str(src), *args, "-o", str(outdir)], capture_output=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    return p.returncode, p.stdout.decode(), p.stderr.decode(), read_files(outdir)


def test_pipeline_equals_reference_outputs(golden, tmp_path):
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    assert BIN.exists(), "build first: make -C teloscope_b200/csrc"
    cases, fixtures = golden
    todo = []
    for c in cases:
        if c["kind"] != "fasta":
            continue
        cfg, flags = ol.config_from_args(c["args"])
        if device_supports(cfg, flags["genome"]):
            todo.append(c)
    assert len(todo) >= 170
    data = {name: fixtures[name].tobytes() for name in sorted({c["fixture"] for c in todo})}   # the .npz reader is not thread-safe

    def one(k):
        # every third case with the FASTA parser cutting records into 97-byte pieces (runs stitched across the cuts)
        c = todo[k]
        d = tmp_path / f"case{k}"
        d.mkdir()
        return run_case(d, c["fixture"], data[c["fixture"]], c["args"], env={"TLS_PARSE_PIECE": "97"} if k % 3 == 0 else None)

    # a case is mostly process start + CUDA context: six at a time, each in a directory of its own
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=6) as pool:
        results = list(pool.map(one, range(len(todo))))
    for c, (rc, out, err, files) in zip(todo, results):
        tag = (c["fixture"], c["args"])
        assert rc == 0, (tag, err)
        assert out == c["stdout"], tag
        assert err == c["stderr"], tag
        assert sorted(files) == sorted(c["files"]), tag
        for name, lines in c["files"].items():
            assert files[name] == lines, (tag, name)


def test_fastq_subset_equals_reference_output(golden, tmp_path):
    """--fastq-subset: stdout (the kept records, verbatim) by SHA-256 and stderr against the reference binary's for its
    own FASTQ fixtures (CRLF, blank lines, thresholds), kept read names for the seeded synthetic reads; plus the -o form."""
    import hashlib
    cases, fixtures = golden
    done = 0
    for c in cases:
        if c["kind"] != "fastq":
            continue
        src = tmp_path / c["fixture"]
        src.write_bytes(fixtures[c["fixture"]].tobytes())
        p = subprocess.run([str(BIN), "--fastq-subset", *c["args"], str(src)], capture_output=True, timeout=600)
        tag = (c["fixture"], c["args"])
        assert p.returncode == 0, (tag, p.stderr)
        kept = [l[1:].split()[0] for l in p.stdout.decode().replace("\r", "").splitlines()[0::4] if l]
        assert kept == c["kept"], tag
        if "stdout_sha" in c:
            assert hashlib.sha256(p.stdout).hexdigest() == c["stdout_sha"], tag
        assert p.stderr.decode() == c["stderr"], tag
        done += 1
    assert done >= 15
    c = next(c for c in cases if c["kind"] == "fastq" and "stdout_sha" in c and c["kept"])
    src = tmp_path / (c["fixture"] + ".gz")
    src.write_bytes(gzip.compress(fixtures[c["fixture"]].tobytes()))
    outdir = tmp_path / "fq_out"
    outdir.mkdir()
    p = subprocess.run([str(BIN), "--fastq-subset", *c["args"], "-f", str(src), "-o", str(outdir)], capture_output=True, timeout=600)
    assert p.returncode == 0 and p.stdout == b"", p.stderr
    written = outdir / (src.name + "_telomeric.fastq")
    assert hashlib.sha256(written.read_bytes()).hexdigest() == c["stdout_sha"]
    assert f"Wrote telomeric reads to {written}.".encode() in p.stderr
    bad = tmp_path / "bad.fq"
    bad.write_text("@r1\nACGT\n+\nIII\n")
    p = subprocess.run([str(BIN), "--fastq-subset", str(bad)], capture_output=True)
    assert p.returncode == 1 and b"Error: FASTQ record 1: sequence and quality length differ." in p.stderr


def test_pipeline_refuses_what_it_does_not_cover(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">s\nACGT\n")
    p = subprocess.run([str(BIN), "-f", str(fa), "--include-prefix", "s", "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 2 and b"record filtering" in p.stderr
    fa.write_text(">s\n" + "ACGT" * 10 + "\n")
    p = subprocess.run([str(BIN), "--bam-subset", str(fa)], capture_output=True)
    assert p.returncode == 1 and b"BAM subset failed: input is not BGZF-compressed BAM" in p.stderr
    p = subprocess.run([str(BIN), "-f", str(fa), "-w", "10", "-s", "20", "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 1 and b"Step size (20) cannot be larger than window size (10)" in p.stderr


def test_gfa_annotation_equals_reference_binary(tmp_path):
    """GFA input: annotated GFA, colours CSV, stdout and stderr of teloscope-b200 against the unmodified
    reference binary (-j 1: the reference's node order depends on thread completion otherwise) for every GFA fixture of the
    reference, with the flag sets of its own .tst cases and with the defaults."""
    ref = ROOT / "oracle" / "_ref" / "teloscope"
    fixtures = sorted((ROOT / "oracle" / "_ref" / "harness" / "testFiles").glob("*.gfa"))
    if not ref.exists() or not fixtures:
        pytest.skip("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    done = annotated = 0
    for f in fixtures:
        extra = {"gfa_plant_small": ["-c", "CCCTAAA"], "gfa_internal_only_small": ["-t", "100"]}.get(f.stem, [])
        for flags in (["-x", "0", "-l", "60", *extra], extra):
            outs = {}
            for tag, exe in (("ref", ref), ("ours", BIN)):
                d = tmp_path / f"{f.stem}_{tag}_{len(flags)}"
                d.mkdir()
                p = subprocess.run([str(exe), "-f", str(f), "-j", "1", *flags, "-o", str(d)], capture_output=True, timeout=600)
                assert p.returncode == 0, (tag, f.name, p.stderr)
                outs[tag] = (p.stdout, p.stderr, {x.name: x.read_bytes() for x in d.iterdir()})
            assert outs["ours"] == outs["ref"], (f.name, flags)
            annotated += b"telomere_" in outs["ref"][2][f.name + ".telo.annotated.gfa"]
            done += 1
    assert done >= 30 and annotated >= 12


def test_pipeline_over_several_devices_equals_one(tmp_path):
    """teloscope-b200 --gpus N (one context and feeder thread per device, batch k on device k mod N, results consumed in
    input order): every output file, stdout and stderr identical to --gpus 1.  Devices beyond the visible ones wrap around,
    so the multi-context path also runs on a one-GPU box."""
    import numpy as np
    from test_gpu_rows import make_contig
    rng = np.random.default_rng(99)
    fa = tmp_path / "multi.fa"
    with open(fa, "wb") as fh:
        for i in range(30):
            n = int(rng.integers(20_000, 1_500_000))
            s = make_contig(rng, n, "TTAGGG", i % 4 == 0).replace(b"\r", b"A")
            gap = b"N" * int(rng.integers(10, 500)) if i % 3 else b""
            fh.write(b">rec%d some text\n" % i)
            body = s[: n // 2] + gap + s[n // 2:]
            fh.write(b"\n".join(body[k:k + 70] for k in range(0, len(body), 70)) + b"\n")
    outs = {}
    for gpus in (1, 2, 3):
        d = tmp_path / f"out{gpus}"
        d.mkdir()
        env = dict(__import__("os").environ, TLS_BATCH_BYTES=str(3 << 20))
        p = subprocess.run([str(BIN), "-f", str(fa), "-r", "-g", "-e", "-m", "-i", "--gpus", str(gpus), "-o", str(d)], capture_output=True, timeout=900, env=env)
        assert p.returncode == 0, p.stderr[-2000:]
        outs[gpus] = (p.stdout, p.stderr, {f.name: f.read_bytes() for f in d.iterdir()})
    assert len(outs[1][2]) == 11
    assert outs[2] == outs[1] and outs[3] == outs[1]


def test_degenerate_overlap_is_refused_with_the_reason(golden, tmp_path):
    """-s < L or -w minus -s < L (L = pattern length): the reference's scan start of a non-first window is
    min(step - L, overlap - L) in uint32 (src/teloscope.cpp:413-415).  It wraps or lands near the window's end, so most of
    every window is never scanned.  The golden outputs of the unmodified reference show it (windows of a 50 % GC fixture
    with a GC track below 1 %, a T2T chromosome reported with one telomere); the device path does not reproduce that
    arithmetic and says so."""
    cases, fixtures = golden
    seen = 0
    for c in cases:
        if c["kind"] != "fasta":
            continue
        cfg, flags = ol.config_from_args(c["args"])
        if device_supports(cfg, flags["genome"]):
            continue
        seen += 1
        if "window_gc.bedgraph" in c["files"] and cfg.window_size - cfg.step < 6 <= cfg.step:
            gc = [float(l.split("\t")[3]) for l in c["files"]["window_gc.bedgraph"][1:]]
            assert gc[0] > 40 and min(gc[1:]) < 1, gc                  # the reference only looked at the last few bases of those windows
        rc, out, err, files = run_case(tmp_path, c["fixture"], fixtures[c["fixture"]].tobytes(), c["args"])
        assert rc == 1 and "overlapping windows need" in err, (c["args"], err)
    assert seen >= 2
