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


def run_case(tmp: Path, fixture: str, data: bytes, args):
    indir, outdir = tmp / "in", tmp / "out"
    indir.mkdir(exist_ok=True)
    outdir.mkdir(exist_ok=True)
    for f in outdir.iterdir():
        f.unlink()
    src = indir / fixture
    src.write_bytes(gzip.compress(data, 1) if fixture.endswith(".gz") else data)
    p = subprocess.run([str(BIN), "-f", str(src), *args, "-o", str(outdir)], capture_output=True, timeout=600)
    return p.returncode, p.stdout.decode(), p.stderr.decode(), read_files(outdir)


def test_pipeline_equals_reference_outputs(golden, tmp_path):
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    assert BIN.exists(), "build first: make -C teloscope_b200/csrc"
    cases, fixtures = golden
    done = 0
    for c in cases:
        if c["kind"] != "fasta":
            continue
        cfg, flags = ol.config_from_args(c["args"])
        if not device_supports(cfg, flags["genome"]):
            continue
        rc, out, err, files = run_case(tmp_path, c["fixture"], fixtures[c["fixture"]].tobytes(), c["args"])
        tag = (c["fixture"], c["args"])
        assert rc == 0, (tag, err)
        assert out == c["stdout"], tag
        assert err == c["stderr"], tag
        assert sorted(files) == sorted(c["files"]), tag
        for name, lines in c["files"].items():
            assert files[name] == lines, (tag, name)
        done += 1
    assert done >= 170


def test_pipeline_refuses_what_it_does_not_cover(tmp_path):
    gfa = tmp_path / "g.gfa"
    gfa.write_text("H\tVN:Z:1.2\nS\ts1\tACGT\n")
    p = subprocess.run([str(BIN), "-f", str(gfa), "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 2 and b"GFA input" in p.stderr
    fa = tmp_path / "a.fa"
    fa.write_text(">s\nACGT\n")
    p = subprocess.run([str(BIN), "--fastq-subset", str(fa)], capture_output=True)
    assert p.returncode == 2 and b"--fastq-subset" in p.stderr
    p = subprocess.run([str(BIN), "-f", str(fa), "-w", "10", "-s", "20", "-o", str(tmp_path)], capture_output=True)
    assert p.returncode == 1 and b"Step size (20) cannot be larger than window size (10)" in p.stderr
