"""The reference's OWN acceptance harness pointed at teloscope-b200.

`make -C oracle ref` stages the reference's fixtures, expected outputs and test scripts next to the reference binary it
builds (oracle/_ref/harness/, a git-ignored build directory that travels to the GPU box).  Here a scratch tree is laid out
the way those scripts expect it (build/bin/teloscope = a wrapper around teloscope-b200, build/bin/teloscope-validate = the
reference's validator behind a wrapper that runs a directory of manifests in six processes), and they run unmodified:
  * teloscope-validate validateFiles  (src/validate.cpp:845-887: every .tst, legacy stdout goldens and directive cases)
  * .github/workflows/val.sh          (file counts per flag set, GFA annotation checks)
  * scripts/test_gaps_bed.sh, scripts/test_n50.sh
  * scripts/test_bam_subset.py        (BGZF / BAM subset integration, byte-level)
The pass / fail table goes to gpurun_out/harness_table.md; cases that exercise front ends this binary refuses by design
(record filters, --plot-report, FASTA on stdin) are listed in EXPECTED_UNSUPPORTED and must be the ONLY failures."""
import os
import re
import shutil
import stat
import subprocess
import tempfile
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
HARNESS = ROOT / "oracle" / "_ref" / "harness"
VALIDATE = ROOT / "oracle" / "_ref" / "teloscope-validate"
BIN = ROOT / "teloscope_b200" / "teloscope-b200"

VALIDATE_WRAPPER = r"""#!/bin/bash
real="$(dirname "$0")/teloscope-validate.real"
if [ $# -eq 1 ] && [ -d "$1" ]; then
  files=( "$1"/*.tst ); n=6; pids=(); rc=0; tmp=$(mktemp -d)
  for k in $(seq 0 $((n - 1))); do
    sub=(); for ((i = k; i < ${#files[@]}; i += n)); do sub+=("${files[$i]}"); done
    [ ${#sub[@]} -eq 0 ] && continue
    "$real" "${sub[@]}" > "$tmp/out.$k" 2> "$tmp/err.$k" & pids+=($!)
  done
  for p in "${pids[@]}"; do wait "$p" || rc=1; done
  cat "$tmp"/out.*; cat "$tmp"/err.* >&2; rm -rf "$tmp"; exit $rc
fi
exec "$real" "$@"
"""

# .tst cases of front ends teloscope-b200 does not have (it exits with status 2 and says so)
EXPECTED_UNSUPPORTED = re.compile(r"filter|include|exclude|prefix|plot|stdin")


@pytest.fixture(scope="module")
def tree():
    import teloscope_b200
    if teloscope_b200.load_library().tls_device_count() < 1:
        pytest.fail("no sm_100 device visible: GPU tests must run on a B200")
    if not HARNESS.exists() or not VALIDATE.exists():
        pytest.skip("oracle/_ref/harness not staged (make -C oracle ref needs /root/reference)")
    tmp = Path(tempfile.mkdtemp(prefix="telo_harness_"))
    for d in ("validateFiles", "testFiles", "scripts"):
        shutil.copytree(HARNESS / d, tmp / d)
    (tmp / ".github" / "workflows").mkdir(parents=True)
    shutil.copy(HARNESS / ".github" / "workflows" / "val.sh", tmp / ".github" / "workflows" / "val.sh")
    binDir = tmp / "build" / "bin"
    binDir.mkdir(parents=True)
    wrapper = binDir / "teloscope"
    wrapper.write_text(f"#!/bin/sh\nexec {BIN} \"$@\"\n")
    wrapper.chmod(wrapper.stat().st_mode | stat.S_IEXEC)
    # build/bin/teloscope-validate: the reference's validator, unmodified, behind a wrapper that hands a DIRECTORY of manifests
    # to six validator processes side by side (the validator takes files as well as directories, src/validate.cpp:842-843,
    # and gives every case a temporary directory of its own).  val.sh runs the whole directory too: a case is mostly process
    # start + CUDA context, and 2 x 173 of them one after the other are most of the GPU suite's wall time.
    shutil.copy(VALIDATE, binDir / "teloscope-validate.real")
    shard = binDir / "teloscope-validate"
    shard.write_text(VALIDATE_WRAPPER)
    shard.chmod(shard.stat().st_mode | stat.S_IEXEC)
    yield tmp
    shutil.rmtree(tmp, ignore_errors=True)


def _table(name, rows):
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "harness_table.md", "a") as fh:
        fh.write(f"\n## {name}\n\n| case | result |\n|---|---|\n")
        for case, res in rows:
            fh.write(f"| {case} | {res} |\n")


def _strip(s):
    return re.sub(r"\x1b\[[0-9;]*m", "", s)


def test_validate_files(tree):
    p = subprocess.run(["build/bin/teloscope-validate", "validateFiles"], cwd=tree, capture_output=True, text=True, timeout=3000)
    n_manifests = len(list((tree / "validateFiles").glob("*.tst")))
    rows = []
    for line in _strip(p.stdout).splitlines():
        m = re.match(r"\s*(PASS|FAIL)\s+(\S+)", line)
        if m:
            rows.append((Path(m.group(2)).name, m.group(1)))
    rows.sort()
    _table("teloscope-validate validateFiles", rows)
    (ROOT / "gpurun_out" / "harness_validate.log").write_text(_strip(p.stdout) + "\n---- stderr ----\n" + p.stderr)
    assert len(rows) >= 170 and len({c for c, _ in rows}) == n_manifests, p.stdout[-3000:]
    bad = [c for c, r in rows if r == "FAIL" and not EXPECTED_UNSUPPORTED.search(c)]
    assert not bad, (bad, p.stdout[-3000:])


def test_val_sh(tree):
    p = subprocess.run(["bash", ".github/workflows/val.sh"], cwd=tree, capture_output=True, text=True, timeout=3000)
    rows = [(m.group(2).strip(), m.group(1)) for m in re.finditer(r"^(PASS|FAIL)\s+(.*)$", _strip(p.stdout), re.M)]
    _table(".github/workflows/val.sh", rows)
    (ROOT / "gpurun_out" / "harness_val_sh.log").write_text(_strip(p.stdout) + "\n---- stderr ----\n" + p.stderr)
    bad = [c for c, r in rows if r == "FAIL" and not EXPECTED_UNSUPPORTED.search(c)]
    assert not bad and len(rows) > 200, (bad, len(rows), p.stdout[-2000:], p.stderr[-2000:])


@pytest.mark.parametrize("script", ["test_gaps_bed.sh", "test_n50.sh"])
def test_shell_scripts(tree, script):
    p = subprocess.run(["bash", f"scripts/{script}"], cwd=tree, capture_output=True, text=True, timeout=3000)
    _table(f"scripts/{script}", [(script, "PASS" if p.returncode == 0 else "FAIL")])
    assert p.returncode == 0, (_strip(p.stdout)[-3000:], p.stderr[-2000:])


def test_bam_subset_script(tree):
    env = dict(os.environ, TELOSCOPE=str(tree / "build" / "bin" / "teloscope"))
    p = subprocess.run(["python3", "scripts/test_bam_subset.py"], cwd=tree, capture_output=True, text=True, timeout=3000, env=env)
    _table("scripts/test_bam_subset.py", [("test_bam_subset.py", "PASS" if p.returncode == 0 else "FAIL")])
    (ROOT / "gpurun_out" / "harness_bam.log").write_text(p.stdout + "\n---- stderr ----\n" + p.stderr)
    assert p.returncode == 0 and "PASS BAM subset integration" in p.stdout, (p.stdout[-2000:], p.stderr[-3000:])
