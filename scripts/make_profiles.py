#!/usr/bin/env python3
"""Turn the files of one scripts/gpu_session.sh call (gpurun_out/<tag>/) into the tracked summaries under profiles/:
    python scripts/make_profiles.py gpurun_out/<tag> r02"""
import csv, json, subprocess, sys
from pathlib import Path
src, tag = Path(sys.argv[1]), sys.argv[2]
out = Path(__file__).resolve().parent.parent / "profiles"

def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout

# bench line
b = (src / "bench.json").read_text().strip().splitlines()[-1]
(out / f"{tag}_bench.json").write_text(json.dumps(json.loads(b), indent=1) + "\n")
# launch list
rows = list(csv.reader(open(src / "launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
(out / f"{tag}_launches.csv").write_text("\n".join(",".join(r) for r in rows[hi:]) + "\n")
with open(out / f"{tag}_launches.md", "w") as fh:
    fh.write(f"# {tag}: ncu launch list of one step on the 3.03 Gbp batch (compact records)\n\n"
             "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:Kernel -s 22 -c 24 python scripts/trace_bench.py 1.0` "
             "(per-launch times are cold-cache and serialised: shares, not absolutes)\n\n| kernel | µs |\n|---|---|\n")
    tot = 0.0
    for r in rows[hi + 1:]:
        if len(r) > vi:
            us = float(r[vi].replace(",", "")) / 1000.0
            tot += us
            fh.write(f"| `{r[ki][:70]}` | {us:.1f} |\n")
    fh.write(f"| total | {tot:.1f} |\n")
for rep, kern, name, what in ((src / "prof_rows.ncu-rep", "rowScanKernel", f"{tag}_rows_summary.md",
                               "window-row scan kernel, ONE launch, 0.25-scale assembly (714 Mbp, 714 K window rows), compact records"),
                              (src / "prof_reads.ncu-rep", "readScanKernel", f"{tag}_reads_summary.md",
                               "read scorer, 2.0 Gbp of synthetic HiFi reads (133 K reads)")):
    if not rep.exists():
        continue
    m = run("python", "scripts/ncu_metrics.py", str(rep))
    l = run("python", "scripts/ncu_lines.py", str(rep), kern, "--top", "40")
    (out / name).write_text(f"# {tag} — {what}: ncu --set full --clock-control none --import-source on\n\n{m}\nSource-line attribution "
                            f"(`scripts/ncu_lines.py`, share of executed warp instructions):\n\n```\n{l}```\n")
print("written to", out)
