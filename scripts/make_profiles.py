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
    fh.write(f"# {tag}: ncu launch list of one step on the 3.03 Gbp batch (packed records)\n\n"
             "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:Kernel -s 36 -c 18 env TLS_PACKED=1 python scripts/trace_bench.py 1.0` "
             "(per-launch times are cold-cache and serialised: shares, not absolutes)\n\n| kernel | µs |\n|---|---|\n")
    tot = 0.0
    for r in rows[hi + 1:]:
        if len(r) > vi:
            us = float(r[vi].replace(",", "")) / 1000.0
            tot += us
            fh.write(f"| `{r[ki][:70]}` | {us:.1f} |\n")
    fh.write(f"| total | {tot:.1f} |\n")
for rep, kern, name, what in ((src / "prof_rows.ncu-rep", "rowScanKernel", f"{tag}_rows_summary.md",
                               "window-row scan kernel, ONE launch, 0.25-scale assembly (714 Mbp, 714 K window rows), packed records"),
                              (src / "prof_reads.ncu-rep", "readScanKernel", f"{tag}_reads_summary.md",
                               "read scorer, 2.0 Gbp of synthetic HiFi reads (133 K reads)")):
    if not rep.exists():
        continue
    m = run("python", "scripts/ncu_metrics.py", str(rep))
    l = run("python", "scripts/ncu_lines.py", str(rep), kern, "--top", "40")
    (out / name).write_text(f"# {tag} — {what}: ncu --set full --clock-control none --import-source on\n\n{m}\nSource-line attribution "
                            f"(`scripts/ncu_lines.py`, share of executed warp instructions):\n\n```\n{l}```\n")
# DRAM bytes of the window-row kernel's launch -> what bench.py scales into roofline.traffic
rep, log = src / "prof_rows.ncu-rep", src / "ncu_full.log"
if rep.exists():
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = {}
    for ln in run("python", "scripts/ncu_metrics.py", str(rep)).splitlines():
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if ln.startswith(f"- {key} = "):
                v, u = ln.split(" = ")[1].split()[:2]
                vals[key] = float(v.replace(",", "")) * unit[u]
    bases = [int(ln.split()[1]) for ln in log.read_text().splitlines() if ln.startswith("bases ")] if log.exists() else []
    if len(vals) == 2 and bases:
        rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
        (out / "rows_kernel_traffic.json").write_text(json.dumps({
            "dram_bytes_per_base": (rd + wr) / bases[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "bases": bases[0],
            "source": f"ncu --set full --clock-control none, TLS_PACKED=1 TLS_ROW_PARTS=1 scripts/trace_bench.py 0.25 (profiles/{tag}_rows_summary.md); "
                      "dram__bytes_read.sum + dram__bytes_write.sum of one launch"}, indent=1) + "\n")
# the kernels behind the scan: headline metrics of each
rep = src / "prof_blocks.ncu-rep"
if rep.exists():
    keys = ("Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    with open(out / f"{tag}_block_kernels_summary.md", "w") as fh:
        fh.write(f"# {tag} — the kernels behind the scan on the 3.03 Gbp batch (packed records): ncu --set full --clock-control none, one launch each\n")
        for kern in ("gatherKernel", "gatherDenseKernel", "terminalKernel", "itsConnKernel", "itsWalkKernel"):
            m = run("python", "scripts/ncu_metrics.py", str(rep), kern)
            fh.write(f"\n## {kern}\n\n")
            for ln in m.splitlines():
                if any(ln.startswith(f"- {k} ") for k in keys):
                    fh.write(ln + "\n")
            l = run("python", "scripts/ncu_lines.py", str(rep), kern, "--top", "8")
            fh.write(f"\nhottest source lines (share of executed warp instructions, share of stall samples):\n\n```\n{l}```\n")
# timeline of one step (TLS_TRACE=1) and the file pipeline's stage clock
tl = src / "trace_tls.log"
if tl.exists():
    keep = [ln.strip()[:400] for ln in tl.read_text().splitlines() if "timeline" in ln or ln.startswith("[tls] submit:") or ln.startswith("iter ")]
    (out / f"{tag}_step_timeline.md").write_text(f"# {tag} — one step on the device and on the host (`TLS_TRACE=1 TLS_PACKED=1 python scripts/trace_bench.py 1.0`, "
                                                "four steps; CUDA events and steady_clock marks of csrc/api.cu; the fprintf calls of the trace inflate the host phases)\n\n```\n"
                                                + "\n".join(keep) + "\n```\n")
pl = src / "pipeline_n1.log"
if pl.exists():
    (out / f"{tag}_pipeline_stages.md").write_text(f"# {tag} — teloscope-b200, 3.03 Gbp FASTA -> 11 files on one device: stage clock (`TLS_PIPE_TRACE=1`, "
                                                  "`scripts/pipeline_bench.py 1.0 --repeat 3`)\n\n```\n" + pl.read_text() + "```\n")
print("written to", out)
