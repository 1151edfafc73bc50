#!/usr/bin/env python3
"""Attribute an ncu report's per-SASS-instruction counters to CUDA source lines.

    python scripts/ncu_lines.py <report.ncu-rep> <kernel-regex> [--top N] [--cubin path] [--range A-B]

ncu's `--page source --csv` lists SASS only; the line table comes from `nvdisasm -g` on the
cubin extracted from libteloscope_b200.so, matched to the report by instruction order.
Prints: executed warp instructions and stall samples per source line (inlined callees are
attributed to the innermost line nvdisasm reports), plus an opcode histogram.
"""
import argparse
import csv
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def sass_rows(rep: str, kernel: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[0][1]
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break          # only the first matching launch
            continue
        res.append({"sass": r[col["Source"]].strip(), "inst": float(r[col["Instructions Executed"]] or 0),
                    "thr": float(r[col["Thread Instructions Executed"]] or 0),
                    "samples": float(r[col["# Samples"]] or 0),
                    "conf": float((r[col["L1 Wavefronts Shared Excessive"]] if "L1 Wavefronts Shared Excessive" in col else 0) or 0)})
    return name, res


def line_table(cubin: str, mangled_hint: str):
    txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True, check=True).stdout
    fn, cur, table = None, None, defaultdict(list)
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
            table[fn].append(cur)
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--cubin")
    ap.add_argument("--lib", default=str(ROOT / "teloscope_b200" / "libteloscope_b200.so"))
    ap.add_argument("--by-line", action="store_true", help="list every line of the kernel's own file in source order")
    ap.add_argument("--per", type=float, default=0, help="divide executed instructions by this (e.g. rows processed)")
    ap.add_argument("--sass", action="store_true", help="dump the SASS with executed counts and line numbers")
    args = ap.parse_args()
    name, rows = sass_rows(args.report, args.kernel)
    cubins = [args.cubin] if args.cubin else []
    if not cubins:
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", args.lib], cwd=tmp, check=True, capture_output=True)
        cubins = [str(p) for p in Path(tmp).glob("*.cubin")]
    lines = None
    for cb in cubins:
        for fn, tab in line_table(cb, args.kernel).items():
            if len(tab) == len(rows) and re.search(args.kernel, fn):
                lines = tab
                break
        if lines:
            break
    if lines is None:
        sys.exit(f"no function with {len(rows)} instructions matching {args.kernel} in {cubins}")
    tot_i = sum(r["inst"] for r in rows)
    tot_s = sum(r["samples"] for r in rows) or 1
    per = defaultdict(lambda: [0.0, 0.0, 0.0])
    ops = Counter()
    for r, l in zip(rows, lines):
        per[l][0] += r["inst"]; per[l][1] += r["samples"]; per[l][2] += r["conf"]
        op = r["sass"].split()[0]
        if op.startswith("@"):
            op = r["sass"].split()[1]
        ops[op.split(".")[0]] += r["inst"]
    print(f"kernel: {name}\nSASS instructions: {len(rows)}  executed warp-inst: {tot_i:.0f}  samples: {tot_s:.0f}")
    src_cache = {}
    def src(l):
        if l is None:
            return ""
        f = next(iter(ROOT.rglob(l[0])), None)
        if f is None:
            return ""
        if f not in src_cache:
            src_cache[f] = f.read_text().splitlines()
        return src_cache[f][l[1] - 1].strip()[:100] if l[1] - 1 < len(src_cache[f]) else ""
    if args.sass:
        for r, l in zip(rows, lines):
            n = r["inst"] / args.per if args.per else r["inst"]
            print(f"{n:10.2f} {r['samples']:6.0f}  {(l[0] + ':' + str(l[1])) if l else '?':28s} {r['sass']}")
        return
    if args.by_line:
        print(f"{'inst':>8} {'samp%':>6}  line")
        for l, (i, sm, c) in sorted(per.items(), key=lambda kv: (kv[0][0] if kv[0] else '', kv[0][1] if kv[0] else 0)):
            n = i / args.per if args.per else i
            print(f"{n:8.2f} {100*sm/tot_s:6.2f}  {l[0] if l else '?'}:{l[1] if l else 0}  {src(l)}")
        return
    print(f"{'inst%':>6} {'samp%':>6} {'smem-excess':>11}  line")
    for l, (i, s, c) in sorted(per.items(), key=lambda kv: -kv[1][0])[:args.top]:
        print(f"{100*i/tot_i:6.2f} {100*s/tot_s:6.2f} {c:11.0f}  {l[0] if l else '?'}:{l[1] if l else 0}  {src(l)}")
    print("\nopcode histogram (executed warp-inst %):")
    print("  ".join(f"{k} {100*v/tot_i:.1f}" for k, v in ops.most_common(28)))


if __name__ == "__main__":
    main()
