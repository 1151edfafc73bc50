"""Run every golden FASTA case through the product and the oracle and print the first differences."""
import gzip, json, sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_lib as ol, product_adapter as pa
import teloscope_b200 as tb

cases = json.load(gzip.open("tests/golden/ref_cases.json.gz"))["cases"]
fixtures = np.load("tests/golden/fixtures.npz")
bad = 0
for c in cases:
    if c["kind"] != "fasta": continue
    cfg, flags = ol.config_from_args(c["args"])
    if cfg.step != cfg.window_size: continue
    paths = ol.parse_fasta(fixtures[c["fixture"]].tobytes())
    tips = not flags["genome"]
    sc = tb.Scanner(pa.scan_config(cfg, flags))
    prod = pa.BatchedProduct(sc, cfg, paths, tips)
    orc = ol.Oracle(cfg)
    i = 0
    msgs = []
    for _, comps, _ in paths:
        for kind, a, payload in comps:
            if kind != "seg": continue
            want = orc.scan_segment(payload.upper(), a, tips)
            got = prod.item_as_oracle(i, a)
            if (prod.res.n_fwd[i], prod.res.n_rev[i]) != (len(want["fwd"]), len(want["rev"])):
                msgs.append(f"item {i} len={len(payload)} n_fwd/rev got {prod.res.n_fwd[i]},{prod.res.n_rev[i]} want {len(want['fwd'])},{len(want['rev'])}")
            for k in ("terminal", "its"):
                g, w = got[k], want[k]
                if len(g) != len(w) or any(not np.array_equal(g[n], w[n]) for n in ("start", "len", "counts", "fwd_count", "can_count", "label", "valid_or")):
                    msgs.append(f"item {i} {k}: got {[(int(b['start']), int(b['len']), int(b['counts'])) for b in g]} want {[(int(b['start']), int(b['len']), int(b['counts'])) for b in w]}")
            if not tips:
                for k in ("canonical", "noncanonical"):
                    g, w = got[k], want[k]
                    if len(g) != len(w) or not np.array_equal(g["pos"], w["pos"]):
                        msgs.append(f"item {i} {k}: got n={len(g)} want n={len(w)} first got {g['pos'][:5]} want {w['pos'][:5]}")
                g, w = got["windows"], want["windows"]
                for k in ("can_cov", "noncan_cov", "fwd_cov", "rev_cov", "acgt"):
                    if cfg.count_acgt or k != "acgt":
                        if len(g) != len(w) or not np.array_equal(g[k], w[k]):
                            d = np.flatnonzero(np.any(np.atleast_2d(g[k] != w[k]).reshape(len(g), -1), axis=1))[:3] if len(g) == len(w) else []
                            msgs.append(f"item {i} windows.{k} differ at {list(d)}: got {[g[k][j].tolist() for j in d]} want {[w[k][j].tolist() for j in d]}")
            i += 1
    sc.close()
    if msgs:
        bad += 1
        print("CASE", c["fixture"], " ".join(c["args"]), "tips" if tips else "genome", f"matches={prod.res.n_matches}")
        for m in msgs[:6]: print("   ", m)
        if bad >= 8: break
print("bad cases:", bad)
