#!/usr/bin/env python3
"""Per-kernel segment table from an ncu report's SASS source page: contiguous SASS ranges with similar execution counts,
with stall-sample share, shared-memory wavefronts (and ideal), global tag requests, L2 sectors, FP64 instruction share.
usage: ncu_segments.py <report.ncu-rep> [kernel-substring] [min-share%]"""
import csv, subprocess, io, sys
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
seen = set()
for b in txt.split('"Kernel Name",')[1:]:
    lines = b.splitlines(); name = lines[0].strip('",')
    if name in seen or sub not in name: continue
    seen.add(name)
    rd = list(csv.reader(io.StringIO("\n".join(lines[1:])))); hdr = rd[0]; ix = {h: i for i, h in enumerate(hdr)}
    rows = [r for r in rd[1:] if len(r) == len(hdr)]
    def f(r, h):
        try: return float(r[ix[h]])
        except Exception: return 0.0
    tot = sum(f(r, "# Samples") for r in rows)
    T = {k: sum(f(r, k) for r in rows) for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global", "L2 Theoretical Sectors Global", "Instructions Executed")}
    print("=====", name)
    print("  totals: samples %d, warp-instr %.3g, shared wavefronts %.3g (ideal %.3g), global tag requests %.3g, L2 sectors %.3g" % (
        tot, T["Instructions Executed"], T["L1 Wavefronts Shared"], T["L1 Wavefronts Shared Ideal"], T["L1 Tag Requests Global"], T["L2 Theoretical Sectors Global"]))
    segs = []; cur = None
    for i, r in enumerate(rows):
        e = max(f(r, "Instructions Executed"), 1.0)
        if cur is None or not (cur["e"] / 1.3 <= e <= cur["e"] * 1.3):
            cur = dict(a=i, b=i, e=e, s=0, n=0, fp=0, wf=0, wfi=0, g=0, l2=0, st={}); segs.append(cur)
        src = r[ix["Source"]]; toks = src.split(); op = toks[1] if src.strip().startswith("@") else toks[0]
        cur["b"] = i; cur["s"] += f(r, "# Samples"); cur["n"] += 1; cur["fp"] += op.startswith(("DFMA", "DMUL", "DADD"))
        cur["wf"] += f(r, "L1 Wavefronts Shared"); cur["wfi"] += f(r, "L1 Wavefronts Shared Ideal"); cur["g"] += f(r, "L1 Tag Requests Global"); cur["l2"] += f(r, "L2 Theoretical Sectors Global")
        for h in hdr:
            if h.startswith("stall_") and "Not" not in h: cur["st"][h[6:]] = cur["st"].get(h[6:], 0) + f(r, h)
    for g in segs:
        if g["s"] > tot * minshare / 100:
            st = sorted(g["st"].items(), key=lambda kv: -kv[1])[:3]
            print("  sass %4d-%4d n=%4d exec %9d fp64 %3d | samples %5.1f%% | shared wf %5.1f%% (x%.2f ideal) | global req %5.1f%% | %s" % (
                g["a"], g["b"], g["n"], g["e"], g["fp"], 100 * g["s"] / tot, 100 * g["wf"] / max(T["L1 Wavefronts Shared"], 1), g["wf"] / max(g["wfi"], 1),
                100 * g["g"] / max(T["L1 Tag Requests Global"], 1), " ".join("%s:%.1f" % (k, 100 * v / tot) for k, v in st)))
