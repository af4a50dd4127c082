#!/usr/bin/env python3
"""Compare per-class tables of bench.py JSON files: cmp_bench.py a.json b.json [c.json ...] [--top N] [--fam]"""
import json, sys
files = [a for a in sys.argv[1:] if not a.startswith("--")]
top = 60
for a in sys.argv[1:]:
    if a.startswith("--top="): top = int(a[6:])
L = []
for f in files:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    L.append((d, {(r["bra"], r["ket"]): r for r in d.get("per_class", [])}))
print("step ms:", " ".join("%.2f" % d["ms_per_step"] for d, _ in L), " per-class sum:", " ".join("%.2f" % sum(r["ms"] for r in t.values()) for _, t in L))
keys = sorted(L[0][1], key=lambda k: -L[0][1][k]["ms"])[:top]
for k in keys:
    print("  %s|%s  " % k + "  ".join("%7.3f" % (t[k]["ms"] if k in t else float("nan")) for _, t in L) + "   frac " + " ".join("%.3f" % (t[k]["frac"] if k in t else 0) for _, t in L))
