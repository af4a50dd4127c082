#!/usr/bin/env python3
"""Print the headline numbers and the heaviest class-pair kernels of a bench.py JSON line.  usage: show_bench.py file.json [nrows] [other.json]"""
import json, sys
d = json.load(open(sys.argv[1]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
other = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else None
r = d["roofline"]
print("%s: %.2f ms/build (e2e %.2f), eri %.2f ms, frac_executed %.3f, frac_nominal %.3f, launches %s, setup %.2f s" % (
    d["config"]["workload"].split()[0], d["ms_per_step"], d["e2e"]["ms_per_step"], r["ms"], r.get("frac_executed", float("nan")),
    r.get("frac_nominal", r.get("frac", float("nan"))), d.get("gpu_launches"), d["config"].get("setup_s", float("nan"))))
if "per_class" in d:
    rows = sorted(d["per_class"], key=lambda x: -x["ms"])
    om = {(x["bra"], x["ket"]): x["ms"] for x in other["per_class"]} if other and "per_class" in other else {}
    print("  sum of class-pair kernels %.2f ms%s" % (sum(x["ms"] for x in rows), (" (other: %.2f)" % sum(om.values())) if om else ""))
    for x in rows[:n]:
        extra = ("  other %.3f (%+.0f%%)" % (om[(x["bra"], x["ket"])], 100 * (x["ms"] / om[(x["bra"], x["ket"])] - 1))) if (x["bra"], x["ket"]) in om else ""
        print("   %s|%s %.3f ms frac %.3f%s" % (x["bra"], x["ket"], x["ms"], x.get("frac", float("nan")), extra))
