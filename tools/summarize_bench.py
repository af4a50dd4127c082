#!/usr/bin/env python3
"""Print a compact summary of bench.py JSON lines (gpurun_out/bench_*.json)."""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as e:
        print(path, "ERR", e); continue
    print(path, "ms/step %.3f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "launches", d["gpu_launches"])
    print("  roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k in ("achieved", "peak", "frac", "ms", "flops_alg")})
    print("  clocks", d["clocks"])
    if d.get("cpu_baseline"): print("  cpu %.3e q/s cores %d" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
    rows = sorted(d.get("per_class", []), key=lambda r: -r["ms"])
    print("  per-class total ms %.2f" % sum(r["ms"] for r in rows))
    for r in rows[:int(__import__('os').environ.get("TOP", "60"))]:
        print("   %s|%s q=%d ms=%.3f tf=%.3f frac=%.3f G=%d" % (r["bra"], r["ket"], r["quartets"], r["ms"], r["tflops"], r["frac"], r["group"]))
