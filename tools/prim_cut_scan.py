#!/usr/bin/env python3
"""Accuracy / time scan of the primitive-quartet cutoff of the J/K kernels (developer knob CF_PRIM_CUT, engine.cu).

    python tools/prim_cut_scan.py --workloads c18 fe4s4 h2o64 --cuts 1e-22 1e-20 1e-18 1e-16 --out gpurun_out/primcut.json

Every (workload, cutoff) runs in its own process (the knob is read once).  Densities with O(1) entries: the seeded
stress density of SURVEY 8d scaled by nbf (random signs) and its element-wise absolute value (all positive: errors of
skipped positive integrals add up coherently -- the worst case).  Reports max|dJ|, max|dK| against the first cutoff of the list
and the CUDA-event time of the ERI kernels (median of 3 builds).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(workload, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import bench
    from chinium_b200 import Int4C2E
    from chinium_b200.inputs import load_fixture_molecule
    import scf_harness as H
    fixture, kind = bench.WORKLOADS[workload]
    mol, fb = load_fixture_molecule(fixture)
    thr = bench.DEFAULT_THRESHOLD.get(workload, -1.0)
    eng = Int4C2E(fb, 1.0, thr, device=0)
    n = fb.nbf
    D0 = H.random_symmetric_density(n, 0) * n
    res = {}
    for tag, D in (("signed", D0), ("positive", np.abs(D0))):
        ms = []
        for _ in range(3):
            J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
            ms.append(eng.stats["ms_eri_last"])
        res[tag] = {"ms": float(np.median(ms)), "prim": int(eng.stats["primitive_quartets_executed_last"])}
        np.savez(out + "_" + tag + ".npz", J=J, K=K)
    json.dump(res, open(out + ".json", "w"))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", nargs="+", default=["c18"])
    ap.add_argument("--cuts", nargs="+", default=["1e-22", "1e-20", "1e-18", "1e-16"])
    ap.add_argument("--out", default="gpurun_out/primcut.json")
    ap.add_argument("--child", nargs=2)
    a = ap.parse_args()
    if a.child:
        return child(*a.child)
    import numpy as np
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    table = []
    for w in a.workloads:
        base = {}
        for c in a.cuts:
            stem = "/tmp/primcut_%s_%s" % (w, c)
            env = dict(os.environ, CF_PRIM_CUT=c)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", w, stem], env=env, capture_output=True, text=True, timeout=1500)
            if r.returncode != 0:
                print(w, c, "FAILED", r.stderr[-500:])
                continue
            info = json.load(open(stem + ".json"))
            row = {"workload": w, "prim_cut": float(c)}
            for tag in ("signed", "positive"):
                z = np.load(stem + "_" + tag + ".npz")
                if tag not in base:
                    base[tag] = (z["J"], z["K"])
                row[tag] = {"ms_eri": info[tag]["ms"], "primitive_quartets_executed": info[tag]["prim"],
                            "max_dJ": float(np.abs(z["J"] - base[tag][0]).max()), "max_dK": float(np.abs(z["K"] - base[tag][1]).max()),
                            "max_J": float(np.abs(z["J"]).max())}
            table.append(row)
            print("%-6s cut %-7s  signed: %8.2f ms prim %.4e dJ %.2e dK %.2e | positive: %8.2f ms dJ %.2e dK %.2e (|J|max %.1e)" % (
                w, c, row["signed"]["ms_eri"], row["signed"]["primitive_quartets_executed"], row["signed"]["max_dJ"], row["signed"]["max_dK"],
                row["positive"]["ms_eri"], row["positive"]["max_dJ"], row["positive"]["max_dK"], row["positive"]["max_J"]), flush=True)
    json.dump(table, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
