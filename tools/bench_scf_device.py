#!/usr/bin/env python3
"""ms per SCF iteration with everything resident on the device: J/K build vs SCF linear algebra.
usage: bench_scf_device.py [workload] [iterations]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from chinium_b200 import Int4C2E
from chinium_b200.inputs import load_fixture_molecule
import scf_harness as H
import scf_device

w = sys.argv[1] if len(sys.argv) > 1 else "h2o64"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mol, fb = load_fixture_molecule(w)
thr = 1e-13 if w == "h2o64" else -1.0
t0 = time.perf_counter()
eng = Int4C2E(fb, 1.0, thr)
eng._ensure()
setup = time.perf_counter() - t0
tm = {}
try:
    E, D, it = scf_device.rhf_device(eng, mol.Z, mol.xyz_bohr, mol.nelec // 2, H.nuclear_repulsion(mol.Z, mol.xyz_bohr), max_iter=iters, timings=tm)
    conv = True
except RuntimeError:
    conv = False
print(json.dumps({"workload": w, "nbf": fb.nbf, "setup_s": setup, "iterations_timed": len(tm["jk"]), "converged_within": conv,
                  "jk_ms_per_iter": float(np.median(tm["jk"])), "linalg_ms_per_iter": float(np.median(tm["linalg"])),
                  "jk_ms": tm["jk"], "linalg_ms": tm["linalg"]}))
