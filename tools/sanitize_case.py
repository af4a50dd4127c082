#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): J/K (RHF + UHF), multi-density, gradient, Hessian on
molecules with s..f shells.  usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py [names...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from chinium_b200 import Int4C2E
from chinium_b200.inputs import load_fixture_molecule
import scf_harness as H

mode = os.environ.get("SAN_MODE", "all")
for name in (sys.argv[1:] or ["hf_tz"]):
    mol, fb = load_fixture_molecule(name)
    n = fb.nbf
    eng = Int4C2E(fb, 1.0, -1.0, device=0)
    D, Da, Db = (H.random_symmetric_density(n, s) * n for s in (0, 1, 2))
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    J2, _, Ka, Kb = eng.ContractInts(None, Da, Db, 1, 0)
    print(name, "J/K done", float(np.abs(J).max()), float(np.abs(Ka).max()), flush=True)
    if mode == "all":
        G = eng.ContractInts([D, Da, Db], 1, 0)
        g = eng.ContractGrads(D, D, 0)
        Hm = eng.ContractHesss(D, D, 0)
        print(name, "multi/grad/hess done", float(np.abs(G[0]).max()), float(np.abs(g).max()), float(np.abs(Hm).max()), flush=True)
    eng.close()
