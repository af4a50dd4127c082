#!/usr/bin/env python3
"""Time cf_contract_hess (ContractHesss) on fixtures: usage time_hess.py name [name ...]; prints one JSON line per molecule."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from chinium_b200 import Int4C2E
from chinium_b200.inputs import load_fixture_molecule
import scf_harness as H

for name in sys.argv[1:]:
    mol, fb = load_fixture_molecule(name)
    eng = Int4C2E(fb, 1.0, -1.0, device=0)
    D = H.random_symmetric_density(fb.nbf, 31) * fb.nbf
    eng.ContractHesss(D, D, 0)
    t = time.perf_counter()
    Hm = eng.ContractHesss(D, D, 0)
    wall = time.perf_counter() - t
    st = eng.stats
    natom = Hm.shape[0] // 3
    print(json.dumps({"molecule": name, "nbf": fb.nbf, "natom": natom, "canonical_quartets": int(st["canonical_quartets"]),
                      "hess_kernels_ms": st["ms_grad_last"], "host_call_ms": wall * 1e3, "launches": st["n_launches_last"],
                      "max_abs": float(np.abs(Hm).max()), "asym": float(np.abs(Hm - Hm.T).max()),
                      "translation_residual": float(np.abs(Hm.reshape(3 * natom, natom, 3).sum(axis=1)).max())}), flush=True)
    eng.close()
