#!/usr/bin/env python3
"""Regenerate tests/golden/hess_golden.npz with the CPU oracle's getRepulsion2 restatement (oracle/oracle.c,
ref_getRepulsion2 = src/Integral/Int4C2E.cpp:410-492 of the reference).

    python tests/golden/make_hess_golden.py [--check]

Per molecule (h2o / cc-pVDZ, hf_tz = HF / cc-pVTZ with f shells, bo3h3 / 6-31G**):
    <name>_hess      ContractHesss(., n * D31), EXX = 0.6, [3*natom][3*natom], D31 = scf_harness.random_symmetric_density(n, 31)
    <name>_hess_j    the same with EXX = 0 (Coulomb part only; the reference skips hessiank, :475)
The reference holds no golden vectors for this path; these are the oracle's outputs, pinned by finite differences of the
first-derivative route in tests/test_oracle.py (which Chinium's recorded forces pin, tools/sn2/sn2.cnm.log:211-216).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from chinium_b200.inputs import load_fixture_molecule  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
import scf_harness as H  # noqa: E402

NAMES = ("h2o", "hf_tz", "bo3h3")


def build(names=NAMES):
    o = Oracle()
    o.set_threads(0)
    out = {}
    for name in names:
        mol, fb = load_fixture_molecule(name)
        n = fb.nbf
        D = H.random_symmetric_density(n, 31) * n
        out[name + "_hess"] = o.contract_hess(fb, D, 0.6)
        out[name + "_hess_j"] = o.contract_hess(fb, D, 0.0)
    return out


def main():
    path = os.path.join(HERE, "hess_golden.npz")
    if "--check" in sys.argv:
        old = np.load(path)
        new = build()
        worst = max(float(np.abs(old[k] - new[k]).max()) for k in old.files)
        print("max deviation from the committed fixture: %.3e over %d arrays" % (worst, len(old.files)))
        return 0 if worst < 1e-10 else 1
    new = build()
    np.savez_compressed(path, **new)
    print("wrote", path, sorted(new))
    return 0


if __name__ == "__main__":
    sys.exit(main())
