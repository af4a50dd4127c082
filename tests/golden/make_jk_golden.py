#!/usr/bin/env python3
"""Regenerate tests/golden/jk_golden.npz with the CPU oracle (oracle/oracle.c through tests/oracle_lib.py).

    python tests/golden/make_jk_golden.py [--check]

Per molecule (h2o / cc-pVDZ = examples/h2o.inp, hf_tz = HF / cc-pVTZ with f shells), seeded symmetric densities
(tests/scf_harness.random_symmetric_density):
    <name>_J, <name>_K          ContractInts(D0, 0x0, 0x0), EXX = 1           (the literal stored-integral path)
    <name>_J_ab, _Ka, _Kb       ContractInts(0x0, D1, D2),  EXX = 1
    <name>_counts               (RepulsionLength, ShellQuartetLength) of the reference's screening printout
    <name>_grad                 ContractGrads(n*D21, n*D22), EXX = 0.7         (getRepulsion1 restatement), index 3*atom+xyz
The reference holds no golden vectors for this path, so these are the oracle's own outputs, pinned by the
checks in tests/test_oracle.py; the GPU tests compare the engine with them without needing the oracle at run time.
--check: recompute and compare with the committed file instead of writing it.
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


def build():
    o = Oracle()
    out = {}
    for name in ("h2o", "hf_tz"):
        mol, fb = load_fixture_molecule(name)
        n = fb.nbf
        D, Da, Db = (H.random_symmetric_density(n, s) for s in (0, 1, 2))
        J, K, _, _, cnt = o.reference_jk(fb, D)
        J2, _, Ka, Kb, _ = o.reference_jk(fb, None, Da, Db)
        out[name + "_J"], out[name + "_K"] = J, K
        out[name + "_J_ab"], out[name + "_Ka"], out[name + "_Kb"] = J2, Ka, Kb
        out[name + "_counts"] = np.array(cnt, dtype=np.int64)
        out[name + "_grad"] = o.contract_grads(fb, H.random_symmetric_density(n, 21) * n, H.random_symmetric_density(n, 22) * n, 0.7)
    return out


def main():
    path = os.path.join(HERE, "jk_golden.npz")
    new = build()
    if "--check" in sys.argv:
        old = np.load(path)
        worst = 0.0
        for k in old.files:
            worst = max(worst, float(np.abs(np.asarray(old[k], dtype=np.float64) - np.asarray(new[k], dtype=np.float64)).max()))
        print("max deviation from the committed fixture: %.3e over %d arrays" % (worst, len(old.files)))
        return 0 if worst < 1e-12 else 1
    np.savez_compressed(path, **new)
    print("wrote", path, sorted(new))
    return 0


if __name__ == "__main__":
    sys.exit(main())
