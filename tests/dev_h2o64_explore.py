#!/usr/bin/env python3
"""(H2O)64 / def2-TZVP: Schwarz threshold vs accuracy and time (exploration; run on the GPU box).
For each threshold: setup time, surviving quartets, build time, max error of sampled J/K blocks against the
unscreened CPU oracle, with a density of O(1) entries (the stress density times nbf) -- worst case for screening."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from chinium_b200 import Int4C2E
from chinium_b200.inputs import load_fixture_molecule
from oracle_lib import Oracle
import scf_harness as H

mol, fb = load_fixture_molecule("h2o64")
n = fb.nbf
D = H.random_symmetric_density(n, 0) * n
o = Oracle()
per = fb.nshell // 64      # shells per water
pairs = [(0, 0), (4, 7 * per + 2), (per * 20 + 9, per * 20 + 3), (per * 63 + 8, 5), (per * 33 + 12, per * 32 + 1)]
t = time.time()
ref = [o.jk_block(fb, 2 * D, D, sa, sb) for sa, sb in pairs]
print("oracle blocks %.1f s" % (time.time() - t), flush=True)
for thr in [float(x) for x in (sys.argv[1:] or ["1e-11", "1e-13"])]:
    t = time.time()
    eng = Int4C2E(fb, 1.0, thr, device=0)
    eng._ensure()
    ts = time.time() - t
    st = eng.stats
    t = time.time()
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    tb = time.time() - t
    t = time.time()
    J, K, _, _ = eng.ContractInts(D, None, None, 1, 0)
    tb2 = time.time() - t
    eJ = eK = 0.0
    for (sa, sb), (Jb, Kb) in zip(pairs, ref):
        ia, ib = fb.shell2bf[sa], fb.shell2bf[sb]
        eJ = max(eJ, np.abs(Jb - J[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max())
        eK = max(eK, np.abs(Kb - K[ia:ia + fb.nfun[sa], ib:ib + fb.nfun[sb]]).max())
    print("thr %g: setup %.1f s, quartets %.4g, build %.2f / %.2f s, device ms %.1f, max|dJ| %.2e max|dK| %.2e (|D|~1; /%d for the stress density), |J|max %.2e"
          % (thr, ts, st["canonical_quartets"], tb, tb2, eng.stats["ms_device_last"], eJ, eK, n, np.abs(J).max()), flush=True)
    eng.close()
