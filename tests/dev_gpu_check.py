#!/usr/bin/env python3
"""Developer check run on the GPU box: engine vs oracle on small molecules, with per-class error maps."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from chinium_b200.inputs import load_fixture_molecule
from chinium_b200 import Int4C2E
from chinium_b200.fock import measure_fp64_peak, device_info
from oracle_lib import Oracle
import scf_harness as H

def block_err(fb, A, B):
    out = {}
    for sa in range(fb.nshell):
        for sb in range(fb.nshell):
            la, lb = abs(int(fb.type[sa])), abs(int(fb.type[sb]))
            ia, ib = fb.shell2bf[sa], fb.shell2bf[sb]
            e = np.abs(A[ia:ia+fb.nfun[sa], ib:ib+fb.nfun[sb]] - B[ia:ia+fb.nfun[sa], ib:ib+fb.nfun[sb]]).max()
            k = (max(la, lb), min(la, lb))
            out[k] = max(out.get(k, 0), e)
    return out

def main():
    names = sys.argv[1:] or ["h2o", "hf_tz", "bo3h3"]
    print(device_info(), "fp64 peak TF/s", measure_fp64_peak())
    o = Oracle()
    ok = True
    for name in names:
        mol, fb = load_fixture_molecule(name)
        n = fb.nbf
        D = H.random_symmetric_density(n, 0)
        Da = H.random_symmetric_density(n, 1); Db = H.random_symmetric_density(n, 2)
        t = time.time(); eng = Int4C2E(fb, 1.0, -1.0); eng.getRepulsionDiag(0); t_setup = time.time() - t
        t = time.time(); J, K, _, _ = eng.ContractInts(D, None, None, 1, 0); t_b = time.time() - t
        st = eng.stats
        t = time.time(); Jo, Ko, _, _, cnt = o.direct_jk(fb, D); t_o = time.time() - t
        eJ, eK = np.abs(J - Jo).max(), np.abs(K - Ko).max()
        print(f"{name}: nbf {n} quartets {st['canonical_quartets']} (oracle {cnt[0]}) uniq {st['unique_integrals']} primq {st['primitive_quartets']} (oracle {cnt[1]})"
              f" F_alg {st['flops_alg_jk'][1]:.3e} setup {t_setup:.2f}s build {t_b*1e3:.1f} ms (dev {st['ms_device_last']:.2f} eri {st['ms_eri_last']:.2f}) oracle {t_o:.1f}s scale2^{st['fixedpoint_scale_log2']}")
        print(f"   RHF-type  max|dJ| {eJ:.2e} max|dK| {eK:.2e}   |J| {np.abs(Jo).max():.2e}")
        if eJ > 1e-10 or eK > 1e-10:
            ok = False
            print("   J block errors", {k: float('%.1e' % v) for k, v in sorted(block_err(fb, J, Jo).items())})
            print("   K block errors", {k: float('%.1e' % v) for k, v in sorted(block_err(fb, K, Ko).items())})
        J2, _, Ka, Kb = eng.ContractInts(None, Da, Db, 1, 0)
        Jo2, _, Kao, Kbo, _ = o.direct_jk(fb, None, Da, Db)
        e2 = max(np.abs(J2 - Jo2).max(), np.abs(Ka - Kao).max(), np.abs(Kb - Kbo).max())
        print(f"   UHF-type  max err {e2:.2e}")
        ok = ok and e2 < 1e-10
        dg = o.repulsion_diag(fb)
        ed = np.abs(eng.RepulsionDiags[0] - dg).max()
        print(f"   diag1212 max err {ed:.2e}")
        ok = ok and ed < 1e-10
        # bit stability: two builds identical
        J3, K3, _, _ = eng.ContractInts(D, None, None, 1, 0)
        print("   bitwise repeatable:", bool((J3 == J).all() and (K3 == K).all()))
        eng.close()
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1

if __name__ == "__main__":
    sys.exit(main())
