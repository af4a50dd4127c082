"""Device-resident RHF loop around the engine (TEST / BENCH HARNESS).

D, J, K, F, S, H never leave HBM: the one-electron matrices come from cf_one_electron_device, J/K from
cf_build_jk_device, the SCF linear algebra (S^-1/2, F' = X^T F X, eigh, FDS - SDF, Pulay CDIIS) runs through torch on
the same device (cuSOLVER syevd / cuBLAS GEMMs: plain library calls, like the Eigen calls of Restricted/SP.cpp:40-73).
Only scalars (energy, max|commutator|, the small DIIS matrix) cross the PCIe link.
Restates the same rules as tests/scf_harness.rhf: D = C_occ C_occ^T, F = H + J - K, E = sum D o (2H + J - K),
commutator 2 (F D S - S D F), CDIIS (src/DIIS/CDIIS.cpp:24-47).
"""
import time

import numpy as np
import torch


def rhf_device(eng, Z, xyz_bohr, nocc, e_nuc=0.0, max_iter=100, tol=1e-8, diis_space=12, device=None, timings=None):
    """eng: chinium_b200.Int4C2E (single device).  -> E, D (torch, device), iterations.
    `timings`: optional dict receiving per-iteration milliseconds {'jk': [...], 'linalg': [...]} (CUDA events)."""
    eng._ensure()
    n = eng.nbf
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    f64 = dict(dtype=torch.float64, device=dev)
    S, T, V = (torch.empty((n, n), **f64) for _ in range(3))
    stream = torch.cuda.current_stream(dev).cuda_stream
    eng.one_electron_device(Z, xyz_bohr, S.data_ptr(), T.data_ptr(), V.data_ptr(), stream)
    H = T + V
    w, U = torch.linalg.eigh(S)
    X = (U * w.rsqrt()) @ U.T

    def density(F):
        e, Cp = torch.linalg.eigh(X.T @ F @ X)
        C = X @ Cp[:, :nocc]
        return C @ C.T

    D = density(H)
    J, K = torch.empty((n, n), **f64), torch.empty((n, n), **f64)
    Fs, Rs = [], []
    E = 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(max_iter):
        ev[0].record()
        eng.build_jk_device(D.data_ptr(), None, None, J.data_ptr(), K.data_ptr(), None, None, stream)     # symmetric: layout-free
        ev[1].record()
        F = H + J - K
        E = float((D * (2 * H + J - K)).sum()) + e_nuc
        R = 2 * (F @ D @ S - S @ D @ F)
        err = float(R.abs().max())
        if err < tol:
            ev[2].record(); torch.cuda.synchronize()
            if timings is not None:
                timings.setdefault("jk", []).append(ev[0].elapsed_time(ev[1])); timings.setdefault("linalg", []).append(ev[1].elapsed_time(ev[2]))
            return E, D, it
        Fs.append(F); Rs.append(X.T @ R @ X)
        Fs, Rs = Fs[-diis_space:], Rs[-diis_space:]
        m = len(Fs)
        B = -np.ones((m + 1, m + 1)); B[m, m] = 0
        Rm = torch.stack([r.reshape(-1) for r in Rs])
        B[:m, :m] = (Rm @ Rm.T).cpu().numpy()
        rhs = np.zeros(m + 1); rhs[m] = -1
        try:
            c = np.linalg.solve(B, rhs)[:m]
            Fx = sum(float(ci) * Fi for ci, Fi in zip(c, Fs))
        except np.linalg.LinAlgError:
            Fx = Fs[-1]
        D = density(Fx)
        ev[2].record(); torch.cuda.synchronize()
        if timings is not None:
            timings.setdefault("jk", []).append(ev[0].elapsed_time(ev[1])); timings.setdefault("linalg", []).append(ev[1].elapsed_time(ev[2]))
    raise RuntimeError("Convergence failed!")
