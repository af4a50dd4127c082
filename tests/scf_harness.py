"""Minimal RHF / UHF drivers used as TEST HARNESS around the J/K path (not product code).

Restates the callers' combination rules:
  RHF  (src/HartreeFockKohnSham/Restricted/SP.cpp:38-73):  D = C_occ C_occ^T (no factor 2),
        (J,K) = ContractInts(D,0x0,0x0), F = Hcore + J - K, E = sum D o (2 Hcore + J - K),
        residual 2(F D S - S D F), converged when max|resid| < 1e-6 (:77-79).
  UHF  (Unrestricted/SP.cpp:42-94): (J,Ka,Kb) = ContractInts(0x0,Da,Db), F_s = Hcore + J - K_s,
        E = 1/2 [Da o (Hcore+Fa) + Db o (Hcore+Fb)] (:75).
Pulay CDIIS as in src/DIIS/CDIIS.cpp:24-47 (the ADIIS warm-up phase needs Maniverse and only
changes the path to convergence, so it is skipped).
`jk` is any callable (Dd, Da, Db) -> (J, Kd, Ka, Kb): the oracle or the CUDA engine.
"""
import numpy as np
from scipy.linalg import eigh


def nuclear_repulsion(Z, xyz):
    e = 0.0
    for i in range(len(Z)):
        for j in range(i):
            e += Z[i] * Z[j] / np.linalg.norm(xyz[i] - xyz[j])
    return e


def _diis_extrapolate(Fs, Rs):
    n = len(Fs)
    B = -np.ones((n + 1, n + 1))
    B[n, n] = 0
    for i in range(n):
        for j in range(n):
            B[i, j] = np.vdot(Rs[i], Rs[j])
    rhs = np.zeros(n + 1)
    rhs[n] = -1
    try:
        c = np.linalg.solve(B, rhs)[:n]
    except np.linalg.LinAlgError:
        return Fs[-1]
    return sum(ci * Fi for ci, Fi in zip(c, Fs))


def rhf(S, Hcore, nocc, jk, e_nuc=0.0, D0=None, max_iter=100, tol=1e-8, diis_space=12, verbose=False):
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T
    def density(F):
        e, Cp = np.linalg.eigh(X.T @ F @ X)
        C_ = X @ Cp
        return C_[:, :nocc] @ C_[:, :nocc].T
    D = density(Hcore) if D0 is None else D0
    Fs, Rs = [], []
    E = 0.0
    for it in range(max_iter):
        J, K, _, _ = jk(D, None, None)
        F = Hcore + J - K
        E = np.sum(D * (2 * Hcore + J - K)) + e_nuc
        R = 2 * (F @ D @ S - S @ D @ F)
        err = np.abs(R).max()
        if verbose:
            print("  it %2d  E = %.10f  |R| = %.2e" % (it, E, err))
        if err < tol:
            return E, D, F, it
        Fs.append(F); Rs.append(X.T @ R @ X)
        Fs, Rs = Fs[-diis_space:], Rs[-diis_space:]
        D = density(_diis_extrapolate(Fs, Rs))
    raise RuntimeError("Convergence failed!")  # same message as Restricted/SP.cpp:76


def rhf_incremental(S, Hcore, nocc, jk, e_nuc=0.0, max_iter=100, tol=1e-8, diis_space=12, rebuild_every=8, log=None):
    """Direct-SCF variant of `rhf`: G = J - K is updated with the density DIFFERENCE,
    G_n = G_(n-1) + G[D_n - D_(n-1)] (the J/K build is linear in D), with a full rebuild every `rebuild_every`
    iterations so that screening errors cannot pile up.  `jk(D, full)` -> (J, K); `full` tells the caller whether
    the matrix is a whole density (no density-weighted screening) or a difference.  Same convergence test as `rhf`."""
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T
    def density(F):
        e, Cp = np.linalg.eigh(X.T @ F @ X)
        C_ = X @ Cp
        return C_[:, :nocc] @ C_[:, :nocc].T
    D = density(Hcore)
    Fs, Rs = [], []
    G, D_prev = None, None
    for it in range(max_iter):
        full = G is None or it % rebuild_every == 0
        if full:
            J, K = jk(D, True)
            G = J - K
        else:
            J, K = jk(D - D_prev, False)
            G = G + (J - K)
        D_prev = D
        if log is not None:
            log.append(full)
        F = Hcore + G
        E = np.sum(D * (2 * Hcore + G)) + e_nuc
        R = 2 * (F @ D @ S - S @ D @ F)
        err = np.abs(R).max()
        if err < tol:
            return E, D, F, it
        Fs.append(F); Rs.append(X.T @ R @ X)
        Fs, Rs = Fs[-diis_space:], Rs[-diis_space:]
        D = density(_diis_extrapolate(Fs, Rs))
    raise RuntimeError("Convergence failed!")


def uhf(S, Hcore, na, nb, jk, e_nuc=0.0, D0=None, max_iter=200, tol=1e-8, diis_space=12, verbose=False, level_shift=0.0,
        trace=None, raise_on_fail=True):
    """`trace`: list that receives the energy of every iteration; raise_on_fail=False returns the last iterate after
    max_iter instead of raising (parity of two runs that take the same steps does not need convergence)."""
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T
    def density(F, n):
        e, Cp = np.linalg.eigh(X.T @ F @ X)
        C_ = X @ Cp
        return C_[:, :n] @ C_[:, :n].T
    if D0 is None:
        Da, Db = density(Hcore, na), density(Hcore, nb)
    else:
        Da, Db = D0
    Fs, Rs = [], []
    for it in range(max_iter):
        J, _, Ka, Kb = jk(None, Da, Db)
        Fa, Fb = Hcore + J - Ka, Hcore + J - Kb
        E = 0.5 * (np.sum(Da * (Hcore + Fa)) + np.sum(Db * (Hcore + Fb))) + e_nuc
        Ra = Fa @ Da @ S - S @ Da @ Fa
        Rb = Fb @ Db @ S - S @ Db @ Fb
        err = max(np.abs(Ra).max(), np.abs(Rb).max())
        if trace is not None:
            trace.append(E)
        if verbose:
            print("  it %2d  E = %.10f  |R| = %.2e" % (it, E, err))
        if err < tol or (not raise_on_fail and it == max_iter - 1):
            return E, (Da, Db), (Fa, Fb), it
        Fs.append(np.stack([Fa, Fb])); Rs.append(np.stack([X.T @ Ra @ X, X.T @ Rb @ X]))
        Fs, Rs = Fs[-diis_space:], Rs[-diis_space:]
        Fx = _diis_extrapolate(Fs, Rs)
        if level_shift:
            Fx = Fx + level_shift * np.stack([S - S @ Da @ S, S - S @ Db @ S])
        Da, Db = density(Fx[0], na), density(Fx[1], nb)
    raise RuntimeError("Convergence failed!")


def rhf_total_gradient(oracle, fb, mol, D, F, S, nocc, g2e, h=1e-4):
    """Total RHF nuclear gradient around the 2e contraction under test, assembled like Restricted/Grad.cpp:60-70:
    Gradient = 2 D o dT + 2 D o dV - 2 W o dS + ContractGrads(D, D) + dE_nuc, W = C_occ eps C_occ^T (energy-weighted
    density).  The one-electron derivative integrals (Int2C1E in the reference, outside this engine's scope) are taken by
    central differences of the oracle's S/T/V with the atom's basis functions AND its nucleus displaced."""
    import copy
    from scipy.linalg import eigh as _eigh
    eps, C_ = _eigh(F, S)
    W = (C_[:, :nocc] * eps[:nocc]) @ C_[:, :nocc].T
    s2a = np.asarray(fb.shell2atom)
    natom = len(mol.Z)
    g = np.zeros((natom, 3))

    def displaced(atom, x, d):
        fb2 = copy.deepcopy(fb)
        c = np.array(fb2.center_xyz, dtype=np.float64).reshape(-1, 3).copy()
        c[s2a == atom, x] += d
        fb2.center_xyz = c.reshape(np.shape(fb.center_xyz))
        xyz = np.array(mol.xyz_bohr, dtype=np.float64).copy()
        xyz[atom, x] += d
        return fb2, xyz

    for a in range(natom):
        for x in range(3):
            fp, xp = displaced(a, x, h)
            fm, xm = displaced(a, x, -h)
            Sp, Tp, Vp = oracle.one_electron(fp, mol.Z, xp)
            Sm, Tm, Vm = oracle.one_electron(fm, mol.Z, xm)
            dH = ((Tp + Vp) - (Tm + Vm)) / (2 * h)
            dS = (Sp - Sm) / (2 * h)
            dEn = (nuclear_repulsion(mol.Z, xp) - nuclear_repulsion(mol.Z, xm)) / (2 * h)
            g[a, x] = 2 * np.sum(D * dH) - 2 * np.sum(W * dS) + g2e[3 * a + x] + dEn
    return g


# forces (= -gradient, Hartree/bohr) Chinium recorded for CH3ClF- RHF/cc-pVDZ: tools/sn2/sn2.cnm.log:211-216
SN2_FORCES_LOG = np.array([[-0.000000540, -0.000000141, 0.011105685], [-0.000000007, -0.001921870, 0.018273089],
                           [0.001664462, 0.000961020, 0.018272376], [-0.001664066, 0.000960766, 0.018273287],
                           [0.000000384, -0.000000163, -0.006215591], [-0.000000233, 0.000000389, -0.059708847]])


def core_density(S, Hcore, nocc):
    """D_core of SURVEY 8d: occupied projector of Hcore in the S-orthonormal basis."""
    e, C_ = eigh(Hcore, S)
    return C_[:, :nocc] @ C_[:, :nocc].T


def random_symmetric_density(nbf, seed=0):
    """stress density of SURVEY 8d: D=(A+A^T)/2, A_ij ~ U(-1,1)/nbf"""
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (nbf, nbf)) / nbf
    return 0.5 * (A + A.T)
