"""Second, independent CPU route to the ERI values: Obara-Saika / Head-Gordon-Pople recursion in extended precision.

TEST INFRASTRUCTURE (pins the oracle; never a product path).  The oracle (oracle/oracle.c) evaluates ERIs by the
McMurchie-Davidson scheme (Hermite expansion coefficients E and Hermite Coulomb integrals R); the device kernels use
Rys quadrature.  This module is a third formulation with no shared intermediates:
    [00|00]^(m) = 2 pi^(5/2) / (zeta eta sqrt(zeta+eta)) K_ab K_cd F_m(T)                       (Boys function from mpmath)
    VRR  [a+1_i 0|c 0]^(m) = PA_i [a0|c0]^(m) + WP_i [a0|c0]^(m+1)
                             + a_i/(2 zeta) ([a-1_i 0|c0]^(m) - rho/zeta [a-1_i 0|c0]^(m+1)) + c_i/(2(zeta+eta)) [a0|c-1_i 0]^(m+1)
         (and the same with bra <-> ket),
    HRR  [a b+1_i|cd] = [a+1_i b|cd] + AB_i [ab|cd]        (bra, then ket)
(S. Obara, A. Saika, J. Chem. Phys. 84, 3963 (1986); M. Head-Gordon, J. A. Pople, J. Chem. Phys. 89, 5777 (1988)), run in
mpmath with 40 digits so that its results are exact to double rounding.  Cartesian Gaussians x^i y^j z^k exp(-alpha r^2),
components ordered lx descending, then ly descending (the oracle's and the kernels' order).
"""
import functools

import mpmath as mp

mp.mp.dps = 40


def cart_components(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def boys(m, T):
    T = mp.mpf(T)
    if T < mp.mpf("1e-30"):
        return mp.mpf(1) / (2 * m + 1)
    return mp.gammainc(m + mp.mpf(1) / 2, 0, T) / (2 * T ** (m + mp.mpf(1) / 2))


def primitive_quartet(la, lb, lc, ld, ea, eb, ec, ed, A, B, C, D):
    """-> nested dict-free list [ia][ib][ic][id] of UNNORMALISED primitive Cartesian ERIs (mpf)."""
    A, B, C, D = ([mp.mpf(x) for x in v] for v in (A, B, C, D))
    ea, eb, ec, ed = (mp.mpf(x) for x in (ea, eb, ec, ed))
    zeta, eta = ea + eb, ec + ed
    P = [(ea * A[i] + eb * B[i]) / zeta for i in range(3)]
    Q = [(ec * C[i] + ed * D[i]) / eta for i in range(3)]
    W = [(zeta * P[i] + eta * Q[i]) / (zeta + eta) for i in range(3)]
    rho = zeta * eta / (zeta + eta)
    AB2 = sum((A[i] - B[i]) ** 2 for i in range(3))
    CD2 = sum((C[i] - D[i]) ** 2 for i in range(3))
    PQ2 = sum((P[i] - Q[i]) ** 2 for i in range(3))
    Kab = mp.exp(-ea * eb / zeta * AB2)
    Kcd = mp.exp(-ec * ed / eta * CD2)
    pref = 2 * mp.pi ** mp.mpf("2.5") / (zeta * eta * mp.sqrt(zeta + eta)) * Kab * Kcd
    L = la + lb + lc + ld
    F = [pref * boys(m, rho * PQ2) for m in range(L + 1)]
    PA = [P[i] - A[i] for i in range(3)]
    QC = [Q[i] - C[i] for i in range(3)]
    WP = [W[i] - P[i] for i in range(3)]
    WQ = [W[i] - Q[i] for i in range(3)]
    AB = [A[i] - B[i] for i in range(3)]
    CD = [C[i] - D[i] for i in range(3)]

    def dec(t, i):
        return tuple(t[k] - (1 if k == i else 0) for k in range(3))

    def inc(t, i):
        return tuple(t[k] + (1 if k == i else 0) for k in range(3))

    @functools.lru_cache(maxsize=None)
    def vrr(a, c, m):
        if min(a) < 0 or min(c) < 0:
            return mp.mpf(0)
        if a == (0, 0, 0) and c == (0, 0, 0):
            return F[m]
        if a != (0, 0, 0):
            i = next(k for k in range(3) if a[k] > 0)
            a1 = dec(a, i)
            v = PA[i] * vrr(a1, c, m) + WP[i] * vrr(a1, c, m + 1)
            if a1[i] > 0:
                v += a1[i] / (2 * zeta) * (vrr(dec(a1, i), c, m) - rho / zeta * vrr(dec(a1, i), c, m + 1))
            if c[i] > 0:
                v += c[i] / (2 * (zeta + eta)) * vrr(a1, dec(c, i), m + 1)
            return v
        i = next(k for k in range(3) if c[k] > 0)
        c1 = dec(c, i)
        v = QC[i] * vrr(a, c1, m) + WQ[i] * vrr(a, c1, m + 1)
        if c1[i] > 0:
            v += c1[i] / (2 * eta) * (vrr(a, dec(c1, i), m) - rho / eta * vrr(a, dec(c1, i), m + 1))
        return v          # a == 0 here: no cross term

    @functools.lru_cache(maxsize=None)
    def hrr_ket(a, c, d):
        if d == (0, 0, 0):
            return vrr(a, c, 0)
        i = next(k for k in range(3) if d[k] > 0)
        d1 = dec(d, i)
        return hrr_ket(a, inc(c, i), d1) + CD[i] * hrr_ket(a, c, d1)

    @functools.lru_cache(maxsize=None)
    def hrr_bra(a, b, c, d):
        if b == (0, 0, 0):
            return hrr_ket(a, c, d)
        i = next(k for k in range(3) if b[k] > 0)
        b1 = dec(b, i)
        return hrr_bra(inc(a, i), b1, c, d) + AB[i] * hrr_bra(a, b1, c, d)

    ca, cb, cc, cd = (cart_components(l) for l in (la, lb, lc, ld))
    return [[[[hrr_bra(a, b, c, d) for d in cd] for c in cc] for b in cb] for a in ca]


def contracted_quartet(shells):
    """shells: four tuples (l, exps, coefs, centre); coefs multiply the bare Cartesian primitives.  -> numpy float64 block."""
    import numpy as np
    (la, xa, ka, A), (lb, xb, kb, B), (lc, xc, kc, C), (ld, xd, kd, D) = shells
    na, nb, nc, nd = (len(cart_components(l)) for l in (la, lb, lc, ld))
    tot = [[[[mp.mpf(0)] * nd for _ in range(nc)] for _ in range(nb)] for _ in range(na)]
    for ea, wa in zip(xa, ka):
        for eb, wb in zip(xb, kb):
            for ec, wc in zip(xc, kc):
                for ed, wd in zip(xd, kd):
                    blk = primitive_quartet(la, lb, lc, ld, ea, eb, ec, ed, A, B, C, D)
                    w = mp.mpf(wa) * mp.mpf(wb) * mp.mpf(wc) * mp.mpf(wd)
                    for i in range(na):
                        for j in range(nb):
                            for k in range(nc):
                                for l in range(nd):
                                    tot[i][j][k][l] += w * blk[i][j][k][l]
    return np.array([[[[float(tot[i][j][k][l]) for l in range(nd)] for k in range(nc)] for j in range(nb)] for i in range(na)])
