/* oracle.c -- CPU ORACLE for the J/K path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (chinium_b200/csrc) never does.
 *
 * What it restates, and from where (paths relative to /root/reference):
 *   - enumeration, uniqueness predicate, degeneracy weights, digestion, symmetrisation,
 *     EXX scaling: literal restatement of src/Integral/Int4C2E.cpp:19-302 and :589-671
 *     (functions ref_*, same loop nests and index types).
 *   - the ERI VALUES: the reference gets them from libint2 (linked as -lint2, makefile:66;
 *     NOT vendored, NO version pinned anywhere in the repo).  libint2 cannot be built here, so
 *     the values come from an independent textbook McMurchie-Davidson scheme (Hermite
 *     expansion coefficients E, Hermite Coulomb integrals R from the Boys function), in
 *     libint2's conventions for shell order / pure ordering / normalisation
 *     (src/Integral/Macro.h:1-25, src/Grid/AO/Pure*.hpp, src/Integral/Normalization.cpp).
 *   - PARITY PIN: weak external pin only -- the RHF/cc-pVDZ energy of CH3ClF- recorded in
 *     tools/sn2/sn2.cnm.log:204 (reproduced by tests/test_oracle.py through this oracle to
 *     < 1e-7 Eh), plus first-principles checks (Boys vs mpmath, permutational symmetry,
 *     diag(S)=1, brute-force einsum); for the derivative path the forces recorded in the same
 *     log (:211-216), reproduced to < 7e-6 Eh/bohr (tests/test_oracle.py::test_sn2_recorded_forces); the 12 first- and 78
 *     second-derivative buffers agree to 1e-12 with analytic derivatives assembled from 40-digit Obara-Saika integrals
 *     (tests/test_oracle.py::test_derivative_buffers_against_obara_saika).
 *     The device kernels use a DIFFERENT algorithm (Rys
 *     quadrature), so oracle == device agreement is a two-route check.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/chinium_fock.h"

#define LMAX 6
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)
#define NCMAX NCART(LMAX)
#define L4MAX (4 * LMAX)
#define PI 3.14159265358979323846264338327950288

/* ---------------------------------------------------------------- Boys function F_m(T), m=0..mmax */
static void boys(int mmax, double T, double* F) {
    if (T < 1e-15) {
        for (int m = 0; m <= mmax; m++) F[m] = 1.0 / (2 * m + 1);
        return;
    }
    double eT = exp(-T);
    if (T < 50.0 + mmax) {
        /* all-positive series for the top order, then downward recursion (both stable) */
        double term = 1.0 / (2 * mmax + 1), sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2.0 * T / (2 * mmax + 2 * k + 1);
            sum += term;
            if (term < 1e-18 * sum) break;
        }
        F[mmax] = eT * sum;
        for (int m = mmax; m > 0; m--) F[m - 1] = (2.0 * T * F[m] + eT) / (2 * m - 1);
    } else {
        F[0] = 0.5 * sqrt(PI / T) * erf(sqrt(T));
        for (int m = 0; m < mmax; m++) F[m + 1] = ((2 * m + 1) * F[m] - eT) / (2.0 * T);
    }
}
void oracle_boys(int mmax, double T, double* F) { boys(mmax, T, F); }

/* ---------------------------------------------------------------- shells */
static inline int sh_l(const cf_basis* b, int s) { return abs(b->type[s]); }
static inline int sh_nfun(const cf_basis* b, int s) {
    int t = b->type[s], l = abs(t);
    return t < 0 ? 2 * l + 1 : NCART(l);
}
int oracle_nbf(const cf_basis* b) {
    int n = 0;
    for (int s = 0; s < b->nshell; s++) n += sh_nfun(b, s);
    return n;
}
static void shell2bf(const cf_basis* b, int* s2bf) {
    int n = 0;
    for (int s = 0; s < b->nshell; s++) { s2bf[s] = n; n += sh_nfun(b, s); }
}

/* Cartesian component order: lx descending, then ly descending (p: x,y,z). */
static void cart_components(int l, int (*c)[3]) {
    int n = 0;
    for (int lx = l; lx >= 0; lx--)
        for (int ly = l - lx; ly >= 0; ly--) { c[n][0] = lx; c[n][1] = ly; c[n][2] = l - lx - ly; n++; }
}
static int cart_index(int l, int lx, int ly) { /* position of (lx,ly,l-lx-ly) in the order above */
    int n = 0;
    for (int x = l; x > lx; x--) n += l - x + 1;
    return n + (l - lx - ly);
}

static double fact(int n) { double r = 1; for (int i = 2; i <= n; i++) r *= i; return r; }
static double binom(int n, int k) { if (k < 0 || k > n) return 0; return fact(n) / (fact(k) * fact(n - k)); }

/* Racah-normalised real solid harmonics as monomial coefficients, m = -l..l
 * (the polynomials of src/Grid/AO/PureD.hpp .. PureI.hpp).  C is [2l+1][NCART(l)]. */
static void pure_matrix(int l, double* C) {
    int nc = NCART(l);
    memset(C, 0, sizeof(double) * (2 * l + 1) * nc);
    for (int m = -l; m <= l; m++) {
        int am = abs(m);
        double N = sqrt(2.0 * fact(l + am) * fact(l - am) / (m == 0 ? 2.0 : 1.0)) / (pow(2.0, am) * fact(l));
        for (int t = 0; t <= (l - am) / 2; t++)
            for (int u = 0; u <= t; u++) {
                /* 2v runs over even (m>=0) or odd (m<0) integers <= |m| */
                for (int v2 = (m >= 0 ? 0 : 1); v2 <= am; v2 += 2) {
                    int k = (m >= 0) ? v2 / 2 : (v2 - 1) / 2;
                    double c = ((t + k) % 2 ? -1.0 : 1.0) * pow(0.25, t) * binom(l, t) * binom(l - t, am + t) *
                               binom(t, u) * binom(am, v2);
                    int ex = 2 * t + am - 2 * u - v2, ey = 2 * u + v2;
                    if (ex < 0) continue;
                    C[(m + l) * nc + cart_index(l, ex, ey)] += N * c;
                }
            }
    }
}
/* transformation of one shell: rows = functions of the shell as the reference orders them */
static int shell_transform(int type, double* C) {
    int l = abs(type), nc = NCART(l);
    if (type >= 0) { /* S, Cartesian P (x,y,z): identity (Cartesian l>=2 is never produced by the reference's reader) */
        memset(C, 0, sizeof(double) * nc * nc);
        for (int i = 0; i < nc; i++) C[i * nc + i] = 1.0;
        return nc;
    }
    pure_matrix(l, C);
    return 2 * l + 1;
}
void oracle_pure_matrix(int l, double* C) { pure_matrix(l, C); }

/* ---------------------------------------------------------------- Hermite expansion coefficients
 * E[i][j][t], 0<=i<=la, 0<=j<=lb, 0<=t<=i+j, for one Cartesian direction; E^{00}_0 = 1
 * (the Gaussian-product prefactor is carried separately). */
#define EDIM (LMAX + 1)
typedef struct { double e[EDIM][EDIM][2 * LMAX + 1]; } ecoef;
static void hermite_E(int la, int lb, double p, double PA, double PB, ecoef* E) {
    memset(E, 0, sizeof(*E));
    double h = 0.5 / p;
    E->e[0][0][0] = 1.0;
    for (int i = 0; i <= la; i++) {
        if (i > 0)
            for (int t = 0; t <= i; t++) {
                double v = PA * E->e[i - 1][0][t];
                if (t > 0) v += h * E->e[i - 1][0][t - 1];
                if (t + 1 <= i - 1) v += (t + 1) * E->e[i - 1][0][t + 1];
                E->e[i][0][t] = v;
            }
        for (int j = 1; j <= lb; j++)
            for (int t = 0; t <= i + j; t++) {
                double v = PB * E->e[i][j - 1][t];
                if (t > 0) v += h * E->e[i][j - 1][t - 1];
                if (t + 1 <= i + j - 1) v += (t + 1) * E->e[i][j - 1][t + 1];
                E->e[i][j][t] = v;
            }
    }
}

/* per-thread scratch that grows on demand and is reused: no allocation inside the primitive / shell-quartet loops
 * (a CPU baseline that mallocs per primitive quartet would be a strawman) */
#define TLS_SLOTS 6
static __thread double* tls_buf[TLS_SLOTS];
static __thread size_t tls_cap[TLS_SLOTS];
static double* tls_get(int slot, size_t n) {
    if (tls_cap[slot] < n) {
        free(tls_buf[slot]);
        tls_buf[slot] = (double*)malloc(sizeof(double) * n);
        tls_cap[slot] = n;
    }
    return tls_buf[slot];
}

/* Hermite Coulomb integrals R_{tuv} = R^0_{tuv}(alpha, PQ), t+u+v <= L. out[t][u][v], dim L+1 each */
static void hermite_R(int L, double alpha, const double* PQ, double* out /* (L+1)^3 */) {
    double F[L4MAX + 1];
    double T = alpha * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]);
    boys(L, T, F);
    int d = L + 1;
    /* work[n][t][u][v] built downward in n (every entry read below has been written: level n holds t+u+v <= L-n) */
    size_t sz = (size_t)d * d * d;
    double* w = tls_get(0, sz * (L + 1));
    memset(w, 0, sz * sizeof(double));      /* level 0 is copied out whole; entries with t+u+v > L stay 0 */
#define RW(n, t, u, v) w[(size_t)(n) * sz + ((size_t)(t) * d + (u)) * d + (v)]
    double f = 1.0;
    for (int n = 0; n <= L; n++) { RW(n, 0, 0, 0) = f * F[n]; f *= -2.0 * alpha; }
    for (int n = L - 1; n >= 0; n--) {
        int M = L - n; /* t+u+v <= M at level n */
        for (int t = 0; t <= M; t++)
            for (int u = 0; t + u <= M; u++)
                for (int v = 0; t + u + v <= M; v++) {
                    if (t + u + v == 0) continue;
                    double r;
                    if (t > 0) {
                        r = PQ[0] * RW(n + 1, t - 1, u, v);
                        if (t > 1) r += (t - 1) * RW(n + 1, t - 2, u, v);
                    } else if (u > 0) {
                        r = PQ[1] * RW(n + 1, t, u - 1, v);
                        if (u > 1) r += (u - 1) * RW(n + 1, t, u - 2, v);
                    } else {
                        r = PQ[2] * RW(n + 1, t, u, v - 1);
                        if (v > 1) r += (v - 1) * RW(n + 1, t, u, v - 2);
                    }
                    RW(n, t, u, v) = r;
                }
    }
    memcpy(out, w, sz * sizeof(double));
#undef RW
}

/* ---------------------------------------------------------------- primitive-pair data of a shell pair */
typedef struct {
    int la, lb, npp;
    double AB[3];
    double* p;      /* [npp] */
    double* P;      /* [npp][3] */
    double* K;      /* [npp] ca*cb*exp(-mu AB^2) */
    ecoef* E;       /* [npp][3] */
} pairdata;

static void make_pair(const cf_basis* b, int sa, int sb, pairdata* pd) {
    int la = sh_l(b, sa), lb = sh_l(b, sb);
    int na = b->nprim[sa], nb = b->nprim[sb];
    const double* A = b->center_xyz + 3 * sa;
    const double* B = b->center_xyz + 3 * sb;
    pd->la = la; pd->lb = lb; pd->npp = na * nb;
    double r2 = 0;
    for (int x = 0; x < 3; x++) { pd->AB[x] = A[x] - B[x]; r2 += pd->AB[x] * pd->AB[x]; }
    pd->p = (double*)malloc(sizeof(double) * pd->npp);
    pd->P = (double*)malloc(sizeof(double) * 3 * pd->npp);
    pd->K = (double*)malloc(sizeof(double) * pd->npp);
    pd->E = (ecoef*)malloc(sizeof(ecoef) * 3 * pd->npp);
    int n = 0;
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++, n++) {
            double a = b->exps[b->prim_offset[sa] + i], bb = b->exps[b->prim_offset[sb] + j];
            double ca = b->coefs_normalized[b->prim_offset[sa] + i], cb = b->coefs_normalized[b->prim_offset[sb] + j];
            double p = a + bb;
            pd->p[n] = p;
            pd->K[n] = ca * cb * exp(-a * bb / p * r2);
            for (int x = 0; x < 3; x++) {
                double Px = (a * A[x] + bb * B[x]) / p;
                pd->P[3 * n + x] = Px;
                hermite_E(la, lb, p, Px - A[x], Px - B[x], &pd->E[3 * n + x]);
            }
        }
}
static void free_pair(pairdata* pd) { free(pd->p); free(pd->P); free(pd->K); free(pd->E); }

/* Cartesian (ab|cd) of a shell quartet from two pairdata; out[na_c][nb_c][nc_c][nd_c] (accumulated from 0) */
static void eri_cart(const pairdata* ab, const pairdata* cd, double* out) {
    int la = ab->la, lb = ab->lb, lc = cd->la, ld = cd->lb;
    int nca = NCART(la), ncb = NCART(lb), ncc = NCART(lc), ncd = NCART(ld);
    int Lab = la + lb, Lcd = lc + ld, L = Lab + Lcd, d = L + 1;
    int ca[NCMAX][3], cb[NCMAX][3], cc[NCMAX][3], cdd[NCMAX][3];
    cart_components(la, ca); cart_components(lb, cb); cart_components(lc, cc); cart_components(ld, cdd);
    size_t ntot = (size_t)nca * ncb * ncc * ncd;
    memset(out, 0, sizeof(double) * ntot);
    double* R = tls_get(1, (size_t)d * d * d);
    int dab = Lab + 1;
    double* G = tls_get(2, (size_t)dab * dab * dab);
    for (int i = 0; i < ab->npp; i++)
        for (int j = 0; j < cd->npp; j++) {
            double p = ab->p[i], q = cd->p[j];
            double alpha = p * q / (p + q);
            double PQ[3] = {ab->P[3 * i] - cd->P[3 * j], ab->P[3 * i + 1] - cd->P[3 * j + 1], ab->P[3 * i + 2] - cd->P[3 * j + 2]};
            double pref = 2.0 * pow(PI, 2.5) / (p * q * sqrt(p + q)) * ab->K[i] * cd->K[j];
            hermite_R(L, alpha, PQ, R);
            const ecoef* Eab = &ab->E[3 * i];
            const ecoef* Ecd = &cd->E[3 * j];
            for (int ic = 0; ic < ncc; ic++)
                for (int id = 0; id < ncd; id++) {
                    int mx = cc[ic][0] + cdd[id][0], my = cc[ic][1] + cdd[id][1], mz = cc[ic][2] + cdd[id][2];
                    const double* ex = Ecd[0].e[cc[ic][0]][cdd[id][0]];
                    const double* ey = Ecd[1].e[cc[ic][1]][cdd[id][1]];
                    const double* ez = Ecd[2].e[cc[ic][2]][cdd[id][2]];
                    /* G[t][u][v] = sum_{tau,nu,phi} (-1)^(tau+nu+phi) Ecd R[t+tau][u+nu][v+phi] */
                    for (int t = 0; t <= Lab; t++)
                        for (int u = 0; t + u <= Lab; u++)
                            for (int v = 0; t + u + v <= Lab; v++) {
                                double s = 0;
                                for (int a = 0; a <= mx; a++)
                                    for (int bq = 0; bq <= my; bq++) {
                                        double exy = ex[a] * ey[bq];
                                        for (int c = 0; c <= mz; c++) {
                                            double sg = ((a + bq + c) & 1) ? -1.0 : 1.0;
                                            s += sg * exy * ez[c] * R[((t + a) * d + (u + bq)) * d + (v + c)];
                                        }
                                    }
                                G[(t * dab + u) * dab + v] = s;
                            }
                    for (int ia = 0; ia < nca; ia++)
                        for (int ib = 0; ib < ncb; ib++) {
                            int nx = ca[ia][0] + cb[ib][0], ny = ca[ia][1] + cb[ib][1], nz = ca[ia][2] + cb[ib][2];
                            const double* fx = Eab[0].e[ca[ia][0]][cb[ib][0]];
                            const double* fy = Eab[1].e[ca[ia][1]][cb[ib][1]];
                            const double* fz = Eab[2].e[ca[ia][2]][cb[ib][2]];
                            double s = 0;
                            for (int t = 0; t <= nx; t++)
                                for (int u = 0; u <= ny; u++) {
                                    double fxy = fx[t] * fy[u];
                                    for (int v = 0; v <= nz; v++) s += fxy * fz[v] * G[(t * dab + u) * dab + v];
                                }
                            out[(((size_t)ia * ncb + ib) * ncc + ic) * ncd + id] += pref * s;
                        }
                }
        }
}

/* transform index `which` (0..3) of a 4-index tensor with dims n[4] by matrix C [nnew][n[which]] */
static void transform_index(const double* in, double* out, const int* n, int which, const double* C, int nnew) {
    size_t outer = 1, inner = 1;
    for (int i = 0; i < which; i++) outer *= n[i];
    for (int i = which + 1; i < 4; i++) inner *= n[i];
    int nold = n[which];
    for (size_t o = 0; o < outer; o++)
        for (int m = 0; m < nnew; m++)
            for (size_t k = 0; k < inner; k++) {
                double s = 0;
                for (int c = 0; c < nold; c++) s += C[m * nold + c] * in[(o * nold + c) * inner + k];
                out[(o * nnew + m) * inner + k] = s;
            }
}

/* (s1 s2|s3 s4) in the reference's function order; buf is the dense [n1][n2][n3][n4] tensor,
 * f4 fastest -- the layout the reference reads at src/Integral/Int4C2E.cpp:276-283. */
typedef struct { pairdata* pd; int nshell; } paircache;

static void eri_shell_quartet_pd(const cf_basis* b, const pairdata* ab, const pairdata* cd,
                                 int s1, int s2, int s3, int s4, double* buf) {
    int sh[4] = {s1, s2, s3, s4};
    int n[4];
    for (int i = 0; i < 4; i++) n[i] = NCART(sh_l(b, sh[i]));
    size_t ntot = (size_t)n[0] * n[1] * n[2] * n[3];
    double* t1 = tls_get(3, ntot);
    double* t2 = tls_get(4, ntot);
    eri_cart(ab, cd, t1);
    double C[(2 * LMAX + 1) * NCMAX];
    for (int i = 0; i < 4; i++) {
        int nnew = shell_transform(b->type[sh[i]], C);
        transform_index(t1, t2, n, i, C, nnew);
        n[i] = nnew;
        double* tmp = t1; t1 = t2; t2 = tmp;
    }
    memcpy(buf, t1, sizeof(double) * (size_t)n[0] * n[1] * n[2] * n[3]);
}
void oracle_eri_shell_quartet(const cf_basis* b, int s1, int s2, int s3, int s4, double* buf) {
    pairdata ab, cd;
    make_pair(b, s1, s2, &ab); make_pair(b, s3, s4, &cd);
    eri_shell_quartet_pd(b, &ab, &cd, s1, s2, s3, s4, buf);
    free_pair(&ab); free_pair(&cd);
}

/* all shell pairs s1>=s2 (plus s1<s2 on demand is never needed: every loop below has s2<=s1, s4<=s3 or
 * builds the pair explicitly) */
static pairdata* all_pairs(const cf_basis* b) {
    int ns = b->nshell;
    pairdata* pd = (pairdata*)malloc(sizeof(pairdata) * (size_t)ns * ns);
    for (int i = 0; i < ns; i++)
        for (int j = 0; j < ns; j++) make_pair(b, i, j, &pd[(size_t)i * ns + j]);
    return pd;
}
static void free_all_pairs(const cf_basis* b, pairdata* pd) {
    size_t n = (size_t)b->nshell * b->nshell;
    for (size_t i = 0; i < n; i++) free_pair(&pd[i]);
    free(pd);
}

/* ================================================================ one-electron integrals (harness) */
static void one_electron(const cf_basis* b, int natom, const double* Z, const double* xyz,
                         double* S, double* T, double* V) {
    int nbf = oracle_nbf(b), ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
#pragma omp parallel for schedule(dynamic, 1)   /* every (sa, sb) writes its own block of S, T, V */
    for (int sa = 0; sa < ns; sa++)
        for (int sb = 0; sb < ns; sb++) {
            int la = sh_l(b, sa), lb = sh_l(b, sb);
            int nca = NCART(la), ncb = NCART(lb);
            int ca[NCMAX][3], cb[NCMAX][3];
            cart_components(la, ca); cart_components(lb, cb);
            double sc[NCMAX * NCMAX] = {0}, tc[NCMAX * NCMAX] = {0}, vc[NCMAX * NCMAX] = {0};
            const double* A = b->center_xyz + 3 * sa;
            const double* B = b->center_xyz + 3 * sb;
            double r2 = 0;
            for (int x = 0; x < 3; x++) r2 += (A[x] - B[x]) * (A[x] - B[x]);
            int Lab = la + lb, d = Lab + 1;
            double* R = (double*)malloc(sizeof(double) * d * d * d);
            for (int i = 0; i < b->nprim[sa]; i++)
                for (int j = 0; j < b->nprim[sb]; j++) {
                    double a = b->exps[b->prim_offset[sa] + i], bb = b->exps[b->prim_offset[sb] + j];
                    double cc = b->coefs_normalized[b->prim_offset[sa] + i] * b->coefs_normalized[b->prim_offset[sb] + j];
                    double p = a + bb, K = cc * exp(-a * bb / p * r2);
                    double P[3];
                    ecoef E[3]; /* need j up to lb+2 for kinetic: build with lb+2 (<= LMAX+2 guard) */
                    double e2[3][EDIM][EDIM + 2]; /* overlap 1D s[i][j], j up to lb+2 */
                    for (int x = 0; x < 3; x++) {
                        P[x] = (a * A[x] + bb * B[x]) / p;
                        hermite_E(la, lb, p, P[x] - A[x], P[x] - B[x], &E[x]);
                        /* 1D overlaps S_ij = E^{ij}_0 sqrt(pi/p) via Obara-Saika-like recursion up to j = lb+2 */
                        double PA = P[x] - A[x], PB = P[x] - B[x], h = 0.5 / p;
                        for (int ii = 0; ii <= la; ii++)
                            for (int jj = 0; jj <= lb + 2; jj++) {
                                double v;
                                if (ii == 0 && jj == 0) v = sqrt(PI / p);
                                else if (jj == 0) v = PA * e2[x][ii - 1][0] + (ii > 1 ? (ii - 1) * h * e2[x][ii - 2][0] : 0);
                                else v = PB * e2[x][ii][jj - 1] + (ii > 0 ? ii * h * e2[x][ii - 1][jj - 1] : 0) +
                                         (jj > 1 ? (jj - 1) * h * e2[x][ii][jj - 2] : 0);
                                e2[x][ii][jj] = v;
                            }
                    }
                    for (int ia = 0; ia < nca; ia++)
                        for (int ib = 0; ib < ncb; ib++) {
                            double s1d[3], k1d[3];
                            for (int x = 0; x < 3; x++) {
                                int ii = ca[ia][x], jj = cb[ib][x];
                                s1d[x] = e2[x][ii][jj];
                                k1d[x] = -2.0 * bb * bb * e2[x][ii][jj + 2] + bb * (2 * jj + 1) * e2[x][ii][jj] -
                                         (jj > 1 ? 0.5 * jj * (jj - 1) * e2[x][ii][jj - 2] : 0);
                            }
                            sc[ia * ncb + ib] += K * s1d[0] * s1d[1] * s1d[2];
                            tc[ia * ncb + ib] += K * (k1d[0] * s1d[1] * s1d[2] + s1d[0] * k1d[1] * s1d[2] + s1d[0] * s1d[1] * k1d[2]);
                        }
                    for (int c = 0; c < natom; c++) {
                        double PC[3] = {P[0] - xyz[3 * c], P[1] - xyz[3 * c + 1], P[2] - xyz[3 * c + 2]};
                        hermite_R(Lab, p, PC, R);
                        for (int ia = 0; ia < nca; ia++)
                            for (int ib = 0; ib < ncb; ib++) {
                                double s = 0;
                                for (int t = 0; t <= ca[ia][0] + cb[ib][0]; t++)
                                    for (int u = 0; u <= ca[ia][1] + cb[ib][1]; u++)
                                        for (int v = 0; v <= ca[ia][2] + cb[ib][2]; v++)
                                            s += E[0].e[ca[ia][0]][cb[ib][0]][t] * E[1].e[ca[ia][1]][cb[ib][1]][u] *
                                                 E[2].e[ca[ia][2]][cb[ib][2]][v] * R[(t * d + u) * d + v];
                                vc[ia * ncb + ib] += -Z[c] * 2.0 * PI / p * K * s;
                            }
                    }
                }
            free(R);
            /* cart -> reference function order */
            double Ca[(2 * LMAX + 1) * NCMAX], Cb[(2 * LMAX + 1) * NCMAX];
            int na = shell_transform(b->type[sa], Ca), nb = shell_transform(b->type[sb], Cb);
            double* mats[3] = {sc, tc, vc};
            double* outs[3] = {S, T, V};
            for (int m = 0; m < 3; m++)
                for (int i = 0; i < na; i++)
                    for (int j = 0; j < nb; j++) {
                        double s = 0;
                        for (int x = 0; x < nca; x++)
                            for (int y = 0; y < ncb; y++) s += Ca[i * nca + x] * Cb[j * ncb + y] * mats[m][x * ncb + y];
                        outs[m][(size_t)(s2bf[sb] + j) * nbf + (s2bf[sa] + i)] = s; /* col-major (symmetric anyway) */
                    }
        }
    free(s2bf);
}
void oracle_one_electron(const cf_basis* b, int natom, const double* Z, const double* xyz, double* S, double* T, double* V) {
    one_electron(b, natom, Z, xyz, S, T, V);
}

/* ================================================================ literal restatement of the reference path
 * Index variables are `short` where the reference uses short int (Int4C2E.h:19-29). Matrices col-major. */
#define M(A, i, j) (A)[(size_t)(j) * nbf + (i)]

/* ::getRepulsionDiag, Int4C2E.cpp:19-77 (only Diag1212 is ever consumed: :515, :544) */
void ref_getRepulsionDiag(const cf_basis* b, double* Diag1212) {
    int nbf = oracle_nbf(b), ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
    memset(Diag1212, 0, sizeof(double) * (size_t)nbf * nbf);
    for (short s1 = 0; s1 < (short)ns; s1++) {
        short bf1_first = s2bf[s1], n1 = sh_nfun(b, s1);
        for (short s2 = 0; s2 <= s1; s2++) {
            short bf2_first = s2bf[s2], n2 = sh_nfun(b, s2);
            double* buf = (double*)malloc(sizeof(double) * (size_t)n1 * n2 * n1 * n2);
            oracle_eri_shell_quartet(b, s1, s2, s1, s2, buf);
            int f1234 = 0;
            for (short f1 = 0; f1 < n1; f1++) {
                short bf1 = f1 + bf1_first;
                for (short f2 = 0; f2 < n2; f2++) {
                    short bf2 = f2 + bf2_first;
                    for (short f3 = 0; f3 < n1; f3++) {
                        short bf3 = f3 + bf1_first;
                        for (short f4 = 0; f4 < n2; f4++, f1234++) {
                            short bf4 = f4 + bf2_first;
                            if (bf1 == bf3 && bf2 == bf4 && bf1 >= bf2) {
                                M(Diag1212, bf1, bf2) = buf[f1234];
                                M(Diag1212, bf2, bf1) = buf[f1234];
                            }
                        }
                    }
                }
            }
            free(buf);
        }
    }
    free(s2bf);
}

/* ::getRepulsionLength (:79-128) and ::getRepulsionIndices (:130-182) in one routine: when shells != NULL the
 * (s1,s2,s3,s4) are recorded with the `>=` test of :166, the counts use the `>` test of :111. */
void ref_getRepulsionLengthIndices(const cf_basis* b, const double* diag, double threshold,
                                   long* n2integrals_out, long* nshellquartets_out,
                                   short* shellis, short* shelljs, short* shellks, short* shellls) {
    int nbf = oracle_nbf(b), ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
    long n2integrals = 0, nshellquartets = 0, nrec = 0;
    for (short s1 = 0; s1 < (short)ns; s1++) {
        short bf1_first = s2bf[s1], n1 = sh_nfun(b, s1);
        for (short s2 = 0; s2 <= s1; s2++) {
            short bf2_first = s2bf[s2], n2 = sh_nfun(b, s2);
            for (short s3 = 0; s3 <= s1; s3++) {
                short bf3_first = s2bf[s3], n3 = sh_nfun(b, s3);
                for (short s4 = 0; s4 <= (s2 > s3 ? s2 : s3); s4++) {
                    short bf4_first = s2bf[s4], n4 = sh_nfun(b, s4);
                    int discard = 1, discard_ge = 1, uniquebf = 0;
                    for (short f1 = 0; f1 < n1; f1++) {
                        short bf1 = f1 + bf1_first;
                        for (short f2 = 0; f2 < n2; f2++) {
                            short bf2 = f2 + bf2_first;
                            for (short f3 = 0; f3 < n3; f3++) {
                                short bf3 = f3 + bf3_first;
                                for (short f4 = 0; f4 < n4; f4++) {
                                    short bf4 = f4 + bf4_first;
                                    if (bf2 <= bf1 && bf3 <= bf1 && bf4 <= ((bf1 == bf3) ? bf2 : bf3)) {
                                        uniquebf++;
                                        double upperbound = sqrt(fabs(M(diag, bf1, bf2) * M(diag, bf3, bf4)));
                                        if (upperbound > threshold) discard = 0;
                                        if (upperbound >= threshold) discard_ge = 0;
                                    }
                                }
                            }
                        }
                    }
                    if (!discard) { nshellquartets++; n2integrals += uniquebf; }
                    if (!discard_ge && shellis) {
                        shellis[nrec] = s1; shelljs[nrec] = s2; shellks[nrec] = s3; shellls[nrec] = s4;
                        nrec++;
                    }
                }
            }
        }
    }
    *n2integrals_out = n2integrals; *nshellquartets_out = nshellquartets;
    free(s2bf);
}

/* getRepulsion0 (:233-302): evaluate, keep unique function quartets, store value * abcd_deg.
 * The equal-count thread split of ::getThreadPointers (:184-231) only decides which thread writes which
 * slice of the SAME arrays, so a serial fill in list order produces the identical arrays. */
long ref_getRepulsion0(const cf_basis* b, long nsq, const short* shellis, const short* shelljs,
                       const short* shellks, const short* shellls,
                       short* basisis, short* basisjs, short* basisks, short* basisls, double* repulsionints) {
    int ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
    pairdata* pd = all_pairs(b);
    /* first pass: offsets per quartet so the fill can run in parallel like the reference's (:244) */
    long* head = (long*)malloc(sizeof(long) * (nsq + 1));
    head[0] = 0;
    for (long q = 0; q < nsq; q++) {
        short s1 = shellis[q], s2 = shelljs[q], s3 = shellks[q], s4 = shellls[q];
        long cnt = 0;
        for (short f1 = 0; f1 < sh_nfun(b, s1); f1++) for (short f2 = 0; f2 < sh_nfun(b, s2); f2++)
            for (short f3 = 0; f3 < sh_nfun(b, s3); f3++) for (short f4 = 0; f4 < sh_nfun(b, s4); f4++) {
                short bf1 = s2bf[s1] + f1, bf2 = s2bf[s2] + f2, bf3 = s2bf[s3] + f3, bf4 = s2bf[s4] + f4;
                if (bf2 <= bf1 && bf3 <= bf1 && bf4 <= ((bf1 == bf3) ? bf2 : bf3)) cnt++;
            }
        head[q + 1] = head[q] + cnt;
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (long q = 0; q < nsq; q++) {
        short s1 = shellis[q], s2 = shelljs[q], s3 = shellks[q], s4 = shellls[q];
        short bf1_first = s2bf[s1], bf2_first = s2bf[s2], bf3_first = s2bf[s3], bf4_first = s2bf[s4];
        short n1 = sh_nfun(b, s1), n2 = sh_nfun(b, s2), n3 = sh_nfun(b, s3), n4 = sh_nfun(b, s4);
        double* buf = (double*)malloc(sizeof(double) * (size_t)n1 * n2 * n3 * n4);
        eri_shell_quartet_pd(b, &pd[(size_t)s1 * ns + s2], &pd[(size_t)s3 * ns + s4], s1, s2, s3, s4, buf);
        long w = head[q];
        int f1234 = 0;
        for (short f1 = 0; f1 < n1; f1++) {
            short bf1 = bf1_first + f1;
            for (short f2 = 0; f2 < n2; f2++) {
                short bf2 = bf2_first + f2;
                char ab_deg = (bf1 == bf2) ? 1 : 2;
                for (short f3 = 0; f3 < n3; f3++) {
                    short bf3 = bf3_first + f3;
                    for (short f4 = 0; f4 < n4; f4++, f1234++) {
                        short bf4 = bf4_first + f4;
                        char cd_deg = (bf3 == bf4) ? 1 : 2;
                        char ab_cd_deg = (bf1 == bf3) ? (bf2 == bf4 ? 1 : 2) : 2;
                        char abcd_deg = ab_deg * cd_deg * ab_cd_deg;
                        if (bf2 <= bf1 && bf3 <= bf1 && bf4 <= ((bf1 == bf3) ? bf2 : bf3)) {
                            basisis[w] = bf1; basisjs[w] = bf2; basisks[w] = bf3; basisls[w] = bf4;
                            repulsionints[w] = buf[f1234] * abcd_deg;
                            w++;
                        }
                    }
                }
            }
        }
        free(buf);
    }
    long total = head[nsq];
    free(head); free_all_pairs(b, pd); free(s2bf);
    return total;
}

/* Gunified (:601-671). Dd/Da/Db may be NULL (= the reference's 0x0 matrices). Outputs nbf x nbf each. */
void ref_Gunified(const short* is, const short* js, const short* ks, const short* ls, const double* ints, long length,
                  int nbf, const double* Dd, const double* Da, const double* Db, double kscale, int nthreads,
                  double* J, double* Kd, double* Ka, double* Kb) {
    size_t n2 = (size_t)nbf * nbf;
    double* Dtot = (double*)calloc(n2, sizeof(double));
    for (size_t i = 0; i < n2; i++) Dtot[i] = (Dd ? 2 * Dd[i] : 0) + (Da ? Da[i] : 0) + (Db ? Db[i] : 0);
    double* raw = (double*)calloc(4 * n2 * nthreads, sizeof(double)); /* thread-private rawJ,rawKd,rawKa,rawKb */
    long fewer = length / nthreads;
    int nfewers = (int)(nthreads - length + fewer * nthreads);
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
    for (int ith = 0; ith < nthreads; ith++) {
        long head = 0;
        for (int t = 0; t < ith; t++) head += (t < nfewers) ? fewer : fewer + 1;
        long nints = (ith < nfewers) ? fewer : fewer + 1;
        double* rawJ = raw + (size_t)ith * 4 * n2;
        double* rawKd = rawJ + n2; double* rawKa = rawKd + n2; double* rawKb = rawKa + n2;
        for (long q = head; q < head + nints; q++) {
            short i = is[q], j = js[q], k = ks[q], l = ls[q];
            double repulsion = ints[q];
            M(rawJ, i, j) += M(Dtot, k, l) * repulsion;
            M(rawJ, k, l) += M(Dtot, i, j) * repulsion;
            if (kscale > 0.) {
                if (Dd) { M(rawKd, i, k) += M(Dd, j, l) * repulsion; M(rawKd, j, l) += M(Dd, i, k) * repulsion;
                          M(rawKd, i, l) += M(Dd, j, k) * repulsion; M(rawKd, j, k) += M(Dd, i, l) * repulsion; }
                if (Da) { M(rawKa, i, k) += M(Da, j, l) * repulsion; M(rawKa, j, l) += M(Da, i, k) * repulsion;
                          M(rawKa, i, l) += M(Da, j, k) * repulsion; M(rawKa, j, k) += M(Da, i, l) * repulsion; }
                if (Db) { M(rawKb, i, k) += M(Db, j, l) * repulsion; M(rawKb, j, l) += M(Db, i, k) * repulsion;
                          M(rawKb, i, l) += M(Db, j, k) * repulsion; M(rawKb, j, k) += M(Db, i, l) * repulsion; }
            }
        }
    }
    for (int t = 1; t < nthreads; t++)
        for (size_t i = 0; i < 4 * n2; i++) raw[i] += raw[(size_t)t * 4 * n2 + i];
    double* outs[4] = {J, Kd, Ka, Kb};
    double sc[4] = {0.25, 0.125 * kscale, 0.125 * kscale, 0.125 * kscale};
    for (int m = 0; m < 4; m++) {
        if (!outs[m]) continue;
        const double* r = raw + m * n2;
        for (int i = 0; i < nbf; i++)
            for (int j = 0; j < nbf; j++) M(outs[m], i, j) = sc[m] * (M(r, i, j) + M(r, j, i));
    }
    free(raw); free(Dtot);
}

/* ================================================================ direct build over canonical shell quartets
 * (B2 of BASELINE.md): same ERI values, integrals digested on the fly with shell-level degeneracy
 * weights; mathematically identical to the stored path.  stride/offset select a 1/stride sample of the
 * bra pairs for bounded timing runs (results are then partial).  counts: [0] canonical quartets visited,
 * [1] primitive quartets. */
void oracle_direct_jk(const cf_basis* b, int nbf, const double* Dd, const double* Da, const double* Db, double kscale,
                      double* J, double* Kd, double* Ka, double* Kb, int nthreads, int stride, int offset, long* counts) {
    int ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
    size_t n2 = (size_t)nbf * nbf;
    double* Dtot = (double*)calloc(n2, sizeof(double));
    for (size_t i = 0; i < n2; i++) Dtot[i] = (Dd ? 2 * Dd[i] : 0) + (Da ? Da[i] : 0) + (Db ? Db[i] : 0);
    pairdata* pd = all_pairs(b);
    if (nthreads < 1) nthreads = 1;
    double* raw = (double*)calloc(4 * n2 * nthreads, sizeof(double));
    const double* DX[3] = {Dd, Da, Db};
    long npairs = (long)ns * (ns + 1) / 2;
    long nq = 0, npq = 0;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1) reduction(+ : nq, npq)
    for (long ip = npairs - 1; ip >= 0; ip--) {
        if (stride > 1 && (ip % stride) != offset) continue;
        int s1 = (int)((sqrt(8.0 * ip + 1) - 1) / 2);
        while ((long)(s1 + 1) * (s1 + 2) / 2 <= ip) s1++;
        while ((long)s1 * (s1 + 1) / 2 > ip) s1--;
        int s2 = (int)(ip - (long)s1 * (s1 + 1) / 2);
#ifdef _OPENMP
        int ith = omp_get_thread_num();
#else
        int ith = 0;
#endif
        double* rawJ = raw + (size_t)ith * 4 * n2;
        int n1 = sh_nfun(b, s1), n2f = sh_nfun(b, s2);
        double* buf = (double*)malloc(sizeof(double) * (size_t)n1 * n2f * (2 * LMAX + 1) * (2 * LMAX + 1) * 4);
        for (int s3 = 0; s3 <= s1; s3++)
            for (int s4 = 0; s4 <= (s3 == s1 ? s2 : s3); s4++) {
                int n3 = sh_nfun(b, s3), n4 = sh_nfun(b, s4);
                const pairdata* ab = &pd[(size_t)s1 * ns + s2];
                const pairdata* cd = &pd[(size_t)s3 * ns + s4];
                eri_shell_quartet_pd(b, ab, cd, s1, s2, s3, s4, buf);
                nq++; npq += (long)ab->npp * cd->npp;
                double w = (s1 == s2 ? 1.0 : 2.0) * (s3 == s4 ? 1.0 : 2.0) * ((s1 == s3 && s2 == s4) ? 1.0 : 2.0);
                int f = 0;
                for (int f1 = 0; f1 < n1; f1++) for (int f2 = 0; f2 < n2f; f2++)
                    for (int f3 = 0; f3 < n3; f3++) for (int f4 = 0; f4 < n4; f4++, f++) {
                        int i = s2bf[s1] + f1, j = s2bf[s2] + f2, k = s2bf[s3] + f3, l = s2bf[s4] + f4;
                        double v = buf[f] * w;
                        M(rawJ, i, j) += M(Dtot, k, l) * v;
                        M(rawJ, k, l) += M(Dtot, i, j) * v;
                        if (kscale > 0.)
                            for (int x = 0; x < 3; x++) {
                                if (!DX[x]) continue;
                                double* rK = rawJ + (x + 1) * n2;
                                const double* D = DX[x];
                                M(rK, i, k) += M(D, j, l) * v; M(rK, j, l) += M(D, i, k) * v;
                                M(rK, i, l) += M(D, j, k) * v; M(rK, j, k) += M(D, i, l) * v;
                            }
                    }
            }
        free(buf);
    }
    for (int t = 1; t < nthreads; t++)
        for (size_t i = 0; i < 4 * n2; i++) raw[i] += raw[(size_t)t * 4 * n2 + i];
    double* outs[4] = {J, Kd, Ka, Kb};
    double sc[4] = {0.25, 0.125 * kscale, 0.125 * kscale, 0.125 * kscale};
    for (int m = 0; m < 4; m++) {
        if (!outs[m]) continue;
        const double* r = raw + m * n2;
        for (int i = 0; i < nbf; i++)
            for (int j = 0; j < nbf; j++) M(outs[m], i, j) = sc[m] * (M(r, i, j) + M(r, j, i));
    }
    if (counts) { counts[0] = nq; counts[1] = npq; }
    free(raw); free(Dtot); free_all_pairs(b, pd); free(s2bf);
}

/* Exact J and K blocks for ONE shell pair (sa,sb) of a large system (sampled parity at full size):
 *   Jblk[i][j] = sum_kl (ij|kl) Dtot_kl      i in sa, j in sb     (row-major na x nb)
 *   Kblk[i][k] = sum_jl (ij|kl) D_jl         i in sa, k in sb     (row-major na x nb), unscaled by exx
 * brute force over all shells, no symmetry. */
void oracle_jk_block(const cf_basis* b, int nbf, const double* Dtot, const double* D, int sa, int sb,
                     double* Jblk, double* Kblk, int nthreads) {
    int ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
    int na = sh_nfun(b, sa), nb = sh_nfun(b, sb);
    memset(Jblk, 0, sizeof(double) * na * nb);
    memset(Kblk, 0, sizeof(double) * na * nb);
    if (nthreads < 1) nthreads = 1;
    double* part = (double*)calloc((size_t)2 * na * nb * nthreads, sizeof(double));
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int s3 = 0; s3 < ns; s3++) {
#ifdef _OPENMP
        int ith = omp_get_thread_num();
#else
        int ith = 0;
#endif
        double* pj = part + (size_t)ith * 2 * na * nb;
        double* pk = pj + na * nb;
        double* buf = (double*)malloc(sizeof(double) * (2 * LMAX + 1) * (2 * LMAX + 1) * (2 * LMAX + 1) * (2 * LMAX + 1));
        for (int s4 = 0; s4 < ns; s4++) {
            int n3 = sh_nfun(b, s3), n4 = sh_nfun(b, s4);
            /* J: (sa sb | s3 s4) */
            oracle_eri_shell_quartet(b, sa, sb, s3, s4, buf);
            for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++)
                for (int k = 0; k < n3; k++) for (int l = 0; l < n4; l++)
                    pj[i * nb + j] += buf[((i * nb + j) * n3 + k) * n4 + l] * M(Dtot, s2bf[s3] + k, s2bf[s4] + l);
            /* K: (sa s3 | sb s4) D_{s3,s4} */
            if (D) {
                oracle_eri_shell_quartet(b, sa, s3, sb, s4, buf);
                for (int i = 0; i < na; i++) for (int j = 0; j < n3; j++)
                    for (int k = 0; k < nb; k++) for (int l = 0; l < n4; l++)
                        pk[i * nb + k] += buf[((i * n3 + j) * nb + k) * n4 + l] * M(D, s2bf[s3] + j, s2bf[s4] + l);
            }
        }
        free(buf);
    }
    for (int t = 0; t < nthreads; t++)
        for (int i = 0; i < na * nb; i++) { Jblk[i] += part[(size_t)t * 2 * na * nb + i]; Kblk[i] += part[(size_t)t * 2 * na * nb + na * nb + i]; }
    free(part); free(s2bf);
}

/* ================================================================ first-derivative ERIs
 * The reference asks libint2 for deriv_order = 1 (Int4C2E.cpp:339) and reads 12 buffers: d/dR of the integral for the
 * centres of s1..s4 x (x,y,z) (:377-389).  Restated here from first principles: for a primitive Cartesian Gaussian
 * d/dA_x [ (x-A)^l e^{-a r_A^2} ] = 2a (x-A)^{l+1} e^{..} - l (x-A)^{l-1} e^{..}, so a derivative block is the
 * (l+1) block with primitive prefactors scaled by 2a minus l_x times the (l-1) block; the pure transformation of the
 * ORIGINAL shell is applied afterwards (it is linear and independent of the centre). */
static void make_pair_shift(const cf_basis* b, int sa, int sb, int da, int db, int ea, int eb, pairdata* pd) {
    make_pair(b, sa, sb, pd);
    free(pd->E);
    int la = sh_l(b, sa) + da, lb = sh_l(b, sb) + db;
    int na = b->nprim[sa], nb = b->nprim[sb];
    const double* A = b->center_xyz + 3 * sa;
    const double* B = b->center_xyz + 3 * sb;
    pd->la = la; pd->lb = lb;
    pd->E = (ecoef*)malloc(sizeof(ecoef) * 3 * pd->npp);
    int n = 0;
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++, n++) {
            double a = b->exps[b->prim_offset[sa] + i], bb = b->exps[b->prim_offset[sb] + j];
            for (int k = 0; k < ea; k++) pd->K[n] *= 2.0 * a;     /* ea, eb: powers of 2*alpha (0..2) */
            for (int k = 0; k < eb; k++) pd->K[n] *= 2.0 * bb;
            for (int x = 0; x < 3; x++) hermite_E(la, lb, pd->p[n], pd->P[3 * n + x] - A[x], pd->P[3 * n + x] - B[x], &pd->E[3 * n + x]);
        }
}

/* buf12[(pos*3 + dir)][n1][n2][n3][n4]: d (s1 s2|s3 s4) / d R_pos,dir in the reference's function order */
void oracle_eri_deriv_quartet(const cf_basis* b, int s1, int s2, int s3, int s4, double* buf12) {
    int sh[4] = {s1, s2, s3, s4};
    int l[4], nc[4], nf[4];
    for (int i = 0; i < 4; i++) { l[i] = sh_l(b, sh[i]); nc[i] = NCART(l[i]); nf[i] = sh_nfun(b, sh[i]); }
    size_t ncart = (size_t)nc[0] * nc[1] * nc[2] * nc[3], nfun = (size_t)nf[0] * nf[1] * nf[2] * nf[3];
    pairdata ab0, cd0;
    make_pair(b, s1, s2, &ab0); make_pair(b, s3, s4, &cd0);
    double* dcart = (double*)malloc(sizeof(double) * ncart);
    double* t1 = (double*)malloc(sizeof(double) * ncart);
    double* t2 = (double*)malloc(sizeof(double) * ncart);
    double C[(2 * LMAX + 1) * NCMAX];
    for (int pos = 0; pos < 4; pos++) {
        int lp = l[pos];
        int ncp = NCART(lp + 1), ncm = lp > 0 ? NCART(lp - 1) : 0;
        int nn[4] = {nc[0], nc[1], nc[2], nc[3]};
        nn[pos] = ncp;
        size_t nplus = (size_t)nn[0] * nn[1] * nn[2] * nn[3];
        nn[pos] = ncm;
        size_t nminus = (size_t)nn[0] * nn[1] * nn[2] * nn[3];
        double* plus = (double*)malloc(sizeof(double) * nplus);
        double* minus = ncm ? (double*)malloc(sizeof(double) * nminus) : NULL;
        pairdata sp, sm;
        int da = (pos == 0 || pos == 2) ? 1 : 0, db = 1 - da;
        if (pos < 2) { make_pair_shift(b, s1, s2, da, db, da, db, &sp); eri_cart(&sp, &cd0, plus); }
        else { make_pair_shift(b, s3, s4, da, db, da, db, &sp); eri_cart(&ab0, &sp, plus); }
        free_pair(&sp);
        if (ncm) {
            if (pos < 2) { make_pair_shift(b, s1, s2, -da, -db, 0, 0, &sm); eri_cart(&sm, &cd0, minus); }
            else { make_pair_shift(b, s3, s4, -da, -db, 0, 0, &sm); eri_cart(&ab0, &sm, minus); }
            free_pair(&sm);
        }
        int comp[NCMAX][3];
        cart_components(lp, comp);
        /* strides of the 4-index Cartesian tensors around index `pos` */
        size_t outer = 1, inner = 1;
        for (int i = 0; i < pos; i++) outer *= nc[i];
        for (int i = pos + 1; i < 4; i++) inner *= nc[i];
        for (int dir = 0; dir < 3; dir++) {
            for (size_t o = 0; o < outer; o++)
                for (int c = 0; c < nc[pos]; c++) {
                    int e[3] = {comp[c][0], comp[c][1], comp[c][2]};
                    e[dir] += 1;
                    int cp = cart_index(lp + 1, e[0], e[1]);
                    int cm = -1;
                    if (comp[c][dir] > 0) { e[dir] -= 2; cm = cart_index(lp - 1, e[0], e[1]); }
                    for (size_t k = 0; k < inner; k++) {
                        double v = plus[(o * ncp + cp) * inner + k];
                        if (cm >= 0) v -= comp[c][dir] * minus[(o * ncm + cm) * inner + k];
                        dcart[(o * nc[pos] + c) * inner + k] = v;
                    }
                }
            int n[4] = {nc[0], nc[1], nc[2], nc[3]};
            memcpy(t1, dcart, sizeof(double) * ncart);
            double* src = t1; double* dst = t2;
            for (int i = 0; i < 4; i++) {
                int nnew = shell_transform(b->type[sh[i]], C);
                transform_index(src, dst, n, i, C, nnew);
                n[i] = nnew;
                double* tmp = src; src = dst; dst = tmp;
            }
            memcpy(buf12 + (size_t)(pos * 3 + dir) * nfun, src, sizeof(double) * nfun);
        }
        free(plus); if (minus) free(minus);
    }
    free(dcart); free(t1); free(t2);
    free_pair(&ab0); free_pair(&cd0);
}

/* getRepulsion1 (Int4C2E.cpp:312-408) restated: loop nest of the reference's quartet list (:83-93), function-level
 * uniqueness predicate and abcd_deg (:377-383), the six scatter updates per atom and direction (:388-397), then
 * gs = 1/2 (rawj + rawj^T) - 1/2 kscale * 1/4 (rawk + rawk^T) (:403-405).  G: [3*natom][nbf*nbf] col-major. */
void ref_getRepulsion1(const cf_basis* b, int natom, const double* D, double kscale, double* G, int nthreads) {
    int nbf = oracle_nbf(b), ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
    size_t n2 = (size_t)nbf * nbf, nmat = (size_t)3 * natom;
    if (nthreads < 1) nthreads = 1;
    double* raw = (double*)calloc(2 * nmat * n2 * nthreads, sizeof(double));   /* per thread: rawjs | rawks */
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int s1 = ns - 1; s1 >= 0; s1--) {
#ifdef _OPENMP
        int ith = omp_get_thread_num();
#else
        int ith = 0;
#endif
        double* rawj = raw + (size_t)ith * 2 * nmat * n2;
        double* rawk = rawj + nmat * n2;
        double* buf = (double*)malloc(sizeof(double) * 12 * (size_t)(2 * LMAX + 1) * (2 * LMAX + 1) * (2 * LMAX + 1) * (2 * LMAX + 1));
        short bf1_first = s2bf[s1], n1 = sh_nfun(b, s1);
        for (short s2 = 0; s2 <= s1; s2++) {
            short bf2_first = s2bf[s2], n2f = sh_nfun(b, s2);
            for (short s3 = 0; s3 <= s1; s3++) {
                short bf3_first = s2bf[s3], n3 = sh_nfun(b, s3);
                for (short s4 = 0; s4 <= (s2 > s3 ? s2 : s3); s4++) {
                    short bf4_first = s2bf[s4], n4 = sh_nfun(b, s4);
                    size_t nfun = (size_t)n1 * n2f * n3 * n4;
                    oracle_eri_deriv_quartet(b, s1, s2, s3, s4, buf);
                    const int atomlist[4] = {b->shell2atom[s1], b->shell2atom[s2], b->shell2atom[s3], b->shell2atom[s4]};
                    int f1234 = 0;
                    for (short f1 = 0; f1 != n1; f1++) {
                        const short bf1 = bf1_first + f1;
                        for (short f2 = 0; f2 != n2f; f2++) {
                            const short bf2 = bf2_first + f2;
                            const double ab_deg = (bf1 == bf2) ? 1 : 2;
                            for (short f3 = 0; f3 != n3; f3++) {
                                const short bf3 = bf3_first + f3;
                                for (short f4 = 0; f4 != n4; f4++, f1234++) {
                                    const short bf4 = bf4_first + f4;
                                    if (bf2 <= bf1 && bf3 <= bf1 && bf4 <= ((bf1 == bf3) ? bf2 : bf3)) {
                                        const double cd_deg = (bf3 == bf4) ? 1 : 2;
                                        const double ab_cd_deg = (bf1 == bf3) ? (bf2 == bf4 ? 1 : 2) : 2;
                                        const double abcd_deg = ab_deg * cd_deg * ab_cd_deg;
                                        for (int p = 0, pt = 0; p < 4; p++) {
                                            const int atom = atomlist[p];
                                            for (int t = 0; t < 3; t++, pt++) {
                                                const double tmp = abcd_deg * buf[(size_t)pt * nfun + f1234];
                                                double* rj = rawj + (size_t)(atom * 3 + t) * n2;
                                                double* rk = rawk + (size_t)(atom * 3 + t) * n2;
                                                M(rj, bf1, bf2) += tmp * M(D, bf3, bf4);
                                                M(rj, bf3, bf4) += tmp * M(D, bf1, bf2);
                                                if (kscale > 0) {
                                                    M(rk, bf1, bf3) += tmp * M(D, bf2, bf4);
                                                    M(rk, bf2, bf4) += tmp * M(D, bf1, bf3);
                                                    M(rk, bf1, bf4) += tmp * M(D, bf2, bf3);
                                                    M(rk, bf2, bf3) += tmp * M(D, bf1, bf4);
                                                }
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        free(buf);
    }
    for (int t = 1; t < nthreads; t++)
        for (size_t i = 0; i < 2 * nmat * n2; i++) raw[i] += raw[(size_t)t * 2 * nmat * n2 + i];
    for (size_t m = 0; m < nmat; m++) {
        const double* rj = raw + m * n2;
        const double* rk = raw + (nmat + m) * n2;
        double* g = G + m * n2;
        for (int i = 0; i < nbf; i++)
            for (int j = 0; j < nbf; j++)
                M(g, i, j) = 0.5 * (M(rj, i, j) + M(rj, j, i)) - 0.5 * kscale * 0.25 * (M(rk, i, j) + M(rk, j, i));
    }
    free(raw); free(s2bf);
}

/* Int4C2E::ContractGrads(D1, D2) (Int4C2E.cpp:747-763): grad[j] = sum D1 o G_j[D2], j = 3*atom + xyz */
void ref_ContractGrads(const cf_basis* b, int natom, const double* D1, const double* D2, double kscale, double* grad, int nthreads) {
    int nbf = oracle_nbf(b);
    size_t n2 = (size_t)nbf * nbf;
    double* G = (double*)malloc(sizeof(double) * 3 * natom * n2);
    ref_getRepulsion1(b, natom, D2, kscale, G, nthreads);
    for (int j = 0; j < 3 * natom; j++) {
        double s = 0;
        for (size_t i = 0; i < n2; i++) s += D1[i] * G[(size_t)j * n2 + i];
        grad[j] = s;
    }
    free(G);
}

/* ================================================================ second-derivative ERIs (tail)
 * The reference asks libint2 for deriv_order = 2 (Int4C2E.cpp:432) and reads 78 buffers: the upper triangle of the
 * 12 x 12 matrix d^2 (s1 s2|s3 s4) / dR_{p,t} dR_{q,s}, rows/columns ordered (centre of s1..s4) x (x,y,z), row-major with
 * (q,s) >= (p,t) (:468-472).  Restated from first principles like the first derivatives: each application of d/dR_{p,t}
 * to a primitive Cartesian Gaussian gives 2a * (component raised by one in t) - n_t * (component lowered by one), so a
 * second derivative is a combination of <= 4 Cartesian blocks with angular momenta shifted by -2..+2 and primitive
 * prefactors scaled by (2a)^0..2.  Blocks are computed on demand and cached for the quartet; the pure transformation of
 * the ORIGINAL shells is applied at the end. */
typedef struct { double c; int n[4][3]; int e[4]; } dterm;

static int apply_deriv(const dterm* in, int nin, int pos, int dir, dterm* out) {
    int nout = 0;
    for (int i = 0; i < nin; i++) {
        dterm up = in[i];
        up.n[pos][dir] += 1; up.e[pos] += 1;
        out[nout++] = up;
        if (in[i].n[pos][dir] > 0) {
            dterm dn = in[i];
            dn.c *= -(double)in[i].n[pos][dir];
            dn.n[pos][dir] -= 1;
            out[nout++] = dn;
        }
    }
    return nout;
}

typedef struct { int key; int l[4]; int nc[4]; double* v; } d2block;
typedef struct { int key; pairdata pd; } d2pair;

static pairdata* d2_get_pair(const cf_basis* b, d2pair* cache, int* ncache, int sa, int sb, int da, int db, int ea, int eb) {
    int key = (((da + 2) * 5 + (db + 2)) * 3 + ea) * 3 + eb;
    for (int i = 0; i < *ncache; i++) if (cache[i].key == key) return &cache[i].pd;
    cache[*ncache].key = key;
    make_pair_shift(b, sa, sb, da, db, ea, eb, &cache[*ncache].pd);
    return &cache[(*ncache)++].pd;
}

void oracle_eri_deriv2_quartet(const cf_basis* b, int s1, int s2, int s3, int s4, double* buf78) {
    int sh[4] = {s1, s2, s3, s4};
    int l[4], nc[4], nf[4];
    for (int i = 0; i < 4; i++) { l[i] = sh_l(b, sh[i]); nc[i] = NCART(l[i]); nf[i] = sh_nfun(b, sh[i]); }
    size_t ncart = (size_t)nc[0] * nc[1] * nc[2] * nc[3], nfun = (size_t)nf[0] * nf[1] * nf[2] * nf[3];
    d2pair bra[64], ket[64];
    int nbra = 0, nket = 0;
    d2block blk[128];
    int nblk = 0;
    double* dcart = (double*)malloc(sizeof(double) * ncart);
    double* t1 = (double*)malloc(sizeof(double) * ncart);
    double* t2 = (double*)malloc(sizeof(double) * ncart);
    double C[(2 * LMAX + 1) * NCMAX];
    int comp[4][NCMAX][3];
    for (int i = 0; i < 4; i++) cart_components(l[i], comp[i]);
    int ptqs = 0;
    for (int p = 0; p < 4; p++) for (int t = 0; t < 3; t++)
        for (int q = p; q < 4; q++) for (int s = (q == p ? t : 0); s < 3; s++, ptqs++) {
            size_t n = 0;
            for (int ia = 0; ia < nc[0]; ia++) for (int ib = 0; ib < nc[1]; ib++)
                for (int ic = 0; ic < nc[2]; ic++) for (int id = 0; id < nc[3]; id++, n++) {
                    dterm a0, a1[2], a2[4];
                    a0.c = 1.0;
                    int idx[4] = {ia, ib, ic, id};
                    for (int i = 0; i < 4; i++) { a0.e[i] = 0; for (int x = 0; x < 3; x++) a0.n[i][x] = comp[i][idx[i]][x]; }
                    int n1 = apply_deriv(&a0, 1, p, t, a1);
                    int n2 = apply_deriv(a1, n1, q, s, a2);
                    double v = 0.0;
                    for (int k = 0; k < n2; k++) {
                        int ll[4], key = 0;
                        for (int i = 0; i < 4; i++) {
                            ll[i] = a2[k].n[i][0] + a2[k].n[i][1] + a2[k].n[i][2];
                            key = (key * 5 + (ll[i] - l[i] + 2)) * 3 + a2[k].e[i];
                        }
                        d2block* B = NULL;
                        for (int i = 0; i < nblk; i++) if (blk[i].key == key) { B = &blk[i]; break; }
                        if (!B) {
                            B = &blk[nblk++];
                            B->key = key;
                            size_t sz = 1;
                            for (int i = 0; i < 4; i++) { B->l[i] = ll[i]; B->nc[i] = NCART(ll[i]); sz *= B->nc[i]; }
                            B->v = (double*)malloc(sizeof(double) * sz);
                            pairdata* ab = d2_get_pair(b, bra, &nbra, s1, s2, ll[0] - l[0], ll[1] - l[1], a2[k].e[0], a2[k].e[1]);
                            pairdata* cd = d2_get_pair(b, ket, &nket, s3, s4, ll[2] - l[2], ll[3] - l[3], a2[k].e[2], a2[k].e[3]);
                            eri_cart(ab, cd, B->v);
                        }
                        size_t off = 0;
                        for (int i = 0; i < 4; i++) off = off * B->nc[i] + cart_index(B->l[i], a2[k].n[i][0], a2[k].n[i][1]);
                        v += a2[k].c * B->v[off];
                    }
                    dcart[n] = v;
                }
            int nn[4] = {nc[0], nc[1], nc[2], nc[3]};
            memcpy(t1, dcart, sizeof(double) * ncart);
            double* src = t1; double* dst = t2;
            for (int i = 0; i < 4; i++) {
                int nnew = shell_transform(b->type[sh[i]], C);
                transform_index(src, dst, nn, i, C, nnew);
                nn[i] = nnew;
                double* tmp = src; src = dst; dst = tmp;
            }
            memcpy(buf78 + (size_t)ptqs * nfun, src, sizeof(double) * nfun);
        }
    for (int i = 0; i < nblk; i++) free(blk[i].v);
    for (int i = 0; i < nbra; i++) free_pair(&bra[i].pd);
    for (int i = 0; i < nket; i++) free_pair(&ket[i].pd);
    free(dcart); free(t1); free(t2);
}

/* getRepulsion2 (Int4C2E.cpp:410-492) restated: the reference's quartet list (:83-93), uniqueness predicate and abcd_deg
 * (:461-466), the 78 buffers scattered to (3*atom[p]+t, 3*atom[q]+s) with scale 2 where two DIFFERENT positions land on the
 * same nuclear coordinate (:471-477), hessianj *= 2, raw = hessianj - 1/2 kscale hessiank, H = raw + raw^T - diag(raw)
 * (:487-491).  H: [3*natom][3*natom] col-major (symmetric). */
void ref_getRepulsion2(const cf_basis* b, int natom, const double* D, double kscale, double* H, int nthreads) {
    int nbf = oracle_nbf(b), ns = b->nshell;
    int* s2bf = (int*)malloc(sizeof(int) * ns);
    shell2bf(b, s2bf);
    int nh = 3 * natom;
    size_t nh2 = (size_t)nh * nh;
    if (nthreads < 1) nthreads = 1;
    double* raw = (double*)calloc(2 * nh2 * nthreads, sizeof(double));   /* per thread: hessianj | hessiank */
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int s1 = ns - 1; s1 >= 0; s1--) {
#ifdef _OPENMP
        int ith = omp_get_thread_num();
#else
        int ith = 0;
#endif
        double* hj = raw + (size_t)ith * 2 * nh2;
        double* hk = hj + nh2;
        double* buf = (double*)malloc(sizeof(double) * 78 * (size_t)(2 * LMAX + 1) * (2 * LMAX + 1) * (2 * LMAX + 1) * (2 * LMAX + 1));
        short bf1_first = s2bf[s1], n1 = sh_nfun(b, s1);
        for (short s2 = 0; s2 <= s1; s2++) {
            short bf2_first = s2bf[s2], n2f = sh_nfun(b, s2);
            for (short s3 = 0; s3 <= s1; s3++) {
                short bf3_first = s2bf[s3], n3 = sh_nfun(b, s3);
                for (short s4 = 0; s4 <= (s2 > s3 ? s2 : s3); s4++) {
                    short bf4_first = s2bf[s4], n4 = sh_nfun(b, s4);
                    size_t nfun = (size_t)n1 * n2f * n3 * n4;
                    oracle_eri_deriv2_quartet(b, s1, s2, s3, s4, buf);
                    const int atomlist[4] = {b->shell2atom[s1], b->shell2atom[s2], b->shell2atom[s3], b->shell2atom[s4]};
                    int f1234 = 0;
                    for (short f1 = 0; f1 != n1; f1++) {
                        const short bf1 = bf1_first + f1;
                        for (short f2 = 0; f2 != n2f; f2++) {
                            const short bf2 = bf2_first + f2;
                            const double ab_deg = (bf1 == bf2) ? 1 : 2;
                            for (short f3 = 0; f3 != n3; f3++) {
                                const short bf3 = bf3_first + f3;
                                for (short f4 = 0; f4 != n4; f4++, f1234++) {
                                    const short bf4 = bf4_first + f4;
                                    if (bf2 <= bf1 && bf3 <= bf1 && bf4 <= ((bf1 == bf3) ? bf2 : bf3)) {
                                        const double cd_deg = (bf3 == bf4) ? 1 : 2;
                                        const double ab_cd_deg = (bf1 == bf3) ? (bf2 == bf4 ? 1 : 2) : 2;
                                        const double abcd_deg = ab_deg * cd_deg * ab_cd_deg;
                                        const double dj = M(D, bf1, bf2) * M(D, bf3, bf4);
                                        const double dk = M(D, bf1, bf3) * M(D, bf2, bf4) + M(D, bf1, bf4) * M(D, bf2, bf3);
                                        for (int p = 0, ptqs = 0; p < 4; p++) for (int t = 0; t < 3; t++) {
                                            const int xpert = 3 * atomlist[p] + t;
                                            for (int q = p; q < 4; q++) for (int s = (q == p ? t : 0); s < 3; s++, ptqs++) {
                                                const int ypert = 3 * atomlist[q] + s;
                                                const double scale = (xpert == ypert && p != q) ? 2 : 1;
                                                const double tmp = scale * abcd_deg * buf[(size_t)ptqs * nfun + f1234];
                                                hj[(size_t)ypert * nh + xpert] += tmp * dj;
                                                if (kscale > 0) hk[(size_t)ypert * nh + xpert] += tmp * dk;
                                            }
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        free(buf);
    }
    for (int t = 1; t < nthreads; t++)
        for (size_t i = 0; i < 2 * nh2; i++) raw[i] += raw[(size_t)t * 2 * nh2 + i];
    const double* hj = raw;
    const double* hk = raw + nh2;
    for (int x = 0; x < nh; x++)
        for (int y = 0; y < nh; y++) {
            const double rxy = 2.0 * hj[(size_t)y * nh + x] - 0.5 * kscale * hk[(size_t)y * nh + x];
            const double ryx = 2.0 * hj[(size_t)x * nh + y] - 0.5 * kscale * hk[(size_t)x * nh + y];
            H[(size_t)y * nh + x] = (x == y) ? rxy : rxy + ryx;
        }
    free(raw); free(s2bf);
}

/* The whole reference path in one call (setup :49-53 of SelfConsistentField.cpp + ContractInts).
 * counts: [0] RepulsionLength, [1] ShellQuartetLength. Returns 0. */
int ref_full_path(const cf_basis* b, double threshold, double exx, int nthreads,
                  const double* Dd, const double* Da, const double* Db,
                  double* J, double* Kd, double* Ka, double* Kb, long* counts) {
    int nbf = oracle_nbf(b);
    double* diag = (double*)malloc(sizeof(double) * (size_t)nbf * nbf);
    ref_getRepulsionDiag(b, diag);
    long nint, nsq;
    ref_getRepulsionLengthIndices(b, diag, threshold, &nint, &nsq, NULL, NULL, NULL, NULL);
    short* sh = (short*)malloc(sizeof(short) * 4 * (nsq + 1));
    long dummy1, dummy2;
    ref_getRepulsionLengthIndices(b, diag, threshold, &dummy1, &dummy2, sh, sh + nsq, sh + 2 * nsq, sh + 3 * nsq);
    short* bf = (short*)malloc(sizeof(short) * 4 * (nint + 1));
    double* ints = (double*)calloc(nint + 1, sizeof(double));
    ref_getRepulsion0(b, nsq, sh, sh + nsq, sh + 2 * nsq, sh + 3 * nsq, bf, bf + nint, bf + 2 * nint, bf + 3 * nint, ints);
    ref_Gunified(bf, bf + nint, bf + 2 * nint, bf + 3 * nint, ints, nint, nbf, Dd, Da, Db, exx, nthreads, J, Kd, Ka, Kb);
    if (counts) { counts[0] = nint; counts[1] = nsq; }
    free(diag); free(sh); free(bf); free(ints);
    return 0;
}

/* Stored-integral handle for timing B1 (the reference's real per-iteration cost): build once, contract many. */
typedef struct { long nint; short* bf; double* ints; int nbf; } ref_store;
ref_store* ref_store_build(const cf_basis* b, double threshold) {
    ref_store* st = (ref_store*)calloc(1, sizeof(ref_store));
    int nbf = oracle_nbf(b);
    double* diag = (double*)malloc(sizeof(double) * (size_t)nbf * nbf);
    ref_getRepulsionDiag(b, diag);
    long nint, nsq, d1, d2;
    ref_getRepulsionLengthIndices(b, diag, threshold, &nint, &nsq, NULL, NULL, NULL, NULL);
    short* sh = (short*)malloc(sizeof(short) * 4 * (nsq + 1));
    ref_getRepulsionLengthIndices(b, diag, threshold, &d1, &d2, sh, sh + nsq, sh + 2 * nsq, sh + 3 * nsq);
    st->bf = (short*)malloc(sizeof(short) * 4 * (nint + 1));
    st->ints = (double*)calloc(nint + 1, sizeof(double));
    st->nint = nint; st->nbf = nbf;
    ref_getRepulsion0(b, nsq, sh, sh + nsq, sh + 2 * nsq, sh + 3 * nsq, st->bf, st->bf + nint, st->bf + 2 * nint, st->bf + 3 * nint, st->ints);
    free(diag); free(sh);
    return st;
}
long ref_store_len(const ref_store* st) { return st->nint; }
void ref_store_contract(const ref_store* st, const double* Dd, const double* Da, const double* Db, double exx, int nthreads,
                        double* J, double* Kd, double* Ka, double* Kb) {
    long n = st->nint;
    ref_Gunified(st->bf, st->bf + n, st->bf + 2 * n, st->bf + 3 * n, st->ints, n, st->nbf, Dd, Da, Db, exx, nthreads, J, Kd, Ka, Kb);
}
void ref_store_free(ref_store* st) { free(st->bf); free(st->ints); free(st); }
/* set the OpenMP thread count explicitly (launchers such as torchrun export OMP_NUM_THREADS=1); n <= 0: all online cores */
int oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
