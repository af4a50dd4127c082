// oneint.cuh -- one-electron integrals S, T, V on the device.
//
// Replaces getTwoCenter0 with Operator::overlap / kinetic / nuclear (src/Integral/Int2C1E.cpp:18-67) as called by
// Int2C1E::CalculateIntegrals(0, ...) (:313-333; the multipole matrices and the ECP term of that function are outside
// this engine's scope: none of the BASELINE configurations carries an ECP).
// One CTA per shell pair (s1 >= s2).  Overlap and kinetic energy: Obara-Saika 1-D overlap tables per primitive pair;
// nuclear attraction: Rys quadrature (the ERI recurrences in the limit of an infinitely tight ket Gaussian):
//     V_ab = -sum_C Z_C (2 pi / p) K_ab sum_r w_r(T) prod_d I_d(t_r^2),  T = p |P - C|^2,
//     I_{i+1} = (PA_d - t^2 PC_d) I_i + i (1 - t^2) / (2p) I_{i-1},  then the a -> b transfer with AB_d.
// Threads of the CTA share the atoms; per-thread partial blocks are combined in a FIXED order (deterministic).
#pragma once
#include "cf_common.cuh"
#include "eri_generic.cuh"

#define ONEINT_THREADS 64
#define ONEINT_LMAX CF_LMAX_DEV
#define ONEINT_NC ((ONEINT_LMAX + 1) * (ONEINT_LMAX + 2) / 2)

struct OneIntTask {
    int ns, nbf, natom;
    const int* l; const int* nprim; const int* prim_off;    // per shell
    const double* exps; const double* coefs; const double* xyz;
    const double* Z; const double* atom_xyz;                  // [natom], [3 natom]
    const double* ctrans; const int* ct_off; const int* bf_off; const int* nfun;
    RysTablesDev rys;
    double* S; double* T; double* V;                          // nbf x nbf col-major (symmetric)
};

template <int NR>
__device__ __forceinline__ void oneint_roots(const RysTablesDev& rys, double T, double* x, double* w) {
#pragma unroll
    for (int r = 0; r < NR; r++) { x[r] = rys_value<NR>(rys, T, r); w[r] = rys_value<NR>(rys, T, NR + r); }
}

__global__ void __launch_bounds__(ONEINT_THREADS) oneint_kernel(const OneIntTask t) {
    // shell pair of this CTA
    const long long e = blockIdx.x;
    long long s1 = (long long)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while (s1 * (s1 + 1) / 2 > e) s1--;
    while ((s1 + 1) * (s1 + 2) / 2 <= e) s1++;
    const int sa = (int)s1, sb = (int)(e - s1 * (s1 + 1) / 2);
    const int la = t.l[sa], lb = t.l[sb], nca = cf_ncart(la), ncb = cf_ncart(lb), nab = nca * ncb, lab = la + lb;
    const int nr = lab / 2 + 1;
    const double Ax = t.xyz[3 * sa], Ay = t.xyz[3 * sa + 1], Az = t.xyz[3 * sa + 2];
    const double Bx = t.xyz[3 * sb], By = t.xyz[3 * sb + 1], Bz = t.xyz[3 * sb + 2];
    const double AB[3] = {Ax - Bx, Ay - By, Az - Bz};
    const double r2 = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
    const int tid = threadIdx.x;

    __shared__ double sS[ONEINT_NC * ONEINT_NC], sT[ONEINT_NC * ONEINT_NC], sV[ONEINT_NC * ONEINT_NC];
    __shared__ double red[ONEINT_THREADS / 32];
    // Cartesian exponents of the components
    __shared__ signed char ca[ONEINT_NC][3], cb[ONEINT_NC][3];
    if (tid < nca) { int x, y, z; cart_comp(la, tid, x, y, z); ca[tid][0] = x; ca[tid][1] = y; ca[tid][2] = z; }
    if (tid < ncb) { int x, y, z; cart_comp(lb, tid, x, y, z); cb[tid][0] = x; cb[tid][1] = y; cb[tid][2] = z; }
    for (int i = tid; i < nab; i += ONEINT_THREADS) { sS[i] = 0.0; sT[i] = 0.0; sV[i] = 0.0; }
    __syncthreads();

    double vloc[ONEINT_NC * ONEINT_NC];      // this thread's nuclear-attraction partial block (its atoms, all primitives)
    for (int i = 0; i < nab; i++) vloc[i] = 0.0;

    for (int ia = 0; ia < t.nprim[sa]; ia++)
        for (int jb = 0; jb < t.nprim[sb]; jb++) {
            const double ea = t.exps[t.prim_off[sa] + ia], eb = t.exps[t.prim_off[sb] + jb], p = ea + eb, hp = 0.5 / p;
            const double K = t.coefs[t.prim_off[sa] + ia] * t.coefs[t.prim_off[sb] + jb] * exp(-ea * eb / p * r2);
            const double P[3] = {(ea * Ax + eb * Bx) / p, (ea * Ay + eb * By) / p, (ea * Az + eb * Bz) / p};
            const double PA[3] = {P[0] - Ax, P[1] - Ay, P[2] - Az}, PB[3] = {P[0] - Bx, P[1] - By, P[2] - Bz};
            // ---- overlap / kinetic: 1-D overlaps s[d][i][j], i <= la, j <= lb + 2
            double s[3][ONEINT_LMAX + 1][ONEINT_LMAX + 3];
            const double s00 = sqrt(M_PI / p);
            for (int d = 0; d < 3; d++)
                for (int i = 0; i <= la; i++)
                    for (int j = 0; j <= lb + 2; j++) {
                        double v;
                        if (i == 0 && j == 0) v = s00;
                        else if (j == 0) v = PA[d] * s[d][i - 1][0] + (i > 1 ? (i - 1) * hp * s[d][i - 2][0] : 0.0);
                        else v = PB[d] * s[d][i][j - 1] + (i > 0 ? i * hp * s[d][i - 1][j - 1] : 0.0) + (j > 1 ? (j - 1) * hp * s[d][i][j - 2] : 0.0);
                        s[d][i][j] = v;
                    }
            for (int idx = tid; idx < nab; idx += ONEINT_THREADS) {     // every component pair is owned by one thread
                const int xa = idx / ncb, xb = idx - xa * ncb;
                double s1d[3], k1d[3];
                for (int d = 0; d < 3; d++) {
                    const int i = ca[xa][d], j = cb[xb][d];
                    s1d[d] = s[d][i][j];
                    k1d[d] = -2.0 * eb * eb * s[d][i][j + 2] + eb * (2 * j + 1) * s[d][i][j] - (j > 1 ? 0.5 * j * (j - 1) * s[d][i][j - 2] : 0.0);
                }
                sS[idx] += K * s1d[0] * s1d[1] * s1d[2];
                sT[idx] += K * (k1d[0] * s1d[1] * s1d[2] + s1d[0] * k1d[1] * s1d[2] + s1d[0] * s1d[1] * k1d[2]);
            }
            // ---- nuclear attraction: atoms spread over the threads
            const double pref = -2.0 * M_PI / p * K;
            for (int c = tid; c < t.natom; c += ONEINT_THREADS) {
                const double PC[3] = {P[0] - t.atom_xyz[3 * c], P[1] - t.atom_xyz[3 * c + 1], P[2] - t.atom_xyz[3 * c + 2]};
                const double Tq = p * (PC[0] * PC[0] + PC[1] * PC[1] + PC[2] * PC[2]);
                double x[4], w[4];
                switch (nr) {
                    case 1: oneint_roots<1>(t.rys, Tq, x, w); break;
                    case 2: oneint_roots<2>(t.rys, Tq, x, w); break;
                    case 3: oneint_roots<3>(t.rys, Tq, x, w); break;
                    default: oneint_roots<4>(t.rys, Tq, x, w); break;
                }
                const double zc = pref * t.Z[c];
                for (int r = 0; r < nr; r++) {
                    double g[3][ONEINT_LMAX + 1][ONEINT_LMAX + 1];      // [d][i][j] = 1-D integral (i on a, j on b)
                    const double b10 = (1.0 - x[r]) * hp;
                    for (int d = 0; d < 3; d++) {
                        double h[2 * ONEINT_LMAX + 1];
                        const double c00 = PA[d] - x[r] * PC[d];
                        h[0] = (d == 2) ? zc * w[r] : 1.0;
                        if (lab > 0) h[1] = c00 * h[0];
                        for (int i = 1; i < lab; i++) h[i + 1] = c00 * h[i] + i * b10 * h[i - 1];
                        for (int i = 0; i <= la; i++) g[d][i][0] = h[i];
                        for (int j = 1; j <= lb; j++) {
                            for (int i = 0; i <= lab - j; i++) h[i] = h[i + 1] + AB[d] * h[i];
                            for (int i = 0; i <= la; i++) g[d][i][j] = h[i];
                        }
                    }
                    for (int idx = 0; idx < nab; idx++) {
                        const int xa = idx / ncb, xb = idx - xa * ncb;
                        vloc[idx] += g[0][ca[xa][0]][cb[xb][0]] * g[1][ca[xa][1]][cb[xb][1]] * g[2][ca[xa][2]][cb[xb][2]];
                    }
                }
            }
        }
    // ---- combine the per-thread nuclear blocks: butterfly inside the warp, then the warps in order (fixed order)
    for (int idx = 0; idx < nab; idx++) {
        double v = vloc[idx];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) { double sum = 0.0; for (int k = 0; k < ONEINT_THREADS / 32; k++) sum += red[k]; sV[idx] = sum; }
    }
    __syncthreads();
    // ---- Cartesian -> the reference's function order, both triangles
    const int na = t.nfun[sa], nb = t.nfun[sb];
    const double* Ca = t.ctrans + t.ct_off[sa];
    const double* Cb = t.ctrans + t.ct_off[sb];
    for (int o = tid; o < 3 * na * nb; o += ONEINT_THREADS) {
        const int m = o / (na * nb), rem = o - m * na * nb, i = rem / nb, j = rem - i * nb;
        if (sa == sb && j > i) continue;          // diagonal blocks: lower triangle mirrored, exactly symmetric (Int2C1E.cpp:52-59)
        const double* src = m == 0 ? sS : m == 1 ? sT : sV;
        double* dst = m == 0 ? t.S : m == 1 ? t.T : t.V;
        double sum = 0.0;
        for (int xa = 0; xa < nca; xa++) {
            const double cam = Ca[i * nca + xa];
            if (cam == 0.0) continue;
            for (int xb = 0; xb < ncb; xb++) sum = fma(cam * Cb[j * ncb + xb], src[xa * ncb + xb], sum);
        }
        const size_t bi = t.bf_off[sa] + i, bj = t.bf_off[sb] + j;
        dst[bj * t.nbf + bi] = sum;
        dst[bi * t.nbf + bj] = sum;
    }
}
