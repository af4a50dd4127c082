// eri_hess.cuh -- second-derivative ERIs contracted on the fly into the nuclear Hessian (tail).
//
// Replaces getRepulsion2 + Int4C2E::ContractHesss(D1, D2) (src/Integral/Int4C2E.cpp:410-492, :792-811; the reference
// uses D = D2 only, :793):
//     H[X,Y] = sum over canonical shell quartets  sum_abcd  d^2 (ab|cd) / dX dY  *  Gamma_abcd ,
//     Gamma_abcd = wgt * [ 2 D_ab D_cd - exx/2 (D_ac D_bd + D_ad D_bc) ]          (densities fixed)
// (the reference's degeneracy weights :461-466 folded into `wgt`; its `hessianj *= 2`, `- 0.5 kscale hessiank` and
// raw + raw^T - diag(raw), :487-491, are the chain rule over ALL ordered pairs of the four centres, which is what the
// full 12 x 12 block below adds).  The 78 buffers libint2 hands the reference are never formed.
// Rys quadrature: d/dA_x acts on the x-direction 2-D integral only,
//     d/dA_x I(i)        = ta I(i+1) - i I(i-1)                                   (ta = 2 alpha_a of the primitive)
//     d2/dA_x2 I(i)      = ta^2 I(i+2) - ta (2i+1) I(i) + i (i-1) I(i-2)
//     d2/dA_x dB_x I(i,j) = ta tb I(i+1,j+1) - ta j I(i+1,j-1) - i tb I(i-1,j+1) + i j I(i-1,j-1)
// so one 2-D table with the indices of a, b, c raised by two serves the 9 x 9 block of the centres A, B, C (45 unique
// entries, accumulated per lane); the rows and columns of centre D follow from translational invariance.
// One CTA of G threads per shell quartet (the layout of eri_grad_generic); the 12 x 12 block of a quartet is added to the
// 3 natom x 3 natom matrix with FP64 atomics (this call feeds `derivative 2` jobs, Restricted/Hess.cpp:67, not the SCF loop).
#pragma once
#include "eri_grad.cuh"

// roots per pass over the raised 2-D tables: all of them where the tables fit ~200 KB of shared memory next to Gamma,
// otherwise the smallest number of equal batches that does (ff|ff: 8 roots in two batches of 4)
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr int eri_hess_root_batch() {
    constexpr int NR = (LA + LB + LC + LD + 2) / 2 + 1;
    constexpr int GSZ = (LA + 3) * (LB + 3) * (LC + 3) * (LD + 1);
    constexpr int NOUT = cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD);
    int nb = NR;
    for (int parts = 1; parts <= NR; parts++) {
        nb = (NR + parts - 1) / parts;
        if (sizeof(double) * (size_t)(nb * 3 * GSZ + NOUT + 1024) <= 200 * 1024) break;
    }
    return nb;
}
template <int LA, int LB, int LC, int LD>
constexpr size_t eri_hess_smem(int G) {
    constexpr int NR = (LA + LB + LC + LD + 2) / 2 + 1;
    constexpr int NRB = eri_hess_root_batch<LA, LB, LC, LD>();
    constexpr int GSZ = (LA + 3) * (LB + 3) * (LC + 3) * (LD + 1);
    constexpr int NOUT = cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD);
    return sizeof(double) * (size_t)(2 * NR + NRB * 3 * GSZ + NOUT + (G / 32) * 45 + 144 + 16);
}

// position of (u, v), u <= v < 9, in the packed upper triangle
__host__ __device__ constexpr int hess_tri(int u, int v) { return u * 9 - u * (u - 1) / 2 + (v - u); }

template <int LA, int LB, int LC, int LD, int G>
__global__ void __launch_bounds__(G) eri_hess_generic(const GradTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NOUT = NA * NB * NC * ND;
    constexpr int NR = (LA + LB + LC + LD + 2) / 2 + 1;
    constexpr int NRB = eri_hess_root_batch<LA, LB, LC, LD>();
    constexpr int GSZ = (LA + 3) * (LB + 3) * (LC + 3) * (LD + 1);
    constexpr int SA = (LB + 3) * (LC + 3) * (LD + 1), SB = (LC + 3) * (LD + 1), SC = (LD + 1);
    extern __shared__ double smem[];
    double* rw = smem;
    double* g = rw + 2 * NR;
    double* gam = g + NRB * 3 * GSZ;        // [NOUT] effective two-particle density of the quartet
    double* red = gam + NOUT;               // [G/32][45]
    double* blk = red + (G / 32) * 45;      // [12][12] block of the quartet
    const int lane = threadIdx.x;
    const size_t ld = (size_t)t.ncart;
    const double* __restrict__ D = t.D1;

    const long long nchunk_total = (t.nquartet + t.chunk - 1) / t.chunk;
    const long long nchunk_local = (nchunk_total - t.rank + t.world - 1) / t.world;
    for (long long lc = blockIdx.x; lc < nchunk_local; lc += gridDim.x) {
        const long long chunk = lc * t.world + t.rank;
        long long q = chunk * t.chunk;
        const long long q_end = min(q + (long long)t.chunk, t.nquartet);
        int lo = 0, hi = t.bra.npair;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (t.qoff[mid] <= q) lo = mid; else hi = mid;
        }
        int ib = lo;
        for (; q < q_end; q++) {
            while (t.qoff[ib + 1] <= q) ib++;
            const int ik = (int)(q - t.qoff[ib]);
            if (t.thr > 0.0 && !(t.bra.Q[ib] * t.ket.Q[ik] > t.thr)) continue;   // uniform across the CTA (Int4C2E.cpp:108-113)
            const int sa = t.bra.sa[ib], sb = t.bra.sb[ib], sc = t.ket.sa[ik], sd = t.ket.sb[ik];
            const double Ax = t.bra.A[3 * ib], Ay = t.bra.A[3 * ib + 1], Az = t.bra.A[3 * ib + 2];
            const double ABx = t.bra.AB[3 * ib], ABy = t.bra.AB[3 * ib + 1], ABz = t.bra.AB[3 * ib + 2];
            const double Cx = t.ket.A[3 * ik], Cy = t.ket.A[3 * ik + 1], Cz = t.ket.A[3 * ik + 2];
            const double CDx = t.ket.AB[3 * ik], CDy = t.ket.AB[3 * ik + 1], CDz = t.ket.AB[3 * ik + 2];
            const int pab0 = t.bra.pbase[ib], npab = t.bra.nprim[ib];
            const int pcd0 = t.ket.pbase[ik], npcd = t.ket.nprim[ik];
            double wgt = (sa == sb ? 1.0 : 2.0) * (sc == sd ? 1.0 : 2.0);
            wgt *= (t.same_class && ib == ik) ? 1.0 : 2.0;
            const int ca = t.bra.cao_a[ib], cb = t.bra.cao_b[ib], cc = t.ket.cao_a[ik], cdd = t.ket.cao_b[ik];

            __syncthreads();                                       // gam / blk of the previous quartet consumed
            for (int n = lane; n < NOUT; n += G) {
                const int a = ca + n / (ND * NC * NB), b = cb + (n / (ND * NC)) % NB, c = cc + (n / ND) % NC, d = cdd + n % ND;
                double v = 2.0 * D[b * ld + a] * D[d * ld + c];
                if (t.exx > 0.0) v -= 0.5 * t.exx * (D[c * ld + a] * D[d * ld + b] + D[d * ld + a] * D[c * ld + b]);
                gam[n] = v * wgt;
            }
            double acc[45];
#pragma unroll
            for (int e = 0; e < 45; e++) acc[e] = 0.0;

            for (int iab = 0; iab < npab; iab++) {
                const int sab = pab0 + iab * CF_PSTRIDE;
                const double p = t.bra.p[sab], cab = t.bra.c[sab];
                const double Px = t.bra.Px[sab], Py = t.bra.Py[sab], Pz = t.bra.Pz[sab];
                const double ta = 2.0 * t.bra_aexp[sab], tb = 2.0 * p - ta;        // 2 alpha_a, 2 alpha_b
                for (int icd = 0; icd < npcd; icd++) {
                    const int scd = pcd0 + icd * CF_PSTRIDE;
                    const double ccd = t.ket.c[scd];
                    if (fabs(cab * ccd) < t.prim_cut) continue;   // uniform across the CTA
                    const double qe = t.ket.p[scd];
                    const double Qx = t.ket.Px[scd], Qy = t.ket.Py[scd], Qz = t.ket.Pz[scd];
                    const double tc = 2.0 * t.ket_aexp[scd];
                    const double pq = p + qe;
                    const double rho = p * qe / pq;
                    const double PQx = Px - Qx, PQy = Py - Qy, PQz = Pz - Qz;
                    const double T = rho * (PQx * PQx + PQy * PQy + PQz * PQz);
                    __syncthreads();                               // previous roots / tables consumed
                    if (lane < 2 * NR) rw[lane] = rys_value<NR>(t.rys, T, lane);
                    const double taa = ta * ta, tbb = tb * tb, tcc = tc * tc, tab = ta * tb, tac = ta * tc, tbc = tb * tc;
                    for (int r0 = 0; r0 < NR; r0 += NRB) {
                    __syncthreads();                               // roots visible / tables of the previous batch consumed
                    for (int tk = lane; tk < 3 * NRB; tk += G) {
                        const int r = r0 + tk / 3, dim = tk % 3;
                        if (r >= NR) continue;
                        const double x = rw[r];
                        const double rx_p = rho * x / p;
                        const double rx_q = rho * x / qe;
                        const double b00 = 0.5 * x / pq;
                        const double b10 = (1.0 - rx_p) * (0.5 / p);
                        const double b01 = (1.0 - rx_q) * (0.5 / qe);
                        double PA, PQ, QC, ab, cd, w0;
                        if (dim == 0) { PA = Px - Ax; PQ = PQx; QC = Qx - Cx; ab = ABx; cd = CDx; w0 = 1.0; }
                        else if (dim == 1) { PA = Py - Ay; PQ = PQy; QC = Qy - Cy; ab = ABy; cd = CDy; w0 = 1.0; }
                        else { PA = Pz - Az; PQ = PQz; QC = Qz - Cz; ab = ABz; cd = CDz; w0 = rw[NR + r] * cab * ccd * rsqrt(pq); }
                        rys_2d<LA + 2, LB + 2, LC + 2, LD>(w0, PA - rx_p * PQ, QC + rx_q * PQ, b10, b01, b00, ab, cd, g + (size_t)tk * GSZ);
                    }
                    __syncthreads();
                    const int nrb = min(NRB, NR - r0);
                    for (int n = lane; n < NOUT; n += G) {
                        int ea[3], eb[3], ec[3], ed[3];
                        cart_comp(LA, n / (ND * NC * NB), ea[0], ea[1], ea[2]);
                        cart_comp(LB, (n / (ND * NC)) % NB, eb[0], eb[1], eb[2]);
                        cart_comp(LC, (n / ND) % NC, ec[0], ec[1], ec[2]);
                        cart_comp(LD, n % ND, ed[0], ed[1], ed[2]);
                        int id3[3];
#pragma unroll
                        for (int d = 0; d < 3; d++) id3[d] = ea[d] * SA + eb[d] * SB + ec[d] * SC + ed[d];
                        const double w = gam[n];
                        for (int r = 0; r < nrb; r++) {
                            const double* gr = g + (size_t)(3 * r) * GSZ;
                            double f[3], d1[3][3], d2[6][3];     // d1[centre][dim]; d2: AA, BB, CC, AB, AC, BC
#pragma unroll
                            for (int d = 0; d < 3; d++) {
                                const double* gd = gr + d * GSZ + id3[d];
                                const int i = ea[d], j = eb[d], k = ec[d];
                                const double f0 = gd[0];
                                const double ap = gd[SA], am = i ? gd[-SA] : 0.0;
                                const double bp = gd[SB], bm = j ? gd[-SB] : 0.0;
                                const double cp = gd[SC], cm = k ? gd[-SC] : 0.0;
                                f[d] = f0;
                                d1[0][d] = ta * ap - i * am;
                                d1[1][d] = tb * bp - j * bm;
                                d1[2][d] = tc * cp - k * cm;
                                d2[0][d] = taa * gd[2 * SA] - ta * (2 * i + 1) * f0 + (i > 1 ? i * (i - 1) * gd[-2 * SA] : 0.0);
                                d2[1][d] = tbb * gd[2 * SB] - tb * (2 * j + 1) * f0 + (j > 1 ? j * (j - 1) * gd[-2 * SB] : 0.0);
                                d2[2][d] = tcc * gd[2 * SC] - tc * (2 * k + 1) * f0 + (k > 1 ? k * (k - 1) * gd[-2 * SC] : 0.0);
                                d2[3][d] = tab * gd[SA + SB] - (j ? ta * j * gd[SA - SB] : 0.0) - (i ? i * tb * gd[SB - SA] : 0.0) +
                                           ((i && j) ? i * j * gd[-SA - SB] : 0.0);
                                d2[4][d] = tac * gd[SA + SC] - (k ? ta * k * gd[SA - SC] : 0.0) - (i ? i * tc * gd[SC - SA] : 0.0) +
                                           ((i && k) ? i * k * gd[-SA - SC] : 0.0);
                                d2[5][d] = tbc * gd[SB + SC] - (k ? tb * k * gd[SB - SC] : 0.0) - (j ? j * tc * gd[SC - SB] : 0.0) +
                                           ((j && k) ? j * k * gd[-SB - SC] : 0.0);
                            }
                            const double fw[3] = {f[0] * w, f[1] * w, f[2] * w};
                            const double pw[3] = {f[1] * f[2] * w, f[0] * f[2] * w, f[0] * f[1] * w};   // product of the other two directions
#pragma unroll
                            for (int u = 0; u < 9; u++) {
#pragma unroll
                                for (int v = u; v < 9; v++) {
                                    const int P = u / 3, tu = u % 3, Q = v / 3, tv = v % 3;
                                    double val;
                                    if (tu == tv) {
                                        const int pi = (P == Q) ? P : (P == 0 ? (Q == 1 ? 3 : 4) : 5);
                                        val = d2[pi][tu] * pw[tu];
                                    } else {
                                        val = d1[P][tu] * d1[Q][tv] * fw[3 - tu - tv];
                                    }
                                    acc[hess_tri(u, v)] += val;
                                }
                            }
                        }
                    }
                    }   // root batches
                }
            }

            // ---- CTA-wide sums of the 45 scalars (fixed order), the 12 x 12 block, then atomics into the Hessian -----------
#pragma unroll
            for (int k = 0; k < 45; k++) {
                double v = acc[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[k] = v;
            }
            __syncthreads();       // tables of the last primitive quartet consumed
            if ((lane & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 45; k++) red[(lane >> 5) * 45 + k] = acc[k];
            }
            __syncthreads();
            for (int e = lane; e < 81; e += G) {       // 9 x 9 part: both triangles
                const int u = e / 9, v = e % 9;
                const int k = u <= v ? hess_tri(u, v) : hess_tri(v, u);
                double s = 0.0;
                for (int wp = 0; wp < G / 32; wp++) s += red[wp * 45 + k];
                blk[u * 12 + v] = s;
            }
            __syncthreads();
            for (int e = lane; e < 27; e += G) {       // rows / columns of centre D from translational invariance: M[D_s][j] = - sum_P M[P_s][j], j < 9
                const int s_ = e / 9, j = e % 9;
                const double v = -(blk[s_ * 12 + j] + blk[(3 + s_) * 12 + j] + blk[(6 + s_) * 12 + j]);
                blk[(9 + s_) * 12 + j] = v;
                blk[j * 12 + 9 + s_] = v;
            }
            __syncthreads();
            if (lane < 9) {        // M[D_s][D_s'] = - sum_P M[D_s][P_s']
                const int s_ = lane / 3, s2 = lane % 3;
                blk[(9 + s_) * 12 + 9 + s2] = -(blk[(9 + s_) * 12 + s2] + blk[(9 + s_) * 12 + 3 + s2] + blk[(9 + s_) * 12 + 6 + s2]);
            }
            __syncthreads();
            {
                const int atA = t.shell2atom[sa], atB = t.shell2atom[sb], atC = t.shell2atom[sc], atD = t.shell2atom[sd];
                for (int e = lane; e < 144; e += G) {
                    const int u = e / 12, v = e % 12;
                    const int au = u < 3 ? atA : u < 6 ? atB : u < 9 ? atC : atD;
                    const int av = v < 3 ? atA : v < 6 ? atB : v < 9 ? atC : atD;
                    atomicAdd(t.gmat + (size_t)(3 * av + v % 3) * t.ngrad + 3 * au + u % 3, blk[e]);
                }
            }
        }
    }
}
