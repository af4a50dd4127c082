// eri_grad.cuh -- first-derivative ERIs contracted on the fly into the nuclear gradient.
//
// Replaces getRepulsion1 + Int4C2E::ContractGrads(D1, D2) (src/Integral/Int4C2E.cpp:312-408, :747-763):
//     grad[A,x] = sum_ij D1_ij * d/dA_x ( J[2 D2] - exx K[D2] )_ij          (densities fixed)
//               = sum over canonical shell quartets  sum_abcd  d(ab|cd)/dA_x * Gamma_abcd ,
//     Gamma_abcd = wgt * [ D1_ab D2_cd + D1_cd D2_ab - exx/4 (D1_ac D2_bd + D1_bc D2_ad + D1_ad D2_bc + D1_bd D2_ac) ]
// (the reference's 8-fold degeneracy weights, :377-383, folded into `wgt`; its 3*natom nbf x nbf matrices are never
// formed -- the 12 derivative buffers libint2 hands it are contracted with Gamma while they are still in registers).
// Rys quadrature: d/dA_x acts on the x-direction 2-D integral only,
//     d/dA_x I_x(i,j,k,l) = 2 alpha_a I_x(i+1,j,k,l) - i I_x(i-1,j,k,l)
// so one 2-D table with every index raised by one serves all centres; centre D follows from translational
// invariance.  One CTA of G threads per shell quartet, same phases as eri_generic.cuh; each lane owns Cartesian
// components, keeps Gamma for them in registers and accumulates 9 scalars.  Per-CTA partial gradients are written
// without atomics to the CTA's own row and summed in a fixed order afterwards (deterministic for a fixed launch).
#pragma once
#include "cf_common.cuh"
#include "eri_generic.cuh"
#include "eri_tpq.cuh"

struct GradTask {
    PairClassDev bra, ket;
    const double* bra_aexp;      // exponent of shell a of every bra primitive pair (slot layout of PairClassDev)
    const double* ket_aexp;      // ... of shell c
    const long long* qoff;       // [bra.npair+1] first quartet of each bra pair
    long long nquartet;
    int chunk, rank, world, same_class, ncart;
    const double* D1;            // Cartesian working basis, symmetric
    const double* D2;
    double exx;                  // <= 0: Coulomb part only
    const int* shell2atom;
    double* gpart;               // [gridDim.x][ngrad] per-CTA partial gradients
    int ngrad;                   // 3 * natom
    double* gmat;                // matrix form (eri_gradmat_generic): [3 * natom][ncart * ncart] raw Cartesian accumulators
    RysTablesDev rys;
    double prim_cut, thr;
};

template <int LA, int LB, int LC, int LD>
constexpr size_t eri_grad_smem(int G) {
    constexpr int NR = (LA + LB + LC + LD + 1) / 2 + 1;
    constexpr int GSZ2 = (LA + 2) * (LB + 2) * (LC + 2) * (LD + 2);
    constexpr int NOUT = cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD);
    return sizeof(double) * (size_t)(2 * NR + NR * 3 * GSZ2 + NOUT + (G / 32) * 9 + 16);
}

template <int LA, int LB, int LC, int LD, int G>
__global__ void __launch_bounds__(G) eri_grad_generic(const GradTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NOUT = NA * NB * NC * ND;
    constexpr int NR = (LA + LB + LC + LD + 1) / 2 + 1;
    constexpr int GSZ2 = (LA + 2) * (LB + 2) * (LC + 2) * (LD + 2);
    constexpr int SA = (LB + 2) * (LC + 2) * (LD + 2), SB = (LC + 2) * (LD + 2), SC = (LD + 2);
    extern __shared__ double smem[];
    double* rw = smem;
    double* g = rw + 2 * NR;
    double* gam = g + NR * 3 * GSZ2;        // [NOUT] effective two-particle density of the quartet
    double* red = gam + NOUT;               // [G/32][9]
    const int lane = threadIdx.x;
    const size_t ld = (size_t)t.ncart;
    double* myrow = t.gpart + (size_t)blockIdx.x * t.ngrad;

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

            // effective two-particle density of the quartet (component n is read back by the lane that wrote it)
            for (int n = lane; n < NOUT; n += G) {
                const int a = ca + n / (ND * NC * NB), b = cb + (n / (ND * NC)) % NB, c = cc + (n / ND) % NC, d = cdd + n % ND;
                double v = t.D1[b * ld + a] * t.D2[d * ld + c] + t.D1[d * ld + c] * t.D2[b * ld + a];
                if (t.exx > 0.0)
                    v -= 0.25 * t.exx * (t.D1[c * ld + a] * t.D2[d * ld + b] + t.D1[c * ld + b] * t.D2[d * ld + a] +
                                         t.D1[d * ld + a] * t.D2[c * ld + b] + t.D1[d * ld + b] * t.D2[c * ld + a]);
                gam[n] = v * wgt;
            }
            double acc[9];
#pragma unroll
            for (int e = 0; e < 9; e++) acc[e] = 0.0;

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
                    __syncthreads();                               // previous tables consumed
                    if (lane < 2 * NR) rw[lane] = rys_value<NR>(t.rys, T, lane);
                    __syncthreads();
                    for (int tk = lane; tk < 3 * NR; tk += G) {
                        const int r = tk / 3, dim = tk - 3 * r;
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
                        rys_2d<LA + 1, LB + 1, LC + 1, LD + 1>(w0, PA - rx_p * PQ, QC + rx_q * PQ, b10, b01, b00, ab, cd, g + (size_t)tk * GSZ2);
                    }
                    __syncthreads();
                    for (int n = lane; n < NOUT; n += G) {
                        int ea[3], eb[3], ec[3], ed[3];
                        cart_comp(LA, n / (ND * NC * NB), ea[0], ea[1], ea[2]);
                        cart_comp(LB, (n / (ND * NC)) % NB, eb[0], eb[1], eb[2]);
                        cart_comp(LC, (n / ND) % NC, ec[0], ec[1], ec[2]);
                        cart_comp(LD, n % ND, ed[0], ed[1], ed[2]);
                        int id3[3];
#pragma unroll
                        for (int d = 0; d < 3; d++) id3[d] = ea[d] * SA + eb[d] * SB + ec[d] * SC + ed[d];
                        double s[9];
#pragma unroll
                        for (int k = 0; k < 9; k++) s[k] = 0.0;
                        for (int r = 0; r < NR; r++) {
                            const double* gr = g + (size_t)(3 * r) * GSZ2;
                            double f[3], dA[3], dB[3], dC[3];
#pragma unroll
                            for (int d = 0; d < 3; d++) {
                                const double* gd = gr + d * GSZ2 + id3[d];
                                f[d] = gd[0];
                                dA[d] = ta * gd[SA] - (ea[d] ? ea[d] * gd[-SA] : 0.0);
                                dB[d] = tb * gd[SB] - (eb[d] ? eb[d] * gd[-SB] : 0.0);
                                dC[d] = tc * gd[SC] - (ec[d] ? ec[d] * gd[-SC] : 0.0);
                            }
                            const double fyz = f[1] * f[2], fxz = f[0] * f[2], fxy = f[0] * f[1];
                            s[0] = fma(dA[0], fyz, s[0]); s[1] = fma(dA[1], fxz, s[1]); s[2] = fma(dA[2], fxy, s[2]);
                            s[3] = fma(dB[0], fyz, s[3]); s[4] = fma(dB[1], fxz, s[4]); s[5] = fma(dB[2], fxy, s[5]);
                            s[6] = fma(dC[0], fyz, s[6]); s[7] = fma(dC[1], fxz, s[7]); s[8] = fma(dC[2], fxy, s[8]);
                        }
#pragma unroll
                        for (int k = 0; k < 9; k++) acc[k] = fma(gam[n], s[k], acc[k]);
                    }
                }
            }

            // ---- CTA-wide sums of the 9 scalars (fixed order), then this CTA's private row --------------------------
#pragma unroll
            for (int k = 0; k < 9; k++) {
                double v = acc[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[k] = v;
            }
            __syncthreads();       // tables of the last primitive quartet consumed; `red` free
            if ((lane & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 9; k++) red[(lane >> 5) * 9 + k] = acc[k];
            }
            __syncthreads();
            if (lane == 0) {
                double tot[9];
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    double v = 0.0;
                    for (int w = 0; w < G / 32; w++) v += red[w * 9 + k];
                    tot[k] = v;
                }
                const int atA = t.shell2atom[sa], atB = t.shell2atom[sb], atC = t.shell2atom[sc], atD = t.shell2atom[sd];
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    myrow[3 * atA + x] += tot[x];
                    myrow[3 * atB + x] += tot[3 + x];
                    myrow[3 * atC + x] += tot[6 + x];
                    myrow[3 * atD + x] -= tot[x] + tot[3 + x] + tot[6 + x];     // translational invariance
                }
            }
        }
    }
}

// ================================================================================================
// Thread-per-quartet gradient kernel for the small and medium classes (<= 108 Cartesian components, raised 2-D tables of
// <= 100 entries, <= 4 roots): the CTA-per-quartet kernel above spends three barriers per primitive quartet with most
// lanes idle there (bo3h3: ps|ss and ss|ss were 50 % of the gradient).  The larger of these classes spill their tables to
// local memory and still win when the task has enough quartets to fill the machine (c18 3.97 -> 1.21 s); tasks with few
// quartets of a larger class stay with the CTA-per-quartet kernel (launch_grad_pair in eri_inst.cu).  Here every thread owns one shell quartet of the same (ib, ik)
// enumeration, keeps the raised tables of one root in registers and accumulates its 12 scalars; the lanes of a warp then
// add them one after the other (fixed order) into the warp's shared-memory row, and the rows go to the CTA's private
// global row at the end -- still no floating-point atomics, still bit-repeatable for a given launch geometry.
// ================================================================================================
#define GRAD_TPQ_THREADS 128
__host__ __device__ constexpr bool grad_tpq_ok(int la, int lb, int lc, int ld) {
#ifndef GRAD_TPQ_MAXOUT
#define GRAD_TPQ_MAXOUT 108
#define GRAD_TPQ_MAXTAB 100
#endif
#ifndef GRAD_TPQ_MAXNR
#define GRAD_TPQ_MAXNR 4
#endif
    return cf_ncart(la) * cf_ncart(lb) * cf_ncart(lc) * cf_ncart(ld) <= GRAD_TPQ_MAXOUT &&
           (la + 2) * (lb + 2) * (lc + 2) * (ld + 2) <= GRAD_TPQ_MAXTAB && (la + lb + lc + ld + 1) / 2 + 1 <= GRAD_TPQ_MAXNR;
}

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(GRAD_TPQ_THREADS) eri_grad_tpq(const GradTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NOUT = NA * NB * NC * ND;
    constexpr int NR = (LA + LB + LC + LD + 1) / 2 + 1;
    constexpr int GSZ2 = (LA + 2) * (LB + 2) * (LC + 2) * (LD + 2);
    constexpr int SA = (LB + 2) * (LC + 2) * (LD + 2), SB = (LC + 2) * (LD + 2), SC = (LD + 2);
    constexpr int TABLEN = tpq_table_len(NR);
    extern __shared__ double smem[];
    double* tab = smem;                                       // Boys rows (1-2 roots) or Chebyshev tables, as in eri_tpq.cuh
    double* rows = smem + TABLEN;                             // [warps][ngrad]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* myrow = rows + (size_t)warp * t.ngrad;
    tpq_stage_tables<NR>(tab, t.rys, threadIdx.x, GRAD_TPQ_THREADS);
    for (int e = threadIdx.x; e < (GRAD_TPQ_THREADS / 32) * t.ngrad; e += GRAD_TPQ_THREADS) rows[e] = 0.0;
    __syncthreads();
    const size_t ld = (size_t)t.ncart;
    const long long nblk_total = (t.nquartet + GRAD_TPQ_THREADS - 1) / GRAD_TPQ_THREADS;
    const long long nblk_local = (nblk_total - t.rank + t.world - 1) / t.world;
    for (long long lb_ = blockIdx.x; lb_ < nblk_local; lb_ += gridDim.x) {
        const long long q = (lb_ * t.world + t.rank) * GRAD_TPQ_THREADS + threadIdx.x;
        bool act = q < t.nquartet;
        int ib = 0, ik = 0;
        if (act) {
            int lo = 0, hi = t.bra.npair;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (t.qoff[mid] <= q) lo = mid; else hi = mid;
            }
            ib = lo; ik = (int)(q - t.qoff[ib]);
            if (t.thr > 0.0 && !(t.bra.Q[ib] * t.ket.Q[ik] > t.thr)) act = false;
        }
        double acc[9];
#pragma unroll
        for (int e = 0; e < 9; e++) acc[e] = 0.0;
        int atA = 0, atB = 0, atC = 0, atD = 0;
        if (act) {
            const int sa = t.bra.sa[ib], sb = t.bra.sb[ib], sc = t.ket.sa[ik], sd = t.ket.sb[ik];
            atA = t.shell2atom[sa]; atB = t.shell2atom[sb]; atC = t.shell2atom[sc]; atD = t.shell2atom[sd];
            const double Ax = t.bra.A[3 * ib], Ay = t.bra.A[3 * ib + 1], Az = t.bra.A[3 * ib + 2];
            const double ABx = t.bra.AB[3 * ib], ABy = t.bra.AB[3 * ib + 1], ABz = t.bra.AB[3 * ib + 2];
            const double Cx = t.ket.A[3 * ik], Cy = t.ket.A[3 * ik + 1], Cz = t.ket.A[3 * ik + 2];
            const double CDx = t.ket.AB[3 * ik], CDy = t.ket.AB[3 * ik + 1], CDz = t.ket.AB[3 * ik + 2];
            const int pab0 = t.bra.pbase[ib], npab = t.bra.nprim[ib];
            const int pcd0 = t.ket.pbase[ik], npcd = t.ket.nprim[ik];
            double wgt = (sa == sb ? 1.0 : 2.0) * (sc == sd ? 1.0 : 2.0);
            wgt *= (t.same_class && ib == ik) ? 1.0 : 2.0;
            const int ca = t.bra.cao_a[ib], cb = t.bra.cao_b[ib], cc = t.ket.cao_a[ik], cdd = t.ket.cao_b[ik];
            double gam[NOUT];
#pragma unroll
            for (int n = 0; n < NOUT; n++) {
                const int a = ca + n / (ND * NC * NB), b = cb + (n / (ND * NC)) % NB, c = cc + (n / ND) % NC, d = cdd + n % ND;
                double v = t.D1[b * ld + a] * t.D2[d * ld + c] + t.D1[d * ld + c] * t.D2[b * ld + a];
                if (t.exx > 0.0)
                    v -= 0.25 * t.exx * (t.D1[c * ld + a] * t.D2[d * ld + b] + t.D1[c * ld + b] * t.D2[d * ld + a] +
                                         t.D1[d * ld + a] * t.D2[c * ld + b] + t.D1[d * ld + b] * t.D2[c * ld + a]);
                gam[n] = v * wgt;
            }
            for (int iab = 0; iab < npab; iab++) {
                const int sab = pab0 + iab * CF_PSTRIDE;
                const double p = t.bra.p[sab], cab = t.bra.c[sab], hp = t.bra.hp[sab];
                const double Px = t.bra.Px[sab], Py = t.bra.Py[sab], Pz = t.bra.Pz[sab];
                const double ta = 2.0 * t.bra_aexp[sab], tb = 2.0 * p - ta;
                for (int icd = 0; icd < npcd; icd++) {
                    const int scd = pcd0 + icd * CF_PSTRIDE;
                    const double cc_ = cab * t.ket.c[scd];
                    if (fabs(cc_) < t.prim_cut) continue;
                    const double qe = t.ket.p[scd], hq = t.ket.hp[scd];
                    const double Qx = t.ket.Px[scd], Qy = t.ket.Py[scd], Qz = t.ket.Pz[scd];
                    const double tc = 2.0 * t.ket_aexp[scd];
                    const double pq = p + qe;
                    const double rs = rsqrt(pq), ipq = rs * rs;
                    const double PQx = Px - Qx, PQy = Py - Qy, PQz = Pz - Qz;
                    const double T = (p * qe * ipq) * fma(PQx, PQx, fma(PQy, PQy, PQz * PQz));
                    double rx[NR], rw[NR];
                    tpq_roots<NR>(tab, T, rx, rw);
                    const double qi = qe * ipq, pi_ = p * ipq, hi = 0.5 * ipq;
#pragma unroll
                    for (int r = 0; r < NR; r++) {
                        const double xr = rx[r];
                        const double rxp = xr * qi, rxq = xr * pi_, b00 = xr * hi;
                        const double b10 = fma(-rxp, hp, hp), b01 = fma(-rxq, hq, hq);
                        double g3[3][GSZ2];
                        rys_2d<LA + 1, LB + 1, LC + 1, LD + 1>(1.0, fma(-rxp, PQx, Px - Ax), fma(rxq, PQx, Qx - Cx), b10, b01, b00, ABx, CDx, g3[0]);
                        rys_2d<LA + 1, LB + 1, LC + 1, LD + 1>(1.0, fma(-rxp, PQy, Py - Ay), fma(rxq, PQy, Qy - Cy), b10, b01, b00, ABy, CDy, g3[1]);
                        rys_2d<LA + 1, LB + 1, LC + 1, LD + 1>(rw[r] * cc_ * rs, fma(-rxp, PQz, Pz - Az), fma(rxq, PQz, Qz - Cz), b10, b01, b00, ABz, CDz, g3[2]);
#pragma unroll
                        for (int n = 0; n < NOUT; n++) {
                            const int id = n % ND, ic = (n / ND) % NC, jb = (n / (ND * NC)) % NB, ia = n / (ND * NC * NB);
                            const int ea[3] = {cart_lx(LA, ia), cart_ly(LA, ia), cart_lz(LA, ia)};
                            const int eb[3] = {cart_lx(LB, jb), cart_ly(LB, jb), cart_lz(LB, jb)};
                            const int ec[3] = {cart_lx(LC, ic), cart_ly(LC, ic), cart_lz(LC, ic)};
                            const int ed[3] = {cart_lx(LD, id), cart_ly(LD, id), cart_lz(LD, id)};
                            double f[3], dA[3], dB[3], dC[3];
#pragma unroll
                            for (int d = 0; d < 3; d++) {
                                const int i0 = ea[d] * SA + eb[d] * SB + ec[d] * SC + ed[d];
                                f[d] = g3[d][i0];
                                dA[d] = ta * g3[d][i0 + SA]; if (ea[d] > 0) dA[d] -= ea[d] * g3[d][i0 - SA];
                                dB[d] = tb * g3[d][i0 + SB]; if (eb[d] > 0) dB[d] -= eb[d] * g3[d][i0 - SB];
                                dC[d] = tc * g3[d][i0 + SC]; if (ec[d] > 0) dC[d] -= ec[d] * g3[d][i0 - SC];
                            }
                            const double fyz = f[1] * f[2] * gam[n], fxz = f[0] * f[2] * gam[n], fxy = f[0] * f[1] * gam[n];
                            acc[0] = fma(dA[0], fyz, acc[0]); acc[1] = fma(dA[1], fxz, acc[1]); acc[2] = fma(dA[2], fxy, acc[2]);
                            acc[3] = fma(dB[0], fyz, acc[3]); acc[4] = fma(dB[1], fxz, acc[4]); acc[5] = fma(dB[2], fxy, acc[5]);
                            acc[6] = fma(dC[0], fyz, acc[6]); acc[7] = fma(dC[1], fxz, acc[7]); acc[8] = fma(dC[2], fxy, acc[8]);
                        }
                    }
                }
            }
        }
        // lanes add one after the other (fixed order) into the warp's row
        const unsigned amask = __ballot_sync(0xffffffffu, act);
        for (unsigned rest = amask; rest; rest &= rest - 1) {
            const int l = __ffs(rest) - 1;
            if (lane == l) {
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    myrow[3 * atA + x] += acc[x];
                    myrow[3 * atB + x] += acc[3 + x];
                    myrow[3 * atC + x] += acc[6 + x];
                    myrow[3 * atD + x] -= acc[x] + acc[3 + x] + acc[6 + x];
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    double* grow = t.gpart + (size_t)blockIdx.x * t.ngrad;
    for (int j = threadIdx.x; j < t.ngrad; j += GRAD_TPQ_THREADS) {
        double s = 0.0;
        for (int w = 0; w < GRAD_TPQ_THREADS / 32; w++) s += rows[(size_t)w * t.ngrad + j];
        grow[j] += s;
    }
}

// ================================================================================================
// Matrix form: the 3*natom matrices G^(A,x)[D] = d/dA_x ( J[2 D] - exx K[D] ) that Int4C2E::ContractGrads(D) returns
// (src/Integral/Int4C2E.cpp:766-790 over getRepulsion1 :312-408; consumer Restricted/Hess.cpp:72).  Same derivative
// integrals as above, but instead of the scalar contraction with Gamma every derivative block is digested like a J/K
// build (the reference's six scatter updates per unique integral, :377-389): in the Cartesian working basis
//     raw^(X)(a,b) += 1/2 w dV(abcd) D(c,d)      raw^(X)(c,d) += 1/2 w dV D(a,b)
//     raw^(X)(a,c) -= exx/8 w dV D(b,d)   (a,d), (b,c), (b,d) alike,          G^(X) = raw + raw^T  (then Cartesian -> pure),
// X = (atom of the differentiated centre, direction); centre D from translational invariance.  One CTA per quartet; the
// nine contracted derivative blocks live in registers / local memory per lane, and each goes through shared memory for
// an owner-computes digestion with FP64 atomics (this path feeds the Hessian driver, not the SCF loop).
// ================================================================================================
template <int LA, int LB, int LC, int LD>
constexpr size_t eri_gradmat_smem(int G) {
    constexpr int NR = (LA + LB + LC + LD + 1) / 2 + 1;
    constexpr int GSZ2 = (LA + 2) * (LB + 2) * (LC + 2) * (LD + 2);
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    return sizeof(double) * (size_t)(2 * NR + NR * 3 * GSZ2 + NA * NB * NC * ND + NA * NB + NC * ND + NA * NC + NA * ND + NB * NC + NB * ND + 16);
}

template <int LA, int LB, int LC, int LD, int G>
__global__ void __launch_bounds__(G) eri_gradmat_generic(const GradTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND, NOUT = NAB * NCD;
    constexpr int NR = (LA + LB + LC + LD + 1) / 2 + 1;
    constexpr int GSZ2 = (LA + 2) * (LB + 2) * (LC + 2) * (LD + 2);
    constexpr int SA = (LB + 2) * (LC + 2) * (LD + 2), SB = (LC + 2) * (LD + 2), SC = (LD + 2);
    constexpr int NPL = (NOUT + G - 1) / G;
    extern __shared__ double smem[];
    double* rw = smem;
    double* g = rw + 2 * NR;
    double* V = g + NR * 3 * GSZ2;          // [NOUT] one derivative block at a time
    double* Dab = V + NOUT;                 // density blocks of the quartet
    double* Dcd = Dab + NAB;
    double* Dac = Dcd + NCD;
    double* Dad = Dac + NA * NC;
    double* Dbc = Dad + NA * ND;
    double* Dbd = Dbc + NB * NC;
    const int lane = threadIdx.x;
    const size_t ld = (size_t)t.ncart, n2c = ld * ld;
    const double* D = t.D2;

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
            if (t.thr > 0.0 && !(t.bra.Q[ib] * t.ket.Q[ik] > t.thr)) continue;
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

            // ---- density blocks of the quartet (bra shell as the column, like the J/K kernels) ----------------------
            __syncthreads();
            for (int e = lane; e < NAB; e += G) Dab[e] = D[(size_t)(cb + e % NB) * ld + ca + e / NB];
            for (int e = lane; e < NCD; e += G) Dcd[e] = D[(size_t)(cdd + e % ND) * ld + cc + e / ND];
            for (int e = lane; e < NA * NC; e += G) Dac[e] = D[(size_t)(ca + e / NC) * ld + cc + e % NC];
            for (int e = lane; e < NA * ND; e += G) Dad[e] = D[(size_t)(ca + e / ND) * ld + cdd + e % ND];
            for (int e = lane; e < NB * NC; e += G) Dbc[e] = D[(size_t)(cb + e / NC) * ld + cc + e % NC];
            for (int e = lane; e < NB * ND; e += G) Dbd[e] = D[(size_t)(cb + e / ND) * ld + cdd + e % ND];
            const int atA = t.shell2atom[sa], atB = t.shell2atom[sb], atC = t.shell2atom[sc], atD = t.shell2atom[sd];
            const double fj = 0.5, fk = t.exx > 0.0 ? -0.125 * t.exx : 0.0;

            // one Cartesian direction at a time (the integrals are recomputed per direction: three accumulators per
            // component instead of nine keep the largest classes out of local memory; this path is not the SCF loop)
            for (int dir = 0; dir < 3; dir++) {
            double dvA[NPL], dvB[NPL], dvC[NPL];       // contracted d/dA_dir, d/dB_dir, d/dC_dir of this lane's components
#pragma unroll
            for (int m = 0; m < NPL; m++) { dvA[m] = 0.0; dvB[m] = 0.0; dvC[m] = 0.0; }
            const int d1 = dir == 0 ? 1 : 0, d2 = dir == 2 ? 1 : 2;     // the two other directions

            for (int iab = 0; iab < npab; iab++) {
                const int sab = pab0 + iab * CF_PSTRIDE;
                const double p = t.bra.p[sab], cab = t.bra.c[sab];
                const double Px = t.bra.Px[sab], Py = t.bra.Py[sab], Pz = t.bra.Pz[sab];
                const double ta = 2.0 * t.bra_aexp[sab], tb = 2.0 * p - ta;
                for (int icd = 0; icd < npcd; icd++) {
                    const int scd = pcd0 + icd * CF_PSTRIDE;
                    const double ccd = t.ket.c[scd];
                    if (fabs(cab * ccd) < t.prim_cut) continue;
                    const double qe = t.ket.p[scd];
                    const double Qx = t.ket.Px[scd], Qy = t.ket.Py[scd], Qz = t.ket.Pz[scd];
                    const double tc = 2.0 * t.ket_aexp[scd];
                    const double pq = p + qe;
                    const double rho = p * qe / pq;
                    const double PQx = Px - Qx, PQy = Py - Qy, PQz = Pz - Qz;
                    const double T = rho * (PQx * PQx + PQy * PQy + PQz * PQz);
                    __syncthreads();
                    if (lane < 2 * NR) rw[lane] = rys_value<NR>(t.rys, T, lane);
                    __syncthreads();
                    for (int tk = lane; tk < 3 * NR; tk += G) {
                        const int r = tk / 3, dim = tk - 3 * r;
                        const double x = rw[r];
                        const double rx_p = rho * x / p, rx_q = rho * x / qe;
                        const double b00 = 0.5 * x / pq, b10 = (1.0 - rx_p) * (0.5 / p), b01 = (1.0 - rx_q) * (0.5 / qe);
                        double PA, PQ, QC, ab, cd, w0;
                        if (dim == 0) { PA = Px - Ax; PQ = PQx; QC = Qx - Cx; ab = ABx; cd = CDx; w0 = 1.0; }
                        else if (dim == 1) { PA = Py - Ay; PQ = PQy; QC = Qy - Cy; ab = ABy; cd = CDy; w0 = 1.0; }
                        else { PA = Pz - Az; PQ = PQz; QC = Qz - Cz; ab = ABz; cd = CDz; w0 = rw[NR + r] * cab * ccd * rsqrt(pq) * wgt; }
                        rys_2d<LA + 1, LB + 1, LC + 1, LD + 1>(w0, PA - rx_p * PQ, QC + rx_q * PQ, b10, b01, b00, ab, cd, g + (size_t)tk * GSZ2);
                    }
                    __syncthreads();
#pragma unroll
                    for (int m = 0; m < NPL; m++) {
                        const int n = lane + m * G;
                        if (n < NOUT) {
                            int ea[3], eb[3], ec[3], ed[3];
                            cart_comp(LA, n / (ND * NC * NB), ea[0], ea[1], ea[2]);
                            cart_comp(LB, (n / (ND * NC)) % NB, eb[0], eb[1], eb[2]);
                            cart_comp(LC, (n / ND) % NC, ec[0], ec[1], ec[2]);
                            cart_comp(LD, n % ND, ed[0], ed[1], ed[2]);
                            const int i0 = ea[dir] * SA + eb[dir] * SB + ec[dir] * SC + ed[dir];      // differentiated direction
                            const int i1 = ea[d1] * SA + eb[d1] * SB + ec[d1] * SC + ed[d1];
                            const int i2 = ea[d2] * SA + eb[d2] * SB + ec[d2] * SC + ed[d2];
                            const int na_ = ea[dir], nb_ = eb[dir], nc_ = ec[dir];
                            double sA = 0.0, sB = 0.0, sC = 0.0;
                            for (int r = 0; r < NR; r++) {
                                const double* gr = g + (size_t)(3 * r) * GSZ2;
                                const double* gd = gr + dir * GSZ2 + i0;
                                const double ff = gr[d1 * GSZ2 + i1] * gr[d2 * GSZ2 + i2];
                                sA = fma(ta * gd[SA] - (na_ ? na_ * gd[-SA] : 0.0), ff, sA);
                                sB = fma(tb * gd[SB] - (nb_ ? nb_ * gd[-SB] : 0.0), ff, sB);
                                sC = fma(tc * gd[SC] - (nc_ ? nc_ * gd[-SC] : 0.0), ff, sC);
                            }
                            dvA[m] += sA; dvB[m] += sB; dvC[m] += sC;
                        }
                    }
                }
            }

            // four derivative blocks of this direction: centres A, B, C, and D = -(A + B + C)
            for (int cen = 0; cen < 4; cen++) {
                __syncthreads();
#pragma unroll
                for (int m = 0; m < NPL; m++) {
                    const int n = lane + m * G;
                    if (n < NOUT) V[n] = cen == 0 ? dvA[m] : cen == 1 ? dvB[m] : cen == 2 ? dvC[m] : -(dvA[m] + dvB[m] + dvC[m]);
                }
                __syncthreads();
                const int atom = cen == 0 ? atA : cen == 1 ? atB : cen == 2 ? atC : atD;
                double* Gm = t.gmat + (size_t)(3 * atom + dir) * n2c;
                for (int e = lane; e < NAB + NCD; e += G) {
                    double s = 0.0;
                    if (e < NAB) {
                        for (int kl = 0; kl < NCD; kl++) s = fma(V[e * NCD + kl], Dcd[kl], s);
                        atomicAdd(Gm + (size_t)(cb + e % NB) * ld + ca + e / NB, fj * s);
                    } else {
                        const int kl = e - NAB;
                        for (int ij = 0; ij < NAB; ij++) s = fma(V[ij * NCD + kl], Dab[ij], s);
                        atomicAdd(Gm + (size_t)(cdd + kl % ND) * ld + cc + kl / ND, fj * s);
                    }
                }
                if (fk != 0.0) {
                    constexpr int NDX = NA * NC + NA * ND + NB * NC + NB * ND;
                    for (int e = lane; e < NDX; e += G) {
                        double s = 0.0;
                        if (e < NA * NC) {
                            const int i = e / NC, k = e % NC;
                            for (int j = 0; j < NB; j++)
                                for (int l = 0; l < ND; l++) s = fma(V[((i * NB + j) * NC + k) * ND + l], Dbd[j * ND + l], s);
                            atomicAdd(Gm + (size_t)(ca + i) * ld + cc + k, fk * s);
                        } else if (e < NA * NC + NA * ND) {
                            const int f = e - NA * NC, i = f / ND, l = f % ND;
                            for (int j = 0; j < NB; j++)
                                for (int k = 0; k < NC; k++) s = fma(V[((i * NB + j) * NC + k) * ND + l], Dbc[j * NC + k], s);
                            atomicAdd(Gm + (size_t)(ca + i) * ld + cdd + l, fk * s);
                        } else if (e < NA * NC + NA * ND + NB * NC) {
                            const int f = e - NA * NC - NA * ND, j = f / NC, k = f % NC;
                            for (int i = 0; i < NA; i++)
                                for (int l = 0; l < ND; l++) s = fma(V[((i * NB + j) * NC + k) * ND + l], Dad[i * ND + l], s);
                            atomicAdd(Gm + (size_t)(cb + j) * ld + cc + k, fk * s);
                        } else {
                            const int f = e - NA * NC - NA * ND - NB * NC, j = f / ND, l = f % ND;
                            for (int i = 0; i < NA; i++)
                                for (int k = 0; k < NC; k++) s = fma(V[((i * NB + j) * NC + k) * ND + l], Dac[i * NC + k], s);
                            atomicAdd(Gm + (size_t)(cb + j) * ld + cdd + l, fk * s);
                        }
                    }
                }
            }
            }   // directions
            __syncthreads();
        }
    }
}
