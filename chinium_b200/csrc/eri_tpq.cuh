// eri_tpq.cuh -- thread-per-quartet Rys kernels with fused J/K digestion (classes with <= TPQ_MAX_NOUT
// Cartesian integrals per shell quartet).
//
// Why this shape (B200): the FP64 pipe retires 64 DFMA/clk/SM while shared memory delivers 16 doubles/clk/SM,
// so any scheme that stages the 2-D Rys integrals in shared memory is LSU-bound by ~4x.  Here one thread owns one
// shell quartet and keeps EVERYTHING in registers: roots, the 2-D recurrences of one root at a time, the
// contracted Cartesian integrals and the digestion partial sums.  All index arithmetic is compile-time.
//   work item  = (one bra shell pair, 128 consecutive ket shell pairs); bra primitive data is staged once per item
//                in shared memory and read as warp-uniform broadcasts; ket primitives are interleaved in HBM so
//                a warp reads them coalesced
//   roots      = 1 or 2 roots: from Boys functions (8-term Taylor rows staged in shared memory + downward
//                recursion, then the closed-form 1- and 2-point Gauss rules); >= 3 roots: Chebyshev tables
//                staged in shared memory
//   digestion  = six J/K blocks formed in registers, added to the global accumulators as 64-bit fixed point
//                (integer adds commute -> bit-identical results for any schedule / grid / GPU count)
// Replaces libint2's engine.compute + the reference's stored-integral digestion
// (src/Integral/Int4C2E.cpp:233-302 and :601-671).
#pragma once
#include "cf_common.cuh"
#include "eri_generic.cuh"

#define TPQ_THREADS 128
#define TPQ_MAXBP 64
#define TPQ_MAX_NOUT 64
#define TPQ_NBRA 9       // p, hp, Px, Py, Pz, c, PAx, PAy, PAz

__host__ __device__ constexpr bool tpq_ok(int la, int lb, int lc, int ld) {
    return cf_ncart(la) * cf_ncart(lb) * cf_ncart(lc) * cf_ncart(ld) <= TPQ_MAX_NOUT;
}
__host__ __device__ constexpr int tpq_table_len(int nroots) {
    return nroots <= 2 ? BOYS_NROW * 8 : (rys_tmax(nroots) / 2) * 2 * nroots * RYS_NC + 2 * nroots;
}
__host__ __device__ constexpr size_t tpq_smem(int nroots) {
    return sizeof(double) * (size_t)(tpq_table_len(nroots) + TPQ_NBRA * TPQ_MAXBP);
}

// Cartesian exponents of component n of angular momentum l (order: lx descending, then ly descending);
// branch-free closed forms so that they fold to constants inside fully unrolled loops (l <= 3)
__host__ __device__ constexpr int cart_row(int n) { return n >= 6 ? 3 : n >= 3 ? 2 : n >= 1 ? 1 : 0; }
__host__ __device__ constexpr int cart_lx(int l, int n) { return l - cart_row(n); }
__host__ __device__ constexpr int cart_lz(int l, int n) { return n - cart_row(n) * (cart_row(n) + 1) / 2; }
__host__ __device__ constexpr int cart_ly(int l, int n) { return cart_row(n) - cart_lz(l, n); }

__device__ __forceinline__ double fast_rcp(double x) {   // x > 0, normal range
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// F_M(T) .. F_0(T) for T < BOYS_TMAX from the staged Taylor rows (row i holds F_M..F_{M+7} at T_i = i/8)
template <int M>
__device__ __forceinline__ void boys_small(const double* __restrict__ tab, double T, double* __restrict__ F) {
    const int i = (int)fma(T, 8.0, 0.5);
    const double mh = fma((double)i, 0.125, -T);   // -(T - T_i), |mh| <= 1/16
    const double2* row = reinterpret_cast<const double2*>(tab + i * 8);
    const double2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3];
    double s = r3.y;
    s = fma(s, mh * (1.0 / 7.0), r3.x);
    s = fma(s, mh * (1.0 / 6.0), r2.y);
    s = fma(s, mh * (1.0 / 5.0), r2.x);
    s = fma(s, mh * (1.0 / 4.0), r1.y);
    s = fma(s, mh * (1.0 / 3.0), r1.x);
    s = fma(s, mh * 0.5, r0.y);
    s = fma(s, mh, r0.x);
    F[M] = s;
    const double e = exp(-T), t2 = T + T;
#pragma unroll
    for (int m = M; m > 0; m--) F[m - 1] = fma(t2, F[m], e) * (1.0 / (2 * m - 1));
}
template <int M>
__device__ __forceinline__ void boys_large(double T, double* __restrict__ F) {   // T >= BOYS_TMAX: erfc(sqrt T) < 1e-16
    const double r = rsqrt(T);
    const double e = exp(-T), i2t = 0.5 * r * r;
    F[0] = 0.88622692545275801365 * r;
#pragma unroll
    for (int m = 0; m < M; m++) F[m + 1] = fma((double)(2 * m + 1), F[m], -e) * i2t;
}

// Rys roots x_r (= t^2) and weights (sum_r w_r x_r^k = F_k(T)) of one primitive quartet, all in registers
template <int NROOTS>
__device__ __forceinline__ void tpq_roots(const double* __restrict__ tab, double T, double* __restrict__ x, double* __restrict__ w) {
    if constexpr (NROOTS <= 2) {
        constexpr int M = 2 * NROOTS - 1;
        double F[M + 1];
        if (T < BOYS_TMAX) boys_small<M>(tab, T, F); else boys_large<M>(T, F);
        if constexpr (NROOTS == 1) {
            w[0] = F[0];
            x[0] = F[1] * fast_rcp(F[0]);
        } else {
            // monic orthogonal polynomial x^2 + a x + b of the weight with moments F0..F3
            const double det = fma(F[0], F[2], -F[1] * F[1]);
            const double idet = fast_rcp(det);
            const double b = fma(F[1], F[3], -F[2] * F[2]) * idet;
            const double ma = fma(F[0], F[3], -F[1] * F[2]) * idet;      // -a = x1 + x2
            const double disc = fma(ma, ma, -4.0 * b);
            const double x2 = 0.5 * (ma + sqrt(disc));
            const double x1 = b * fast_rcp(x2);
            const double w2 = fma(-x1, F[0], F[1]) * fast_rcp(x2 - x1);
            x[0] = x1; x[1] = x2;
            w[0] = F[0] - w2; w[1] = w2;
        }
    } else {
        constexpr int NV = 2 * NROOTS;
        if (T >= (double)rys_tmax(NROOTS)) {
            const double* asym = tab + (rys_tmax(NROOTS) / 2) * NV * RYS_NC;
            const double rs = rsqrt(T), it = rs * rs;
#pragma unroll
            for (int v = 0; v < NROOTS; v++) { x[v] = asym[v] * it; w[v] = asym[NROOTS + v] * rs; }
        } else {
            const int it = (int)(T * 0.5);
            const double u = T - (2.0 * it + 1.0), u2 = u + u;
            const double2* cs = reinterpret_cast<const double2*>(tab + (size_t)it * NV * RYS_NC);
#pragma unroll
            for (int v = 0; v < NV; v++) {
                double2 c[RYS_NC / 2];
#pragma unroll
                for (int k = 0; k < RYS_NC / 2; k++) c[k] = cs[v * (RYS_NC / 2) + k];
                double b1 = 0.0, b2 = 0.0;
#pragma unroll
                for (int k = RYS_NC - 1; k >= 1; k--) {
                    const double ck = (k & 1) ? c[k >> 1].y : c[k >> 1].x;
                    const double tt = fma(u2, b1, ck - b2);
                    b2 = b1;
                    b1 = tt;
                }
                const double val = fma(u, b1, c[0].x - b2);
                if (v < NROOTS) x[v] = val; else w[v - NROOTS] = val;
            }
        }
    }
}

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(TPQ_THREADS) eri_jk_tpq(const QuartetTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND, NOUT = NAB * NCD;
    constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    constexpr int GSZ = (LA + 1) * (LB + 1) * (LC + 1) * (LD + 1);
    constexpr int TABLEN = tpq_table_len(NROOTS);
    extern __shared__ double smem[];
    double* tab = smem;
    double* sbra = smem + TABLEN;   // [TPQ_NBRA][TPQ_MAXBP]

    // ---- stage the root tables ------------------------------------------------------------------
    if constexpr (NROOTS <= 2) {
        constexpr int M = 2 * NROOTS - 1;
        for (int e = threadIdx.x; e < BOYS_NROW * 8; e += TPQ_THREADS) tab[e] = t.rys.boys[(e >> 3) * BOYS_NCOL + M + (e & 7)];
    } else {
        constexpr int NT = (rys_tmax(NROOTS) / 2) * 2 * NROOTS * RYS_NC;
        const double* src = t.rys.table + rys_off(NROOTS);
        for (int e = threadIdx.x; e < NT; e += TPQ_THREADS) tab[e] = src[e];
        if (threadIdx.x < 2 * NROOTS) tab[NT + threadIdx.x] = t.rys.asym[rys_asym_off(NROOTS) + threadIdx.x];
    }

    const long long nitem_local = (t.nitem - t.rank + t.world - 1) / t.world;
    for (long long li = blockIdx.x; li < nitem_local; li += gridDim.x) {
        const long long item = li * t.world + t.rank;
        int ib, chunk;
        if (t.same_class) {
            int lo = 0, hi = t.bra.npair;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (t.item_off[mid] <= item) lo = mid; else hi = mid;
            }
            ib = lo; chunk = (int)(item - t.item_off[lo]);
        } else {
            ib = (int)(item / t.nchunk_ket); chunk = (int)(item - (long long)ib * t.nchunk_ket);
        }
        const int ik = chunk * TPQ_THREADS + threadIdx.x;
        bool active = ik < t.ket.npair && (!t.same_class || ik <= ib);
        if (active && t.thr > 0.0) active = t.bra.Q[ib] * t.ket.Q[ik] > t.thr;

        // bra pair: uniform across the CTA
        const int sa = t.bra.sa[ib], sb = t.bra.sb[ib];
        const double Ax = t.bra.A[3 * ib], Ay = t.bra.A[3 * ib + 1], Az = t.bra.A[3 * ib + 2];
        const double ABx = t.bra.AB[3 * ib], ABy = t.bra.AB[3 * ib + 1], ABz = t.bra.AB[3 * ib + 2];
        const int pab0 = t.bra.pbase[ib], npab = t.bra.nprim[ib];

        int sc = 0, sd = 0, pcd0 = 0, npcd = 0;
        double Cx = 0, Cy = 0, Cz = 0, CDx = 0, CDy = 0, CDz = 0;
        if (active) {
            sc = t.ket.sa[ik]; sd = t.ket.sb[ik];
            Cx = t.ket.A[3 * ik]; Cy = t.ket.A[3 * ik + 1]; Cz = t.ket.A[3 * ik + 2];
            CDx = t.ket.AB[3 * ik]; CDy = t.ket.AB[3 * ik + 1]; CDz = t.ket.AB[3 * ik + 2];
            pcd0 = t.ket.pbase[ik]; npcd = t.ket.nprim[ik];
        }
        double wgt = (sa == sb ? 1.0 : 2.0) * (sc == sd ? 1.0 : 2.0);
        wgt *= (t.same_class && ib == ik) ? 1.0 : 2.0;

        double gout[NOUT];
#pragma unroll
        for (int n = 0; n < NOUT; n++) gout[n] = 0.0;

        for (int b0 = 0; b0 < npab; b0 += TPQ_MAXBP) {
            const int nb = min(TPQ_MAXBP, npab - b0);
            __syncthreads();   // tables staged / previous pass consumed
            if (threadIdx.x < nb) {
                const int s = pab0 + (b0 + threadIdx.x) * CF_PSTRIDE;
                const double px = t.bra.Px[s], py = t.bra.Py[s], pz = t.bra.Pz[s];
                sbra[0 * TPQ_MAXBP + threadIdx.x] = t.bra.p[s];
                sbra[1 * TPQ_MAXBP + threadIdx.x] = t.bra.hp[s];
                sbra[2 * TPQ_MAXBP + threadIdx.x] = px;
                sbra[3 * TPQ_MAXBP + threadIdx.x] = py;
                sbra[4 * TPQ_MAXBP + threadIdx.x] = pz;
                sbra[5 * TPQ_MAXBP + threadIdx.x] = t.bra.c[s];
                sbra[6 * TPQ_MAXBP + threadIdx.x] = px - Ax;
                sbra[7 * TPQ_MAXBP + threadIdx.x] = py - Ay;
                sbra[8 * TPQ_MAXBP + threadIdx.x] = pz - Az;
            }
            __syncthreads();
            if (!active) continue;
            for (int icd = 0; icd < npcd; icd++) {
                const int scd = pcd0 + icd * CF_PSTRIDE;
                const double q = t.ket.p[scd], hq = t.ket.hp[scd], ccd = t.ket.c[scd] * wgt;
                const double Qx = t.ket.Px[scd], Qy = t.ket.Py[scd], Qz = t.ket.Pz[scd];
                const double QCx = Qx - Cx, QCy = Qy - Cy, QCz = Qz - Cz;
                for (int iab = 0; iab < nb; iab++) {
                    const double cc = sbra[5 * TPQ_MAXBP + iab] * ccd;
                    if (fabs(cc) < t.prim_cut) continue;
                    const double p = sbra[iab], hp = sbra[TPQ_MAXBP + iab];
                    const double PQx = sbra[2 * TPQ_MAXBP + iab] - Qx, PQy = sbra[3 * TPQ_MAXBP + iab] - Qy,
                                 PQz = sbra[4 * TPQ_MAXBP + iab] - Qz;
                    const double PAx = sbra[6 * TPQ_MAXBP + iab], PAy = sbra[7 * TPQ_MAXBP + iab], PAz = sbra[8 * TPQ_MAXBP + iab];
                    const double pq = p + q;
                    const double rs = rsqrt(pq), ipq = rs * rs;
                    const double T = (p * q * ipq) * fma(PQx, PQx, fma(PQy, PQy, PQz * PQz));
                    const double pref = cc * rs;
                    double rx[NROOTS], rw[NROOTS];
                    tpq_roots<NROOTS>(tab, T, rx, rw);
                    const double qi = q * ipq, pi_ = p * ipq, hi = 0.5 * ipq;
#pragma unroll
                    for (int r = 0; r < NROOTS; r++) {
                        const double xr = rx[r];
                        const double rxp = xr * qi, rxq = xr * pi_, b00 = xr * hi;
                        const double b10 = fma(-rxp, hp, hp), b01 = fma(-rxq, hq, hq);
                        double gx[GSZ], gy[GSZ], gz[GSZ];
                        rys_2d<LA, LB, LC, LD>(1.0, fma(-rxp, PQx, PAx), fma(rxq, PQx, QCx), b10, b01, b00, ABx, CDx, gx);
                        rys_2d<LA, LB, LC, LD>(1.0, fma(-rxp, PQy, PAy), fma(rxq, PQy, QCy), b10, b01, b00, ABy, CDy, gy);
                        rys_2d<LA, LB, LC, LD>(rw[r] * pref, fma(-rxp, PQz, PAz), fma(rxq, PQz, QCz), b10, b01, b00, ABz, CDz, gz);
#pragma unroll
                        for (int n = 0; n < NOUT; n++) {
                            const int id = n % ND, ic = (n / ND) % NC, jb = (n / (ND * NC)) % NB, ia = n / (ND * NC * NB);
                            const int ix = ((cart_lx(LA, ia) * (LB + 1) + cart_lx(LB, jb)) * (LC + 1) + cart_lx(LC, ic)) * (LD + 1) + cart_lx(LD, id);
                            const int iy = ((cart_ly(LA, ia) * (LB + 1) + cart_ly(LB, jb)) * (LC + 1) + cart_ly(LC, ic)) * (LD + 1) + cart_ly(LD, id);
                            const int iz = ((cart_lz(LA, ia) * (LB + 1) + cart_lz(LB, jb)) * (LC + 1) + cart_lz(LC, ic)) * (LD + 1) + cart_lz(LD, id);
                            gout[n] = fma(gx[ix] * gy[iy], gz[iz], gout[n]);
                        }
                    }
                }
            }
        }
        if (!active) continue;

        // ---- digestion: six blocks, all in registers ------------------------------------------------
        const int ca = t.bra.cao_a[ib], cb = t.bra.cao_b[ib], cc0 = t.ket.cao_a[ik], cd0 = t.ket.cao_b[ik];
        const size_t ld = (size_t)t.ncart;
        {   // J(a,b) += sum_cd V Dtot(c,d) ; J(c,d) += sum_ab V Dtot(a,b)
            double dcd[NCD], jcd[NCD];
#pragma unroll
            for (int kl = 0; kl < NCD; kl++) { dcd[kl] = t.Dtot[(cd0 + kl % ND) * ld + cc0 + kl / ND]; jcd[kl] = 0.0; }
#pragma unroll
            for (int ij = 0; ij < NAB; ij++) {
                const size_t off = (cb + ij % NB) * ld + ca + ij / NB;
                const double dab = t.Dtot[off];
                double s = 0.0;
#pragma unroll
                for (int kl = 0; kl < NCD; kl++) {
                    s = fma(gout[ij * NCD + kl], dcd[kl], s);
                    jcd[kl] = fma(gout[ij * NCD + kl], dab, jcd[kl]);
                }
                fixed_add(t.accJ + off, s, t.scaleJ);
            }
#pragma unroll
            for (int kl = 0; kl < NCD; kl++) fixed_add(t.accJ + (cd0 + kl % ND) * ld + cc0 + kl / ND, jcd[kl], t.scaleJ);
        }
        for (int x = 0; x < t.nk; x++) {
            const double* __restrict__ D = t.Dk[x];
            long long* acc = t.accK[x];
            {   // K(a,c) += sum_bd V D(b,d)
                double d[NB * ND];
#pragma unroll
                for (int e = 0; e < NB * ND; e++) d[e] = D[(cd0 + e % ND) * ld + cb + e / ND];
#pragma unroll
                for (int i = 0; i < NA; i++)
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < NB; j++)
#pragma unroll
                            for (int l = 0; l < ND; l++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[j * ND + l], s);
                        fixed_add(acc + (cc0 + k) * ld + ca + i, s, t.scaleK);
                    }
            }
            {   // K(a,d) += sum_bc V D(b,c)
                double d[NB * NC];
#pragma unroll
                for (int e = 0; e < NB * NC; e++) d[e] = D[(cc0 + e % NC) * ld + cb + e / NC];
#pragma unroll
                for (int i = 0; i < NA; i++)
#pragma unroll
                    for (int l = 0; l < ND; l++) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < NB; j++)
#pragma unroll
                            for (int k = 0; k < NC; k++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[j * NC + k], s);
                        fixed_add(acc + (cd0 + l) * ld + ca + i, s, t.scaleK);
                    }
            }
            {   // K(b,c) += sum_ad V D(a,d)
                double d[NA * ND];
#pragma unroll
                for (int e = 0; e < NA * ND; e++) d[e] = D[(cd0 + e % ND) * ld + ca + e / ND];
#pragma unroll
                for (int j = 0; j < NB; j++)
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        double s = 0.0;
#pragma unroll
                        for (int i = 0; i < NA; i++)
#pragma unroll
                            for (int l = 0; l < ND; l++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[i * ND + l], s);
                        fixed_add(acc + (cc0 + k) * ld + cb + j, s, t.scaleK);
                    }
            }
            {   // K(b,d) += sum_ac V D(a,c)
                double d[NA * NC];
#pragma unroll
                for (int e = 0; e < NA * NC; e++) d[e] = D[(cc0 + e % NC) * ld + ca + e / NC];
#pragma unroll
                for (int j = 0; j < NB; j++)
#pragma unroll
                    for (int l = 0; l < ND; l++) {
                        double s = 0.0;
#pragma unroll
                        for (int i = 0; i < NA; i++)
#pragma unroll
                            for (int k = 0; k < NC; k++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[i * NC + k], s);
                        fixed_add(acc + (cd0 + l) * ld + cb + j, s, t.scaleK);
                    }
            }
        }
    }
}
