// eri_tpq.cuh -- thread-per-quartet Rys kernels with fused J/K digestion (classes with <= TPQ_MAX_NOUT
// Cartesian integrals per shell quartet).
//
// Why this shape (B200): the FP64 pipe retires 64 DFMA/clk/SM while shared memory delivers 16 doubles/clk/SM,
// so any scheme that stages the 2-D Rys integrals in shared memory is LSU-bound by ~4x.  Here one thread owns one
// shell quartet and keeps EVERYTHING in registers: roots, the 2-D recurrences of one root at a time, the
// contracted Cartesian integrals and the digestion partial sums.  All index arithmetic is compile-time.
//   work item  = (one bra shell pair, <= 32 consecutive ket shell pairs), owned by ONE WARP (no CTA barrier in the
//                item loop); bra primitive data is staged per warp in shared memory and read as warp-uniform
//                broadcasts; ket primitives are interleaved in HBM so a warp reads them coalesced
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
#ifndef TPQ_MAX_NOUT
#define TPQ_MAX_NOUT 64  // largest Cartesian block one thread keeps in registers
#endif
#ifndef TPQS_MAX_T
#define TPQS_MAX_T 60    // sliced kernels: largest per-thread share of the block
#endif
#define TPQ_NBRA 9       // p, hp, Px, Py, Pz, c, PAx, PAy, PAz

__host__ __device__ constexpr bool tpq_ok(int la, int lb, int lc, int ld) {
    return cf_ncart(la) * cf_ncart(lb) * cf_ncart(lc) * cf_ncart(ld) <= TPQ_MAX_NOUT;
}
// one-root classes (L <= 1) read F_0 and F_1 from their own Taylor rows F_0..F_9 (no exp, no downward recursion);
// -DCF_BOYS1_EXP restores the F_1-row + exp(-T) + recursion path for A/B measurements
#ifdef CF_BOYS1_EXP
#define BOYS1_COLS 8
#else
#define BOYS1_COLS 10
#endif
// Strides of the STAGED (shared-memory) copies.  Lanes read the same (value, coefficient) position of DIFFERENT rows (their own
// T), so the row / interval stride decides the bank behaviour of every LDS.128: the global layout has 4 chunks of 16 bytes per
// Boys row (2-root classes) and 14 * NROOTS chunks per Chebyshev interval -- an EVEN number, so lanes in different intervals
// reach only 4, 2 or (NROOTS = 4, 8) ONE of the 8 bank groups (ncu r2l, c18 dp|ps: table reads = 96 % of the shared wavefronts
// at 2.7x the ideal count).  The staged copy pads every Boys row to 5 chunks and every interval by one chunk: odd strides,
// all 8 bank groups.  -DCF_TABLE_NOPAD restores the dense copy for A/B runs.
#ifdef CF_TABLE_NOPAD
#define BOYS2_STR 8
__host__ __device__ constexpr int rys_istr(int nroots) { return 2 * nroots * RYS_NC; }
#else
#define BOYS2_STR 10
__host__ __device__ constexpr int rys_istr(int nroots) { return 2 * nroots * RYS_NC + 2; }
#endif
__host__ __device__ constexpr int tpq_table_len(int nroots) {
    return nroots == 1 ? BOYS_NROW * BOYS1_COLS : nroots == 2 ? BOYS_NROW * BOYS2_STR : (rys_tmax(nroots) / 2) * rys_istr(nroots) + 2 * nroots;
}
#define TPQ_WBP 32        // bra primitive pairs staged per pass and per warp by the thread-per-quartet kernel
__host__ __device__ constexpr size_t tpq_smem(int nroots) {
    return sizeof(double) * (size_t)(tpq_table_len(nroots) + (TPQ_THREADS / 32) * TPQ_NBRA * TPQ_WBP);
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
    const double2* row = reinterpret_cast<const double2*>(tab + i * BOYS2_STR);
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

// F_0(T), F_1(T) for T < BOYS_TMAX from the staged rows F_0..F_9 at T_i = i/8: two 8-term Taylor sums sharing the powers
__device__ __forceinline__ void boys01_small(const double* __restrict__ tab, double T, double& F0, double& F1) {
    const int i = (int)fma(T, 8.0, 0.5);
    const double mh = fma((double)i, 0.125, -T);   // -(T - T_i), |mh| <= 1/16
    const double2* row = reinterpret_cast<const double2*>(tab + i * 10);
    const double2 r0 = row[0], r1 = row[1], r2 = row[2], r3 = row[3], r4 = row[4];     // F_0..F_9
    const double h7 = mh * (1.0 / 7.0), h6 = mh * (1.0 / 6.0), h5 = mh * (1.0 / 5.0), h4 = mh * 0.25, h3 = mh * (1.0 / 3.0), h2 = mh * 0.5;
    double s0 = r3.y, s1 = r4.x;
    s0 = fma(s0, h7, r3.x); s1 = fma(s1, h7, r3.y);
    s0 = fma(s0, h6, r2.y); s1 = fma(s1, h6, r3.x);
    s0 = fma(s0, h5, r2.x); s1 = fma(s1, h5, r2.y);
    s0 = fma(s0, h4, r1.y); s1 = fma(s1, h4, r2.x);
    s0 = fma(s0, h3, r1.x); s1 = fma(s1, h3, r1.y);
    s0 = fma(s0, h2, r0.y); s1 = fma(s1, h2, r1.x);
    F0 = fma(s0, mh, r0.x); F1 = fma(s1, mh, r0.y);
}

// stage the root tables of a kernel with NT threads (once per CTA): Boys Taylor rows for 1 and 2 roots, Chebyshev
// coefficient tables + asymptotic constants for 3 and more
template <int NROOTS>
__device__ __forceinline__ void tpq_stage_tables(double* __restrict__ tab, const RysTablesDev& rys, int tid, int nt) {
    if constexpr (NROOTS == 1 && BOYS1_COLS == 10) {
        for (int e = tid; e < BOYS_NROW * 10; e += nt) tab[e] = rys.boys[(e / 10) * BOYS_NCOL + (e % 10)];
    } else if constexpr (NROOTS <= 2) {
        constexpr int M = 2 * NROOTS - 1;
        for (int e = tid; e < BOYS_NROW * 8; e += nt) tab[(e >> 3) * BOYS2_STR + (e & 7)] = rys.boys[(e >> 3) * BOYS_NCOL + M + (e & 7)];
    } else {
        constexpr int PER = 2 * NROOTS * RYS_NC, NTAB = (rys_tmax(NROOTS) / 2) * PER;
        const double* src = rys.table + rys_off(NROOTS);
        for (int e = tid; e < NTAB; e += nt) tab[(e / PER) * rys_istr(NROOTS) + e % PER] = src[e];
        if (tid < 2 * NROOTS) tab[(rys_tmax(NROOTS) / 2) * rys_istr(NROOTS) + tid] = rys.asym[rys_asym_off(NROOTS) + tid];
    }
}

// Rys roots x_r (= t^2) and weights (sum_r w_r x_r^k = F_k(T)) of one primitive quartet, all in registers
template <int NROOTS>
__device__ __forceinline__ void tpq_roots(const double* __restrict__ tab, double T, double* __restrict__ x, double* __restrict__ w) {
    if constexpr (NROOTS == 1 && BOYS1_COLS == 10) {
        double F0, F1;
        if (T < BOYS_TMAX) boys01_small(tab, T, F0, F1);
        else {   // erfc(sqrt T) < 2e-17: F_0 = sqrt(pi/4T); F_1 = (F_0 - e^-T) / 2T with e^-T / F_0 < 2e-15 dropped (root only)
            const double r = rsqrt(T);
            F0 = 0.88622692545275801365 * r;
            F1 = F0 * (0.5 * r * r);
        }
        w[0] = F0;
        x[0] = F1 * fast_rcp(F0);
    } else if constexpr (NROOTS <= 2) {
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
            const double* asym = tab + (rys_tmax(NROOTS) / 2) * rys_istr(NROOTS);
            const double rs = rsqrt(T), it = rs * rs;
#pragma unroll
            for (int v = 0; v < NROOTS; v++) { x[v] = asym[v] * it; w[v] = asym[NROOTS + v] * rs; }
        } else {
            const int it = (int)(T * 0.5);
            const double u = T - (2.0 * it + 1.0), u2 = u + u;
            const double2* cs = reinterpret_cast<const double2*>(tab + (size_t)it * rys_istr(NROOTS));
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

// sum over the 32 lanes in a FIXED butterfly order (deterministic for a given work item); every lane gets the total
__device__ __forceinline__ double warp_sum_fixed(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums of N per-lane values over the 32 lanes by RECURSIVE HALVING: at distance 16, 8, 4, 2, 1 a lane keeps one half of
// its values and adds the partner's copy of that half, so N values cost ~N double shuffles instead of 5 N (ncu r2l: the 18
// warp_sum_fixed calls of J(a,b) were 24 % of the samples of the c18 dp|ps kernel).  The summation tree of every element
// is the butterfly of warp_sum_fixed (own + partner at each distance; addition commutes): results are bit-identical to it.
// On return the lane holds out[i] = total of element `first + i` for i < count (count may be 0).
#ifndef CF_NO_WARP_MULTI_SUM
#define CF_WARP_MULTI_SUM 1
#endif
__host__ __device__ constexpr int wms_half(int n) { return (n + 1) / 2; }
__host__ __device__ constexpr int wms_len(int n) { return wms_half(wms_half(wms_half(wms_half(wms_half(n))))); }
template <int N, int O>
__device__ __forceinline__ void wms_step(const double (&v)[N], double (&w)[(N + 1) / 2], int lane, int& first, int& count) {
    constexpr int H = (N + 1) / 2;
    const bool up = (lane & O) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
        const double lo = v[i];
        const double hi = (i + H < N) ? v[i + H] : 0.0;
        const double keep = up ? hi : lo, send = up ? lo : hi;
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
    }
    if (up) { first += H; count -= H; }
    count = max(0, min(count, H));
}
template <int N>
__device__ __forceinline__ void warp_multi_sum(const double (&v)[N], double (&out)[wms_len(N)], int lane, int& first, int& count) {
    constexpr int N1 = wms_half(N), N2 = wms_half(N1), N3 = wms_half(N2), N4 = wms_half(N3);
    double a1[N1], a2[N2], a3[N3], a4[N4];
    first = 0; count = N;
    wms_step<N, 16>(v, a1, lane, first, count);
    wms_step<N1, 8>(a1, a2, lane, first, count);
    wms_step<N2, 4>(a2, a3, lane, first, count);
    wms_step<N3, 2>(a3, a4, lane, first, count);
    wms_step<N4, 1>(a4, out, lane, first, count);
}

#ifndef TPQ_MINB
#define TPQ_MINB 2
#endif
// resident CTAs the register allocation is held to: the small classes sit right at the 128-register boundary
__host__ __device__ constexpr int tpq_minb(int nout) { return nout <= 9 ? 4 : TPQ_MINB; }

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(TPQ_THREADS, tpq_minb(cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD))) eri_jk_tpq(const QuartetTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND, NOUT = NAB * NCD;
    constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    constexpr int GSZ = (LA + 1) * (LB + 1) * (LC + 1) * (LD + 1);
    constexpr int TABLEN = tpq_table_len(NROOTS);
    constexpr int MAXBP = TPQ_WBP;
    extern __shared__ double smem[];
    const double scaleJ = __ldg(t.scales), scaleK = __ldg(t.scales + 1);
    const double thr = __ldg(t.scales + 4);   // effective Schwarz threshold of this build (scales_kernel)
    const long long jlo = __ldg(t.scales + 6) != 0.0 ? t.jlo_off : 0;
    unsigned cnt_q = 0, cnt_p = 0;            // this lane's evaluated shell quartets / executed primitive quartets
    double* tab = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sbra = smem + TABLEN + warp * (TPQ_NBRA * MAXBP);   // this warp's [TPQ_NBRA][MAXBP]

    // ---- stage the root tables (once per CTA) ----------------------------------------------------
    tpq_stage_tables<NROOTS>(tab, t.rys, threadIdx.x, TPQ_THREADS);
    __syncthreads();

    // ---- work items are WARP-private: (one bra pair) x (<= 32 consecutive ket pairs); no CTA barrier below -------
    const long long nitem_local = (t.nitem - t.rank + t.world - 1) / t.world;
    const long long gw = (long long)blockIdx.x * (TPQ_THREADS / 32) + warp, nw = (long long)gridDim.x * (TPQ_THREADS / 32);
    int4 it_next = make_int4(0, 0, 0, 0);
    if (gw < nitem_local) it_next = __ldg(t.items + gw * t.world + t.rank);
    for (long long li = gw; li < nitem_local; li += nw) {
        const int4 it = it_next;                      // descriptor of the next item is fetched one item ahead
        if (li + nw < nitem_local) it_next = __ldg(t.items + (li + nw) * t.world + t.rank);
        const int ib = it.x;
        const int ik = it.y + lane;
        bool active = lane < it.z;
        if (active && thr > 0.0) active = t.bra.Q[ib] * t.ket.Q[ik] > thr;
        {   // warp-uniform: skip items whose quartets are all screened out; count the evaluated ones
            const unsigned amask = __ballot_sync(0xffffffffu, active);
            if (!amask) continue;
            cnt_q += active ? 1u : 0u;
        }

        // bra pair: uniform across the warp
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
#ifdef CF_PREFETCH_D
        if (active) {   // lines of the D elements read by the digestion -> L1 while the integrals are computed
            const int pca = t.bra.cao_a[ib], pcb = t.bra.cao_b[ib], pcc = t.ket.cao_a[ik], pcd = t.ket.cao_b[ik];
            const size_t pld = (size_t)t.ncart;
#pragma unroll
            for (int l = 0; l < ND; l++) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(t.Dtot + (pcd + l) * pld + pcc));
                for (int x = 0; x < t.nk; x++) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(t.Dk[x] + (pcd + l) * pld + pca));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(t.Dk[x] + (pcd + l) * pld + pcb));
                }
            }
#pragma unroll
            for (int k = 0; k < NC; k++)
                for (int x = 0; x < t.nk; x++) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(t.Dk[x] + (pcc + k) * pld + pca));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(t.Dk[x] + (pcc + k) * pld + pcb));
                }
        }
#endif

        double gout[NOUT];
#pragma unroll
        for (int n = 0; n < NOUT; n++) gout[n] = 0.0;

        for (int b0 = 0; b0 < npab; b0 += MAXBP) {
            const int nb = min(MAXBP, npab - b0);
            __syncwarp();      // previous pass consumed
            if (lane < nb) {
                const int s = pab0 + (b0 + lane) * CF_PSTRIDE;
                const double px = t.bra.Px[s], py = t.bra.Py[s], pz = t.bra.Pz[s];
                sbra[0 * MAXBP + lane] = t.bra.p[s];
                sbra[1 * MAXBP + lane] = t.bra.hp[s];
                sbra[2 * MAXBP + lane] = px;
                sbra[3 * MAXBP + lane] = py;
                sbra[4 * MAXBP + lane] = pz;
                sbra[5 * MAXBP + lane] = t.bra.c[s];
                sbra[6 * MAXBP + lane] = px - Ax;
                sbra[7 * MAXBP + lane] = py - Ay;
                sbra[8 * MAXBP + lane] = pz - Az;
            }
            __syncwarp();
            if (!active) continue;
            for (int icd = 0; icd < npcd; icd++) {
                const int scd = pcd0 + icd * CF_PSTRIDE;
                const double q = t.ket.p[scd], hq = t.ket.hp[scd], ccd = t.ket.c[scd] * wgt;
                const double Qx = t.ket.Px[scd], Qy = t.ket.Py[scd], Qz = t.ket.Pz[scd];
                const double QCx = Qx - Cx, QCy = Qy - Cy, QCz = Qz - Cz;
                for (int iab = 0; iab < nb; iab++) {
                    const double cc = sbra[5 * MAXBP + iab] * ccd;
                    if (fabs(cc) < t.prim_cut) continue;
                    cnt_p++;
                    const double p = sbra[iab], hp = sbra[MAXBP + iab];
                    const double PQx = sbra[2 * MAXBP + iab] - Qx, PQy = sbra[3 * MAXBP + iab] - Qy,
                                 PQz = sbra[4 * MAXBP + iab] - Qz;
                    const double PAx = sbra[6 * MAXBP + iab], PAy = sbra[7 * MAXBP + iab], PAz = sbra[8 * MAXBP + iab];
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
        // ---- digestion: six blocks, all in registers ------------------------------------------------
        // (inactive lanes hold gout == 0 and take part in the warp-wide J(a,b) sums only)
        const int ca = t.bra.cao_a[ib], cb = t.bra.cao_b[ib];
        int cc0 = 0, cd0 = 0;
        if (active) { cc0 = t.ket.cao_a[ik]; cd0 = t.ket.cao_b[ik]; }
        const size_t ld = (size_t)t.ncart;
        for (int xj = 0; xj < t.nj; xj++) {   // J(a,b) += sum_cd V Dj(c,d) ; J(c,d) += sum_ab V Dj(a,b)
            const double* __restrict__ DJ = t.Dj[xj];
            long long* aJ = t.accJm[xj];
            double dcd[NCD], jcd[NCD];
#pragma unroll
            for (int kl = 0; kl < NCD; kl++) { dcd[kl] = active ? DJ[(cd0 + kl % ND) * ld + cc0 + kl / ND] : 0.0; jcd[kl] = 0.0; }
#ifdef CF_WARP_MULTI_SUM
            double sab[NAB];
#endif
#pragma unroll
            for (int ij = 0; ij < NAB; ij++) {
                const size_t off = (cb + ij % NB) * ld + ca + ij / NB;
                const double dab = DJ[off];
                double s = 0.0;
#pragma unroll
                for (int kl = 0; kl < NCD; kl++) {
                    s = fma(gout[ij * NCD + kl], dcd[kl], s);
                    jcd[kl] = fma(gout[ij * NCD + kl], dab, jcd[kl]);
                }
                // the bra pair is common to the warp's 32 quartets: one add per element and warp instead of 32
#ifdef CF_WARP_MULTI_SUM
                sab[ij] = s;
#else
                s = warp_sum_fixed(s);
                if (lane == 0) fixed_add_j(aJ + off, jlo, s, scaleJ);
#endif
            }
#ifdef CF_WARP_MULTI_SUM
            {   // all NAB warp sums at once; the lanes that end up holding totals add them (one atomic instruction per slot)
                double tot[wms_len(NAB)];
                int first, count;
                warp_multi_sum<NAB>(sab, tot, lane, first, count);
#pragma unroll
                for (int i = 0; i < wms_len(NAB); i++)
                    if (i < count) {
                        const int ij = first + i;
                        fixed_add_j(aJ + (size_t)(cb + ij % NB) * ld + ca + ij / NB, jlo, tot[i], scaleJ);
                    }
            }
#endif
            if (active) {
#pragma unroll
                for (int kl = 0; kl < NCD; kl++) fixed_add_j(aJ + (cd0 + kl % ND) * ld + cc0 + kl / ND, jlo, jcd[kl], scaleJ);
            }
        }
        if (!active) continue;
        for (int x = 0; x < t.nk; x++) {
            const double* __restrict__ D = t.Dk[x];
            long long* acc = t.accK[x];
            {   // K(a,c) += sum_bd V D(b,d)
                double d[NB * ND];
#pragma unroll
                for (int e = 0; e < NB * ND; e++) d[e] = D[(cb + e / ND) * ld + cd0 + e % ND];
#pragma unroll
                for (int i = 0; i < NA; i++)
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < NB; j++)
#pragma unroll
                            for (int l = 0; l < ND; l++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[j * ND + l], s);
                        fixed_add(acc + (ca + i) * ld + cc0 + k, s, scaleK);
                    }
            }
            {   // K(a,d) += sum_bc V D(b,c)
                double d[NB * NC];
#pragma unroll
                for (int e = 0; e < NB * NC; e++) d[e] = D[(cb + e / NC) * ld + cc0 + e % NC];
#pragma unroll
                for (int i = 0; i < NA; i++)
#pragma unroll
                    for (int l = 0; l < ND; l++) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < NB; j++)
#pragma unroll
                            for (int k = 0; k < NC; k++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[j * NC + k], s);
                        fixed_add(acc + (ca + i) * ld + cd0 + l, s, scaleK);
                    }
            }
            {   // K(b,c) += sum_ad V D(a,d)
                double d[NA * ND];
#pragma unroll
                for (int e = 0; e < NA * ND; e++) d[e] = D[(ca + e / ND) * ld + cd0 + e % ND];
#pragma unroll
                for (int j = 0; j < NB; j++)
#pragma unroll
                    for (int k = 0; k < NC; k++) {
                        double s = 0.0;
#pragma unroll
                        for (int i = 0; i < NA; i++)
#pragma unroll
                            for (int l = 0; l < ND; l++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[i * ND + l], s);
                        fixed_add(acc + (cb + j) * ld + cc0 + k, s, scaleK);
                    }
            }
            {   // K(b,d) += sum_ac V D(a,c)
                double d[NA * NC];
#pragma unroll
                for (int e = 0; e < NA * NC; e++) d[e] = D[(ca + e / NC) * ld + cc0 + e % NC];
#pragma unroll
                for (int j = 0; j < NB; j++)
#pragma unroll
                    for (int l = 0; l < ND; l++) {
                        double s = 0.0;
#pragma unroll
                        for (int i = 0; i < NA; i++)
#pragma unroll
                            for (int k = 0; k < NC; k++) s = fma(gout[((i * NB + j) * NC + k) * ND + l], d[i * NC + k], s);
                        fixed_add(acc + (cb + j) * ld + cd0 + l, s, scaleK);
                    }
            }
        }
    }
    cf_cnt_flush(t.cnt, cnt_q, cnt_p);
}

// ================================================================================================
// Sliced variant for larger classes: GS threads (in GS different warps of the CTA) share one quartet.
// Thread `s` owns the MA = NA/GS Cartesian components ia = s*MA .. s*MA+MA-1 of shell a and ALL components of
// b, c, d.  Roots, the VRR and the bra transfer are recomputed by each of the GS threads (FP64 is the cheap
// resource; shared memory is not), the a-exponent of an owned component is picked from the register arrays by
// select chains, so ONE instruction stream serves every slice.  J(a,b), K(a,c), K(a,d) are complete per thread;
// the J(c,d), K(b,c), K(b,d) partial sums are combined across the GS threads through shared memory (fixed
// order) before the fixed-point atomics.
// ================================================================================================
#define TPQS_PART_BYTES (40 * 1024)

__host__ __device__ constexpr int tpqs_nq(int gs) { return gs <= 2 ? 64 : 32; }
__host__ __device__ constexpr int tpqs_vc(int gs) { return TPQS_PART_BYTES / (gs * tpqs_nq(gs) * 8); }
__host__ __device__ constexpr size_t tpqs_smem(int nroots, int gs) {
    return sizeof(double) * (size_t)(tpq_table_len(nroots) + TPQ_NBRA * TPQ_MAXBP + tpqs_vc(gs) * gs * tpqs_nq(gs));
}
// slices per quartet for a class (0: not covered by the sliced kernels); per-thread outputs <= 60
__host__ __device__ constexpr int tpqs_gs(int la, int lb, int lc, int ld) {
    const int na = cf_ncart(la), rest = cf_ncart(lb) * cf_ncart(lc) * cf_ncart(ld);
    if (na * rest <= TPQ_MAX_NOUT) return 0;       // plain thread-per-quartet
    for (int gs = 2; gs <= na; gs++)
        if (na % gs == 0 && (na / gs) * rest <= TPQS_MAX_T) return gs;
    return 0;
}

template <int N>
__device__ __forceinline__ double sel_reg(const double* v, int idx) {
    double r = v[0];
#pragma unroll
    for (int i = 1; i < N; i++) r = (idx == i) ? v[i] : r;
    return r;
}

// VRR + bra transfer of one root and one Cartesian direction: b[j][i][k] = [i j | k 0], k <= LC+LD
template <int LA, int LB, int LCD>
__device__ __forceinline__ void rys_2d_bra(double w0, double c00, double c00p, double b10, double b01, double b00, double ab,
                                           double (&b)[LB + 1][LA + 1][LCD + 1]) {
    constexpr int LAB = LA + LB;
    double a[LAB + 1][LCD + 1];
    a[0][0] = w0;
    if (LAB > 0) a[1][0] = c00 * w0;
#pragma unroll
    for (int i = 1; i < LAB; i++) a[i + 1][0] = fma(c00, a[i][0], (i * b10) * a[i - 1][0]);
#pragma unroll
    for (int k = 0; k < LCD; k++) {
#pragma unroll
        for (int i = 0; i <= LAB; i++) {
            double v = c00p * a[i][k];
            if (k > 0) v = fma(k * b01, a[i][k - 1], v);
            if (i > 0) v = fma(i * b00, a[i - 1][k], v);
            a[i][k + 1] = v;
        }
    }
#pragma unroll
    for (int k = 0; k <= LCD; k++) {
        double h[LAB + 1];
#pragma unroll
        for (int i = 0; i <= LAB; i++) h[i] = a[i][k];
#pragma unroll
        for (int i = 0; i <= LA; i++) b[0][i][k] = h[i];
#pragma unroll
        for (int j = 1; j <= LB; j++) {
#pragma unroll
            for (int i = 0; i <= LAB - j; i++) h[i] = fma(ab, h[i], h[i + 1]);
#pragma unroll
            for (int i = 0; i <= LA; i++) b[j][i][k] = h[i];
        }
    }
}

// pick a-exponent `ai` (runtime), then the ket transfer: g[j][k][l] = [ai j | k l]
template <int LA, int LB, int LC, int LD>
__device__ __forceinline__ void rys_2d_ket(const double (&b)[LB + 1][LA + 1][LC + LD + 1], int ai, double cd,
                                           double (&g)[LB + 1][LC + 1][LD + 1]) {
    constexpr int LCD = LC + LD;
#pragma unroll
    for (int j = 0; j <= LB; j++) {
        double c[LCD + 1];
#pragma unroll
        for (int k = 0; k <= LCD; k++) {
            double cand[LA + 1];
#pragma unroll
            for (int i = 0; i <= LA; i++) cand[i] = b[j][i][k];
            c[k] = sel_reg<LA + 1>(cand, ai);
        }
#pragma unroll
        for (int l = 0; l <= LD; l++) {
            if (l > 0) {
#pragma unroll
                for (int k = 0; k <= LCD - l; k++) c[k] = fma(cd, c[k], c[k + 1]);
            }
#pragma unroll
            for (int k = 0; k <= LC; k++) g[j][k][l] = c[k];
        }
    }
}

#ifndef TPQS_MINB
#define TPQS_MINB(nt) 1
#endif
template <int LA, int LB, int LC, int LD, int GS>
__global__ void __launch_bounds__(GS * tpqs_nq(GS), TPQS_MINB(GS * tpqs_nq(GS))) eri_jk_tpqs(const QuartetTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NCD = NC * ND, NBCD = NB * NCD;
    constexpr int MA = NA / GS, NOUT_T = MA * NBCD;
    constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    constexpr int LCD = LC + LD;
    constexpr int TABLEN = tpq_table_len(NROOTS);
    constexpr int NQ = tpqs_nq(GS), NT = GS * NQ, VC = tpqs_vc(GS);
    static_assert(NA % GS == 0, "slices must divide the components of shell a");
    extern __shared__ double smem[];
    const double scaleJ = __ldg(t.scales), scaleK = __ldg(t.scales + 1);
    const double thr = __ldg(t.scales + 4);   // effective Schwarz threshold of this build (scales_kernel)
    const long long jlo = __ldg(t.scales + 6) != 0.0 ? t.jlo_off : 0;
    unsigned cnt_q = 0, cnt_p = 0;            // counted by slice 0 of every quartet
    double* tab = smem;
    double* sbra = smem + TABLEN;                   // [TPQ_NBRA][TPQ_MAXBP]
    double* part = sbra + TPQ_NBRA * TPQ_MAXBP;     // [VC][GS][NQ]

    tpq_stage_tables<NROOTS>(tab, t.rys, threadIdx.x, NT);

    const int q = threadIdx.x % NQ, s = threadIdx.x / NQ;   // s is warp-uniform
    int eax[MA], eay[MA], eaz[MA];
#pragma unroll
    for (int m = 0; m < MA; m++) {
        const int ia = s * MA + m;
        eax[m] = cart_lx(LA, ia); eay[m] = cart_ly(LA, ia); eaz[m] = cart_lz(LA, ia);
    }

    const long long nitem_local = (t.nitem - t.rank + t.world - 1) / t.world;
    int4 it_next = make_int4(0, 0, 0, 0);
    if ((long long)blockIdx.x < nitem_local) it_next = __ldg(t.items + (long long)blockIdx.x * t.world + t.rank);
    for (long long li = blockIdx.x; li < nitem_local; li += gridDim.x) {
        const int4 it = it_next;                      // descriptor of the next item is fetched one item ahead
        if (li + gridDim.x < nitem_local) it_next = __ldg(t.items + (li + gridDim.x) * t.world + t.rank);
        const int ib = it.x;
        const int ik = it.y + q;
        bool active = q < it.z;
        if (active && thr > 0.0) active = t.bra.Q[ib] * t.ket.Q[ik] > thr;
        if (s == 0 && active) cnt_q++;   // slice 0 of every quartet counts

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

        double gout[NOUT_T];
#pragma unroll
        for (int n = 0; n < NOUT_T; n++) gout[n] = 0.0;

        for (int b0 = 0; b0 < npab; b0 += TPQ_MAXBP) {
            const int nb = min(TPQ_MAXBP, npab - b0);
            __syncthreads();
            if (threadIdx.x < nb) {
                const int sl = pab0 + (b0 + threadIdx.x) * CF_PSTRIDE;
                const double px = t.bra.Px[sl], py = t.bra.Py[sl], pz = t.bra.Pz[sl];
                sbra[0 * TPQ_MAXBP + threadIdx.x] = t.bra.p[sl];
                sbra[1 * TPQ_MAXBP + threadIdx.x] = t.bra.hp[sl];
                sbra[2 * TPQ_MAXBP + threadIdx.x] = px;
                sbra[3 * TPQ_MAXBP + threadIdx.x] = py;
                sbra[4 * TPQ_MAXBP + threadIdx.x] = pz;
                sbra[5 * TPQ_MAXBP + threadIdx.x] = t.bra.c[sl];
                sbra[6 * TPQ_MAXBP + threadIdx.x] = px - Ax;
                sbra[7 * TPQ_MAXBP + threadIdx.x] = py - Ay;
                sbra[8 * TPQ_MAXBP + threadIdx.x] = pz - Az;
            }
            __syncthreads();
            if (!active) continue;
            for (int icd = 0; icd < npcd; icd++) {
                const int scd = pcd0 + icd * CF_PSTRIDE;
                const double qe = t.ket.p[scd], hq = t.ket.hp[scd], ccd = t.ket.c[scd] * wgt;
                const double Qx = t.ket.Px[scd], Qy = t.ket.Py[scd], Qz = t.ket.Pz[scd];
                const double QCx = Qx - Cx, QCy = Qy - Cy, QCz = Qz - Cz;
                for (int iab = 0; iab < nb; iab++) {
                    const double cc = sbra[5 * TPQ_MAXBP + iab] * ccd;
                    if (fabs(cc) < t.prim_cut) continue;
                    if (s == 0) cnt_p++;
                    const double p = sbra[iab], hp = sbra[TPQ_MAXBP + iab];
                    const double PQx = sbra[2 * TPQ_MAXBP + iab] - Qx, PQy = sbra[3 * TPQ_MAXBP + iab] - Qy,
                                 PQz = sbra[4 * TPQ_MAXBP + iab] - Qz;
                    const double PAx = sbra[6 * TPQ_MAXBP + iab], PAy = sbra[7 * TPQ_MAXBP + iab], PAz = sbra[8 * TPQ_MAXBP + iab];
                    const double pq = p + qe;
                    const double rs = rsqrt(pq), ipq = rs * rs;
                    const double T = (p * qe * ipq) * fma(PQx, PQx, fma(PQy, PQy, PQz * PQz));
                    const double pref = cc * rs;
                    double rx[NROOTS], rw[NROOTS];
                    tpq_roots<NROOTS>(tab, T, rx, rw);
                    const double qi = qe * ipq, pi_ = p * ipq, hi = 0.5 * ipq;
#pragma unroll
                    for (int r = 0; r < NROOTS; r++) {
                        const double xr = rx[r];
                        const double rxp = xr * qi, rxq = xr * pi_, b00 = xr * hi;
                        const double b10 = fma(-rxp, hp, hp), b01 = fma(-rxq, hq, hq);
                        double bx[LB + 1][LA + 1][LCD + 1], by[LB + 1][LA + 1][LCD + 1], bz[LB + 1][LA + 1][LCD + 1];
                        rys_2d_bra<LA, LB, LCD>(1.0, fma(-rxp, PQx, PAx), fma(rxq, PQx, QCx), b10, b01, b00, ABx, bx);
                        rys_2d_bra<LA, LB, LCD>(1.0, fma(-rxp, PQy, PAy), fma(rxq, PQy, QCy), b10, b01, b00, ABy, by);
                        rys_2d_bra<LA, LB, LCD>(rw[r] * pref, fma(-rxp, PQz, PAz), fma(rxq, PQz, QCz), b10, b01, b00, ABz, bz);
#pragma unroll
                        for (int m = 0; m < MA; m++) {
                            double gx[LB + 1][LC + 1][LD + 1], gy[LB + 1][LC + 1][LD + 1], gz[LB + 1][LC + 1][LD + 1];
                            rys_2d_ket<LA, LB, LC, LD>(bx, eax[m], CDx, gx);
                            rys_2d_ket<LA, LB, LC, LD>(by, eay[m], CDy, gy);
                            rys_2d_ket<LA, LB, LC, LD>(bz, eaz[m], CDz, gz);
#pragma unroll
                            for (int n = 0; n < NBCD; n++) {
                                const int id = n % ND, ic = (n / ND) % NC, jb = n / NCD;
                                gout[m * NBCD + n] = fma(gx[cart_lx(LB, jb)][cart_lx(LC, ic)][cart_lx(LD, id)] *
                                                             gy[cart_ly(LB, jb)][cart_ly(LC, ic)][cart_ly(LD, id)],
                                                         gz[cart_lz(LB, jb)][cart_lz(LC, ic)][cart_lz(LD, id)], gout[m * NBCD + n]);
                            }
                        }
                    }
                }
            }
        }

        // ---- digestion ----------------------------------------------------------------------------
        int ca = 0, cb = 0, cc0 = 0, cd0 = 0;
        ca = t.bra.cao_a[ib]; cb = t.bra.cao_b[ib];          // bra pair: valid for every lane (warp-wide J(a,b) sums)
        if (active) { cc0 = t.ket.cao_a[ik]; cd0 = t.ket.cao_b[ik]; }
        const size_t ld = (size_t)t.ncart;
        const int ia0 = s * MA;
        // cross-slice sum of NV partial values through shared memory, chunk by chunk, then one fixed-point add each
        auto reduce_add = [&](auto nv_tag, const double* pv, auto&& addr_of, long long* acc, double scale, long long lo) {
            constexpr int NV = decltype(nv_tag)::value;
#pragma unroll
            for (int c0 = 0; c0 < NV; c0 += VC) {
                constexpr int dummy = 0; (void)dummy;
                const int cn = (NV - c0) < VC ? (NV - c0) : VC;
                __syncthreads();
#pragma unroll
                for (int v = 0; v < VC; v++)
                    if (c0 + v < NV) part[(v * GS + s) * NQ + q] = pv[c0 + v];
                __syncthreads();
                if (active)
                    for (int v = s; v < cn; v += GS) {
                        double sum = 0.0;
#pragma unroll
                        for (int s2 = 0; s2 < GS; s2++) sum += part[(v * GS + s2) * NQ + q];
                        fixed_add_j(acc + addr_of(c0 + v), lo, sum, scale);
                    }
            }
        };
        for (int xj = 0; xj < t.nj; xj++) {   // J(a,b) complete; J(c,d) partial over the owned a components
            const double* __restrict__ DJ = t.Dj[xj];
            long long* aJ = t.accJm[xj];
            double dcd[NCD], jcd[NCD];
#pragma unroll
            for (int kl = 0; kl < NCD; kl++) { dcd[kl] = active ? DJ[(cd0 + kl % ND) * ld + cc0 + kl / ND] : 0.0; jcd[kl] = 0.0; }
#ifdef CF_WARP_MULTI_SUM
            double sab[MA * NB];
#endif
#pragma unroll
            for (int m = 0; m < MA; m++)
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    const size_t off = (cb + j) * ld + ca + ia0 + m;
                    const double dab = active ? DJ[off] : 0.0;
                    double sum = 0.0;
#pragma unroll
                    for (int kl = 0; kl < NCD; kl++) {
                        sum = fma(gout[(m * NB + j) * NCD + kl], dcd[kl], sum);
                        jcd[kl] = fma(gout[(m * NB + j) * NCD + kl], dab, jcd[kl]);
                    }
                    // the bra pair and the slice are warp-uniform: one add per element and warp
#ifdef CF_WARP_MULTI_SUM
                    sab[m * NB + j] = sum;
#else
                    sum = warp_sum_fixed(sum);
                    if ((threadIdx.x & 31) == 0) fixed_add_j(aJ + off, jlo, sum, scaleJ);
#endif
                }
#ifdef CF_WARP_MULTI_SUM
            {
                double tot[wms_len(MA * NB)];
                int first, count;
                warp_multi_sum<MA * NB>(sab, tot, threadIdx.x & 31, first, count);
#pragma unroll
                for (int i = 0; i < wms_len(MA * NB); i++)
                    if (i < count) {
                        const int e = first + i;
                        fixed_add_j(aJ + (size_t)(cb + e % NB) * ld + ca + ia0 + e / NB, jlo, tot[i], scaleJ);
                    }
            }
#endif
            reduce_add(std::integral_constant<int, NCD>{}, jcd,
                       [&](int kl) { return (size_t)(cd0 + kl % ND) * ld + cc0 + kl / ND; }, aJ, scaleJ, jlo);
        }
        for (int x = 0; x < t.nk; x++) {
            const double* __restrict__ D = t.Dk[x];
            long long* acc = t.accK[x];
            if (active) {
                {   // K(a,c) += sum_bd V D(b,d)   (complete)
                    double d[NB * ND];
#pragma unroll
                    for (int e = 0; e < NB * ND; e++) d[e] = D[(cb + e / ND) * ld + cd0 + e % ND];
#pragma unroll
                    for (int m = 0; m < MA; m++)
#pragma unroll
                        for (int k = 0; k < NC; k++) {
                            double sum = 0.0;
#pragma unroll
                            for (int j = 0; j < NB; j++)
#pragma unroll
                                for (int l = 0; l < ND; l++) sum = fma(gout[((m * NB + j) * NC + k) * ND + l], d[j * ND + l], sum);
                            fixed_add(acc + (ca + ia0 + m) * ld + cc0 + k, sum, scaleK);
                        }
                }
                {   // K(a,d) += sum_bc V D(b,c)   (complete)
                    double d[NB * NC];
#pragma unroll
                    for (int e = 0; e < NB * NC; e++) d[e] = D[(cb + e / NC) * ld + cc0 + e % NC];
#pragma unroll
                    for (int m = 0; m < MA; m++)
#pragma unroll
                        for (int l = 0; l < ND; l++) {
                            double sum = 0.0;
#pragma unroll
                            for (int j = 0; j < NB; j++)
#pragma unroll
                                for (int k = 0; k < NC; k++) sum = fma(gout[((m * NB + j) * NC + k) * ND + l], d[j * NC + k], sum);
                            fixed_add(acc + (ca + ia0 + m) * ld + cd0 + l, sum, scaleK);
                        }
                }
            }
            // K(b,c) += sum_ad V D(a,d) ; K(b,d) += sum_ac V D(a,c)   (partial over the owned a components)
            double kp[NB * NC + NB * ND];
#pragma unroll
            for (int e = 0; e < NB * NC + NB * ND; e++) kp[e] = 0.0;
#pragma unroll
            for (int m = 0; m < MA; m++) {
                double dad[ND], dac[NC];
#pragma unroll
                for (int l = 0; l < ND; l++) dad[l] = active ? D[(ca + ia0 + m) * ld + cd0 + l] : 0.0;
#pragma unroll
                for (int k = 0; k < NC; k++) dac[k] = active ? D[(ca + ia0 + m) * ld + cc0 + k] : 0.0;
#pragma unroll
                for (int j = 0; j < NB; j++)
#pragma unroll
                    for (int k = 0; k < NC; k++)
#pragma unroll
                        for (int l = 0; l < ND; l++) {
                            const double v = gout[((m * NB + j) * NC + k) * ND + l];
                            kp[j * NC + k] = fma(v, dad[l], kp[j * NC + k]);
                            kp[NB * NC + j * ND + l] = fma(v, dac[k], kp[NB * NC + j * ND + l]);
                        }
            }
            reduce_add(std::integral_constant<int, NB * NC + NB * ND>{}, kp,
                       [&](int e) {
                           if (e < NB * NC) return (size_t)(cb + e / NC) * ld + cc0 + e % NC;
                           const int f = e - NB * NC;
                           return (size_t)(cb + f / ND) * ld + cd0 + f % ND;
                       }, acc, scaleK, 0LL);
        }
    }
    cf_cnt_flush(t.cnt, cnt_q, cnt_p);
}
