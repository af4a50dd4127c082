// eri_wg.cuh -- warp-group cooperative Rys kernels with fused J/K digestion for the large classes.
//
// A group of GS lanes of one warp owns one shell quartet (QW = 32/GS quartets per warp).  Notation: H = the
// CTA-uniform ("held") shell pair (a,b): every lane keeps ALL NA*NB Cartesian components of it as register
// accumulators; S = the per-quartet ("spread") shell pair (c,d): its NC*ND components are spread over the lanes of
// the group, MK per lane.  The engine may hand the classes over in either order (integrals are symmetric in
// (ab|cd) <-> (cd|ab)), so H is whichever pair gives the better register/lane fit.
// Per primitive quartet the group works in three warp-synchronous phases (no CTA barrier in the primitive loop):
//   A  2*NROOTS root/weight values, spread over the lanes (Chebyshev tables staged in shared memory)
//   B  3*NROOTS (root, direction) tasks spread over the lanes: VRR + S-side transfer -> shared memory,
//      laid out [c-exp][d-exp][i = 0..LA+LB] so that phase C reads contiguous columns
//   C  every lane: per root and owned S component, 3 columns from shared memory, H-side transfer in registers,
//      then 2*NA*NB DFMAs of assembly with compile-time indices (>= 4 DFMA per shared-memory double, which is
//      the FP64 : shared-memory balance point of the SM)
// Digestion: J(c,d) is complete per lane; J(a,b) and the four K blocks are combined across the lanes of the group
// through the same shared-memory scratch (fixed order), then added as 64-bit fixed point.
// Replaces libint2's engine.compute + the reference's stored-integral digestion
// (src/Integral/Int4C2E.cpp:233-302 and :601-671).
#pragma once
#include "eri_tpq.cuh"

#define WG_WARPS 4

// bring the 128-byte line of a global address into L1 (the digestion reads it ~10^4 cycles later)
#ifndef CF_HAVE_PREFETCH_L1
#define CF_HAVE_PREFETCH_L1
__device__ __forceinline__ void cf_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

template <int B, int E, class F>
__device__ __forceinline__ void wg_static_for(F&& f) {
    if constexpr (B < E) { f(std::integral_constant<int, B>{}); wg_static_for<B + 1, E>(f); }
}

template <int LA, int LB, int LC, int LD, int MK, int HS = 1, int MINB = 2>
struct WgCfg {
    static constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    static constexpr int NAB = NA * NB, NCD = NC * ND;
    static constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    static constexpr int LAB = LA + LB, LCD = LC + LD;
    static constexpr int GS = (NCD + MK - 1) / MK;
    static constexpr int QW = 32 / GS;
    static constexpr int NQ = QW * WG_WARPS;
    // strides are kept == 2 (mod 4) doubles, i.e. an odd number of 16-byte chunks, so that the 8 lanes of a
    // quarter-warp LDS.128/STS.128 phase that address different columns / blocks / quartets land in different banks
    static constexpr int wg_pad(int n) { return n + ((6 - n % 4) % 4); }
    static constexpr int LABP = wg_pad(LAB + 1);                   // column stride
    static constexpr int TSZ = wg_pad((LC + 1) * (LD + 1) * LABP); // one (root, direction)
    static constexpr int RWP = (2 * NROOTS + 1) & ~1;
    static constexpr int TQ = NROOTS * 3 * TSZ + RWP;
    static constexpr int NAP = NA / HS;                             // components of shell a handled per pass
    static constexpr int NE = NAP * NB;                             // register accumulators per owned S component
    static constexpr int R1 = NE * GS, R2 = (NAP + NB) * NCD;
    static constexpr int SCR = wg_pad(TQ > R1 ? (TQ > R2 ? TQ : R2) : (R1 > R2 ? R1 : R2));   // doubles per quartet
    // GTAB: when the staged Chebyshev root tables (18-64 KB) would keep the MINB-th CTA off the SM (or leave the L1 cache
    // next to nothing of the 256 KB), they are read through L1 from global memory instead.  Measured (profiles r02f):
    // ff|fd 3.17 -> 1.74 ms, ff|ps 2.03 -> 1.44 ms on c18; classes that fit anyway are 2-3 % faster with staged tables.
    static constexpr int TABLEN_STAGED = tpq_table_len(NROOTS);
    static constexpr size_t SMEM_STAGED = sizeof(double) * (size_t)(TABLEN_STAGED + TPQ_NBRA * TPQ_MAXBP + WG_WARPS * QW * SCR);
    static constexpr bool GTAB = NROOTS > 2 && MINB * (SMEM_STAGED + 1024) > 216 * 1024;
    static constexpr int TABLEN = GTAB ? 0 : TABLEN_STAGED;
    static constexpr size_t SMEM = sizeof(double) * (size_t)(TABLEN + TPQ_NBRA * TPQ_MAXBP + WG_WARPS * QW * SCR);
    static_assert(GS >= 1 && GS <= 32, "group does not fit a warp");
    static_assert(NA % HS == 0, "passes must split the components of shell a evenly");
};

// one root/weight value (index v: roots 0..n-1, weights n..2n-1) from the staged Chebyshev table
template <int NROOTS>
__device__ __forceinline__ double wg_rys_value(const double* __restrict__ tab, const double* __restrict__ asym, double T, int v) {
    constexpr int NV = 2 * NROOTS;
    if (T >= (double)rys_tmax(NROOTS)) {
        const double a = asym[v];
        const double rs = rsqrt(T);
        return v < NROOTS ? a * rs * rs : a * rs;
    }
    const int it = (int)(T * 0.5);
    const double u = T - (2.0 * it + 1.0), u2 = u + u;
    const double2* cs = reinterpret_cast<const double2*>(tab + (size_t)(it * NV + v) * RYS_NC);
    double2 c[RYS_NC / 2];
#pragma unroll
    for (int k = 0; k < RYS_NC / 2; k++) c[k] = cs[k];
    double b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int k = RYS_NC - 1; k >= 1; k--) {
        const double ck = (k & 1) ? c[k >> 1].y : c[k >> 1].x;
        const double tt = fma(u2, b1, ck - b2);
        b2 = b1;
        b1 = tt;
    }
    return fma(u, b1, c[0].x - b2);
}

// VRR + S-side (c->d) transfer of one root and one direction; out[(k*(LD+1)+l)*LABP + i] = [i 0 | k l]
template <int LAB, int LC, int LD, int LABP>
__device__ __forceinline__ void wg_vrr_ket(double w0, double c00, double c00p, double b10, double b01, double b00, double cd,
                                           double* __restrict__ out) {
    constexpr int LCD = LC + LD;
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
    for (int l = 0; l <= LD; l++) {
        if (l > 0) {
#pragma unroll
            for (int k = 0; k <= LCD - l; k++)
#pragma unroll
                for (int i = 0; i <= LAB; i++) a[i][k] = fma(cd, a[i][k], a[i][k + 1]);
        }
#pragma unroll
        for (int k = 0; k <= LC; k++) {
            double* o = out + (k * (LD + 1) + l) * LABP;
            if constexpr (((LAB + 1) & 1) == 0) {
#pragma unroll
                for (int i = 0; i <= LAB; i += 2) *reinterpret_cast<double2*>(o + i) = make_double2(a[i][k], a[i + 1][k]);
            } else {
#pragma unroll
                for (int i = 0; i + 1 <= LAB; i += 2) *reinterpret_cast<double2*>(o + i) = make_double2(a[i][k], a[i + 1][k]);
                o[LAB] = a[LAB][k];
            }
        }
    }
}

// column [i = 0..LAB] -> H-side (a->b) transfer: g[j][i] = [i j | . .], i <= LA, j <= LB
template <int LA, int LB, int LABP>
__device__ __forceinline__ void wg_bra_hrr(const double* __restrict__ col, double ab, double (&g)[LB + 1][LA + 1]) {
    constexpr int LAB = LA + LB;
    double h[LAB + 2];
#pragma unroll
    for (int i = 0; i <= LAB; i += 2) {
        const double2 v = *reinterpret_cast<const double2*>(col + i);
        h[i] = v.x; h[i + 1] = v.y;
    }
#pragma unroll
    for (int i = 0; i <= LA; i++) g[0][i] = h[i];
#pragma unroll
    for (int j = 1; j <= LB; j++) {
#pragma unroll
        for (int i = 0; i <= LAB - j; i++) h[i] = fma(ab, h[i], h[i + 1]);
#pragma unroll
        for (int i = 0; i <= LA; i++) g[j][i] = h[i];
    }
}

// HS > 1: the H components are processed in HS passes over the primitive loop (NA/HS components of shell a per pass),
// so that the accumulators of one pass fit the register file; phases A and B are repeated per pass.
template <int LA, int LB, int LC, int LD, int MK, int HS, int MINB = 2>
__global__ void __launch_bounds__(32 * WG_WARPS, MINB) eri_jk_wg(const QuartetTask t) {
    using C = WgCfg<LA, LB, LC, LD, MK, HS, MINB>;
    constexpr int NAP = C::NAP, NE = C::NE;
    constexpr int NA = C::NA, NB = C::NB, NC = C::NC, ND = C::ND, NAB = C::NAB, NCD = C::NCD;
    constexpr int NROOTS = C::NROOTS, LAB = C::LAB, GS = C::GS, QW = C::QW, NQ = C::NQ, LABP = C::LABP, TSZ = C::TSZ;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ double smem[];
    const double scaleJ = __ldg(t.scales), scaleK = __ldg(t.scales + 1);
    const double thr = __ldg(t.scales + 4);   // effective Schwarz threshold of this build (scales_kernel)
    const long long jlo = __ldg(t.scales + 6) != 0.0 ? t.jlo_off : 0;
    unsigned cnt_q = 0, cnt_p = 0;            // counted by lane g == 0 of every quartet group
    const double* tab = smem;
    const double* asym = smem;
    double* sbra = smem + C::TABLEN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = lane / GS, g = lane - qi * GS;
    const bool lane_ok = qi < QW;
    double* myq = sbra + TPQ_NBRA * TPQ_MAXBP + (size_t)(warp * QW + (lane_ok ? qi : 0)) * C::SCR;   // this quartet's scratch
    double* myrw = myq + NROOTS * 3 * TSZ;

    // ---- stage the root tables ------------------------------------------------------------------
    if constexpr (NROOTS <= 2) {
        tpq_stage_tables<NROOTS>(smem, t.rys, threadIdx.x, 32 * WG_WARPS);
    } else {
        constexpr int NTAB = (rys_tmax(NROOTS) / 2) * 2 * NROOTS * RYS_NC;
        if constexpr (C::GTAB) {
            tab = t.rys.table + rys_off(NROOTS);
            asym = t.rys.asym + rys_asym_off(NROOTS);
        } else {
            const double* src = t.rys.table + rys_off(NROOTS);
            for (int e = threadIdx.x; e < NTAB; e += 32 * WG_WARPS) smem[e] = src[e];
            if (threadIdx.x < 2 * NROOTS) smem[NTAB + threadIdx.x] = t.rys.asym[rys_asym_off(NROOTS) + threadIdx.x];
            asym = smem + NTAB;
        }
    }

    // owned S components f = g*MK + m = (ic, id); their exponents and column offsets inside one (root, direction) block
    int fic[MK], fid[MK], colx[MK], coly[MK], colz[MK];
    bool fok[MK];
#pragma unroll
    for (int m = 0; m < MK; m++) {
        const int f = g * MK + m;
        fok[m] = lane_ok && f < NCD;
        const int ff = fok[m] ? f : 0;
        fic[m] = ff / ND; fid[m] = ff - fic[m] * ND;
        colx[m] = (cart_lx(LC, fic[m]) * (LD + 1) + cart_lx(LD, fid[m])) * LABP;
        coly[m] = (cart_ly(LC, fic[m]) * (LD + 1) + cart_ly(LD, fid[m])) * LABP;
        colz[m] = (cart_lz(LC, fic[m]) * (LD + 1) + cart_lz(LD, fid[m])) * LABP;
    }

    const long long nitem_local = (t.nitem - t.rank + t.world - 1) / t.world;
    int4 it_next = make_int4(0, 0, 0, 0);
    if ((long long)blockIdx.x < nitem_local) it_next = __ldg(t.items + (long long)blockIdx.x * t.world + t.rank);
    for (long long li = blockIdx.x; li < nitem_local; li += gridDim.x) {
        const int4 it = it_next;                      // descriptor of the next item is fetched one item ahead
        if (li + gridDim.x < nitem_local) it_next = __ldg(t.items + (li + gridDim.x) * t.world + t.rank);
        const int ib = it.x;
        const int ik = it.y + warp * QW + qi;
        bool active = lane_ok && warp * QW + qi < it.z;
        if (active && thr > 0.0) active = t.bra.Q[ib] * t.ket.Q[ik] > thr;
        if (active && g == 0) cnt_q++;

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
        const int npcd_w = __reduce_max_sync(FULL, npcd);

        int ca = 0, cb = 0, cc0 = 0, cd0 = 0;
        ca = t.bra.cao_a[ib]; cb = t.bra.cao_b[ib];
        if (active) { cc0 = t.ket.cao_a[ik]; cd0 = t.ket.cao_b[ik]; }
        const size_t ld = (size_t)t.ncart;
#ifdef CF_PREFETCH_D
        if (active) {   // D elements of the digestion: one line per (column, a|b row run); the L2 round trip overlaps the integral work
#pragma unroll
            for (int m = 0; m < MK; m++) {
                if (!fok[m]) continue;
                cf_prefetch_l1(t.Dtot + (cd0 + fid[m]) * ld + cc0 + fic[m]);
                for (int x = 0; x < t.nk; x++) {
                    cf_prefetch_l1(t.Dk[x] + (ca) * ld + cd0 + fid[m]);
                    cf_prefetch_l1(t.Dk[x] + (cb) * ld + cd0 + fid[m]);
                    cf_prefetch_l1(t.Dk[x] + (ca) * ld + cc0 + fic[m]);
                    cf_prefetch_l1(t.Dk[x] + (cb) * ld + cc0 + fic[m]);
                }
            }
        }
#endif

        wg_static_for<0, HS>([&](auto hp_tag) {
        constexpr int IA0 = decltype(hp_tag)::value * NAP;     // first component of shell a of this pass
        double acc[MK][NE];
#pragma unroll
        for (int m = 0; m < MK; m++)
#pragma unroll
            for (int e = 0; e < NE; e++) acc[m][e] = 0.0;

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
#ifdef CF_PREFETCH_KET
            double n_qe = 1.0, n_hq = 0.5, n_c = 0.0, n_Qx = 0.0, n_Qy = 0.0, n_Qz = 0.0;   // primitive icd+1, loaded one iteration ahead
            if (active && npcd > 0) {
                n_qe = t.ket.p[pcd0]; n_hq = t.ket.hp[pcd0]; n_c = t.ket.c[pcd0];
                n_Qx = t.ket.Px[pcd0]; n_Qy = t.ket.Py[pcd0]; n_Qz = t.ket.Pz[pcd0];
            }
#endif
            for (int icd = 0; icd < npcd_w; icd++) {
                const bool vk = active && icd < npcd;
                double qe = 1.0, hq = 0.5, ccd = 0.0, Qx = 0.0, Qy = 0.0, Qz = 0.0;
#ifdef CF_PREFETCH_KET
                if (vk) { qe = n_qe; hq = n_hq; ccd = n_c * wgt; Qx = n_Qx; Qy = n_Qy; Qz = n_Qz; }
                if (active && icd + 1 < npcd) {
                    const int scd = pcd0 + (icd + 1) * CF_PSTRIDE;
                    n_qe = t.ket.p[scd]; n_hq = t.ket.hp[scd]; n_c = t.ket.c[scd];
                    n_Qx = t.ket.Px[scd]; n_Qy = t.ket.Py[scd]; n_Qz = t.ket.Pz[scd];
                }
#else
                if (vk) {
                    const int scd = pcd0 + icd * CF_PSTRIDE;
                    qe = t.ket.p[scd]; hq = t.ket.hp[scd]; ccd = t.ket.c[scd] * wgt;
                    Qx = t.ket.Px[scd]; Qy = t.ket.Py[scd]; Qz = t.ket.Pz[scd];
                }
#endif
                const double QCx = Qx - Cx, QCy = Qy - Cy, QCz = Qz - Cz;
                for (int iab = 0; iab < nb; iab++) {
                    const double cc = sbra[5 * TPQ_MAXBP + iab] * ccd;
                    const bool valid = vk && fabs(cc) >= t.prim_cut;
                    if (!__any_sync(FULL, valid)) continue;      // warp-uniform
                    if constexpr (decltype(hp_tag)::value == 0) { if (valid && g == 0) cnt_p++; }   // counted once, not per H pass
                    const double p = sbra[iab], hp = sbra[TPQ_MAXBP + iab];
                    const double PQx = sbra[2 * TPQ_MAXBP + iab] - Qx, PQy = sbra[3 * TPQ_MAXBP + iab] - Qy,
                                 PQz = sbra[4 * TPQ_MAXBP + iab] - Qz;
                    const double pq = p + qe;
                    const double rs = rsqrt(pq), ipq = rs * rs;
                    const double T = (p * qe * ipq) * fma(PQx, PQx, fma(PQy, PQy, PQz * PQz));
                    // ---- A: roots and weights ---------------------------------------------------------------
                    if (valid) {
                        if constexpr (NROOTS <= 2) {
                            if (g == 0) { double rx[NROOTS], rw[NROOTS]; tpq_roots<NROOTS>(tab, T, rx, rw);
#pragma unroll
                                for (int r = 0; r < NROOTS; r++) { myrw[r] = rx[r]; myrw[NROOTS + r] = rw[r]; } }
                        } else {
                            for (int v = g; v < 2 * NROOTS; v += GS) myrw[v] = wg_rys_value<NROOTS>(tab, asym, T, v);
                        }
                    }
                    __syncwarp();
                    // ---- B: VRR + S-side transfer, one (root, direction) per lane-task ---------------------------
                    if (valid) {
                        const double pref = cc * rs;
                        const double qi_ = qe * ipq, pi_ = p * ipq, hi = 0.5 * ipq;
                        for (int tk = g; tk < 3 * NROOTS; tk += GS) {
                            const int r = tk / 3, dim = tk - 3 * r;
                            const double xr = myrw[r];
                            const double rxp = xr * qi_, rxq = xr * pi_, b00 = xr * hi;
                            const double b10 = fma(-rxp, hp, hp), b01 = fma(-rxq, hq, hq);
                            const double PQd = dim == 0 ? PQx : dim == 1 ? PQy : PQz;
                            const double PAd = sbra[(6 + dim) * TPQ_MAXBP + iab];
                            const double QCd = dim == 0 ? QCx : dim == 1 ? QCy : QCz;
                            const double CDd = dim == 0 ? CDx : dim == 1 ? CDy : CDz;
                            const double w0 = dim == 2 ? myrw[NROOTS + r] * pref : 1.0;
                            wg_vrr_ket<LAB, LC, LD, LABP>(w0, fma(-rxp, PQd, PAd), fma(rxq, PQd, QCd), b10, b01, b00, CDd, myq + tk * TSZ);
                        }
                    }
                    __syncwarp();
                    // ---- C: H-side transfer + assembly ------------------------------------------------------------
                    if (valid) {
#pragma unroll 1
                        for (int r = 0; r < NROOTS; r++) {
                            const double* blk = myq + r * 3 * TSZ;
#pragma unroll
                            for (int m = 0; m < MK; m++) {
                                if (!fok[m]) continue;
                                double gx[LB + 1][LA + 1], gy[LB + 1][LA + 1], gz[LB + 1][LA + 1];
                                wg_bra_hrr<LA, LB, LABP>(blk + colx[m], ABx, gx);
                                wg_bra_hrr<LA, LB, LABP>(blk + TSZ + coly[m], ABy, gy);
                                wg_bra_hrr<LA, LB, LABP>(blk + 2 * TSZ + colz[m], ABz, gz);
#pragma unroll
                                for (int e = 0; e < NE; e++) {
                                    const int ia = IA0 + e / NB, jb = e % NB;
                                    acc[m][e] = fma(gx[cart_lx(LB, jb)][cart_lx(LA, ia)] * gy[cart_ly(LB, jb)][cart_ly(LA, ia)],
                                                    gz[cart_lz(LB, jb)][cart_lz(LA, ia)], acc[m][e]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }

        // ---- digestion (warp-synchronous; the quartet's scratch is free now) -----------------------------------
        for (int xj = 0; xj < t.nj; xj++) {   // J(c,d) complete per lane; J(a,b) partial over the lanes of the group
            const double* __restrict__ DJ = t.Dj[xj];
            long long* aJ = t.accJm[xj];
            double dcd[MK], jcd[MK];
#pragma unroll
            for (int m = 0; m < MK; m++) { dcd[m] = (active && fok[m]) ? DJ[(cd0 + fid[m]) * ld + cc0 + fic[m]] : 0.0; jcd[m] = 0.0; }
#pragma unroll
            for (int e = 0; e < NE; e++) {
                const double dab = DJ[(cb + e % NB) * ld + ca + IA0 + e / NB];
                double pab = 0.0;
#pragma unroll
                for (int m = 0; m < MK; m++) { pab = fma(acc[m][e], dcd[m], pab); jcd[m] = fma(acc[m][e], dab, jcd[m]); }
                if (active) myq[e * GS + g] = pab;
            }
#pragma unroll
            for (int m = 0; m < MK; m++)
                if (active && fok[m]) fixed_add_j(aJ + (cd0 + fid[m]) * ld + cc0 + fic[m], jlo, jcd[m], scaleJ);
            __syncwarp();
            // J(a,b) is common to every quartet of the item: sum over ALL active quartets of the warp (fixed order) and
            // issue one add per element and warp; lanes map to consecutive rows i, i.e. consecutive addresses
            {
                const unsigned amask = __ballot_sync(FULL, active);
                const double* wq = myq - (size_t)(lane_ok ? qi : 0) * C::SCR;      // scratch of quartet 0 of this warp
                for (int e2 = lane; e2 < NE; e2 += 32) {
                    const int j = e2 / NAP, i = e2 - j * NAP, e = i * NB + j;
                    double s = 0.0;
#pragma unroll
                    for (int q2 = 0; q2 < QW; q2++)
                        if ((amask >> (q2 * GS)) & 1u) {
#pragma unroll
                            for (int g2 = 0; g2 < GS; g2++) s += wq[(size_t)q2 * C::SCR + e * GS + g2];
                        }
                    if (amask) fixed_add_j(aJ + (cb + j) * ld + ca + IA0 + i, jlo, s, scaleJ);
                }
            }
            __syncwarp();
        }
        for (int x = 0; x < t.nk; x++) {
            const double* __restrict__ D = t.Dk[x];
            long long* accK = t.accK[x];
            // half 1: K(a,c) += sum_bd V D(b,d) ; K(b,c) += sum_ad V D(a,d)   -> slots [.., ic, id], summed over id
            if (active) {
#pragma unroll
                for (int m = 0; m < MK; m++) {
                    if (!fok[m]) continue;
                    double dbd[NB], dad[NAP];
#pragma unroll
                    for (int j = 0; j < NB; j++) dbd[j] = D[(cb + j) * ld + cd0 + fid[m]];
#pragma unroll
                    for (int i = 0; i < NAP; i++) dad[i] = D[(ca + IA0 + i) * ld + cd0 + fid[m]];
                    double kbc[NB];
#pragma unroll
                    for (int j = 0; j < NB; j++) kbc[j] = 0.0;
#pragma unroll
                    for (int i = 0; i < NAP; i++) {
                        double kac = 0.0;
#pragma unroll
                        for (int j = 0; j < NB; j++) { kac = fma(acc[m][i * NB + j], dbd[j], kac); kbc[j] = fma(acc[m][i * NB + j], dad[i], kbc[j]); }
                        myq[(i * NC + fic[m]) * ND + fid[m]] = kac;
                    }
#pragma unroll
                    for (int j = 0; j < NB; j++) myq[NAP * NCD + (j * NC + fic[m]) * ND + fid[m]] = kbc[j];
                }
            }
            __syncwarp();
            if (active)
                for (int tg = g; tg < (NAP + NB) * NC; tg += GS) {     // c components fastest: consecutive lanes -> consecutive addresses
                    const int r = tg / NC, k = tg - r * NC;
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < ND; l++) s += myq[(r * NC + k) * ND + l];
                    const int row = r < NAP ? ca + IA0 + r : cb + (r - NAP);
                    fixed_add(accK + (row) * ld + cc0 + k, s, scaleK);
                }
            __syncwarp();
            // half 2: K(a,d) += sum_bc V D(b,c) ; K(b,d) += sum_ac V D(a,c)   -> slots [.., id, ic], summed over ic
            if (active) {
#pragma unroll
                for (int m = 0; m < MK; m++) {
                    if (!fok[m]) continue;
                    double dbc[NB], dac[NAP];
#pragma unroll
                    for (int j = 0; j < NB; j++) dbc[j] = D[(cb + j) * ld + cc0 + fic[m]];
#pragma unroll
                    for (int i = 0; i < NAP; i++) dac[i] = D[(ca + IA0 + i) * ld + cc0 + fic[m]];
                    double kbd[NB];
#pragma unroll
                    for (int j = 0; j < NB; j++) kbd[j] = 0.0;
#pragma unroll
                    for (int i = 0; i < NAP; i++) {
                        double kad = 0.0;
#pragma unroll
                        for (int j = 0; j < NB; j++) { kad = fma(acc[m][i * NB + j], dbc[j], kad); kbd[j] = fma(acc[m][i * NB + j], dac[i], kbd[j]); }
                        myq[(i * ND + fid[m]) * NC + fic[m]] = kad;
                    }
#pragma unroll
                    for (int j = 0; j < NB; j++) myq[NAP * NCD + (j * ND + fid[m]) * NC + fic[m]] = kbd[j];
                }
            }
            __syncwarp();
            if (active)
                for (int tg = g; tg < (NAP + NB) * ND; tg += GS) {
                    const int r = tg / ND, l = tg - r * ND;
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NC; k++) s += myq[(r * ND + l) * NC + k];
                    const int row = r < NAP ? ca + IA0 + r : cb + (r - NAP);
                    fixed_add(accK + (row) * ld + cd0 + l, s, scaleK);
                }
            __syncwarp();
        }
        });   // passes over the H components
    }
    cf_cnt_flush(t.cnt, cnt_q, cnt_p);
}

// class-pair -> warp-group configuration: MK | (swap << 8) | (HS << 12) | (MINB << 16); 0 = not covered (HS field 0 means
// 1, MINB field 0 means 2 resident CTAs per SM).  Classes are (la*(la+1)/2 + lb); `swap` means the launcher hands the LOWER
// class over as the CTA-uniform (H) pair.  `variant` selects among alternatives compiled with -DCF_WG_NVAR=n for A/B
// measurements (env CF_WG_VARIANT); the default build has only variant 0 = the measured best.
// Measured (profiles/r01n_*): more resident CTAs with fewer quartets per warp do NOT help -- throughput follows the
// number of quartets in flight per SM, so the large-H / large-MK shapes stay; three classes gain from MK = 1.
#define WGC(mk, sw, hs, minb) ((mk) | ((sw) << 8) | ((hs) << 12) | ((minb) << 16))
#ifndef CF_WG_NVAR
#define CF_WG_NVAR 1
#endif
__host__ __device__ constexpr int wg_cfg_base(int key) {
    switch (key) {
        case 42: return 3;          // dp|pp
        case 44: return 3;          // dp|dp
        case 52: return 1;          // dd|pp
        case 53: return 1;          // dd|ds
        case 54: return 2;          // dd|dp
        case 55: return 2;          // dd|dd
        case 64: return 3 | 256;    // fs|dp  (H = dp)
        case 65: return WGC(1, 1, 1, 3);   // fs|dd  (H = dd), 10 lanes per quartet, 3 CTAs per SM
        case 72: return WGC(1, 0, 1, 3);   // fp|pp
        case 73: return WGC(1, 0, 1, 3);   // fp|ds
        case 74: return 2;          // fp|dp
        case 75: return 2 | 256;    // fp|dd  (H = dd)
        case 76: return 2;          // fp|fs
        case 77: return 2;          // fp|fp
        case 81: return 1;          // fd|ps
        case 82: return 1;          // fd|pp
        case 83: return 1;          // fd|ds
        case 84: return 4 | 256;    // fd|dp  (H = dp)
        case 85: return 2 | 256;    // fd|dd  (H = dd)
        case 86: return 1;          // fd|fs
        case 87: return 1;          // fd|fp
        case 88: return 2 | (2 << 12);          // fd|fd  two passes of 5 a-components
        case 91: return 1 | (2 << 12);          // ff|ps
        case 92: return 1 | (2 << 12);          // ff|pp
        case 93: return 1 | (2 << 12);          // ff|ds
        case 94: return 4 | 256;                // ff|dp  (H = dp)
        case 95: return 4 | 256 | (2 << 12);    // ff|dd  (H = dd, two passes)
        case 96: return 1 | (2 << 12);          // ff|fs
        case 97: return 4 | 256 | (2 << 12);    // ff|fp  (H = fp, two passes)
        case 98: return 2 | (5 << 12);          // ff|fd  five passes of 2 a-components
        case 99: return 4 | (10 << 12);         // ff|ff  ten passes of 1 a-component (was the generic CTA-per-quartet kernel)
        case 71: return 1;                      // fp|ps  (measured against the sliced thread-per-quartet kernel: 4.63 -> 3.86 ms on c18)
        case 51: return 1;                      // dd|ps  (2.49 -> 2.31 ms)
        case 66: return 1;                      // fs|fs  (1.89 -> 1.48 ms); pp|pp and fs|pp stay with the sliced kernels (slower as warp-group kernels)
        case 43: return 2;                      // dp|ds, two S components per lane (4.67 -> 4.63 ms on c18, 100 -> 86 ms on (H2O)64)
        default: return 0;
    }
}
__host__ __device__ constexpr int wg_cfg_alt1(int key) {   // scratch pad for the next A/B round
    switch (key) {
        default: return wg_cfg_base(key);
    }
}
__host__ __device__ constexpr int wg_cfg_alt2(int key) {
    switch (key) {
        default: return wg_cfg_base(key);
    }
}
__host__ __device__ constexpr int wg_cfg(int bra_cls, int ket_cls, int variant = 0) {
#ifdef CF_NO_WG
    return 0;
#else
    const int key = bra_cls * 10 + ket_cls;
    return variant == 1 ? wg_cfg_alt1(key) : variant == 2 ? wg_cfg_alt2(key) : wg_cfg_base(key);
#endif
}
