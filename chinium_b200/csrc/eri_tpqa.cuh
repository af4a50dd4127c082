// eri_tpqa.cuh -- thread-per-quartet Rys kernel with BRA-LOOP digestion (replaces eri_jk_tpq on the build path).
//
// Why: the thread-per-quartet classes (<= 64 Cartesian integrals per quartet) were bound by the fixed-point
// atomics of the digestion, not by FP64: NCD + (NA+NB)(NC+ND) scattered RED.64 per quartet (ps|ss: 9 atomics for
// ~600 flops; measured 218 G scattered RED/s on this part, profiles/r01i_red_micro.txt).  Here a warp keeps its 32
// ket pairs (c,d) FIXED and loops over a chunk of bra pairs (a,b) that all share shell a:
//   J(c,d)            accumulates in registers over the whole chunk    -> NCD atomics per lane and ITEM
//   K(a,c), K(a,d)    accumulate in registers (a is common)            -> NA(NC+ND) atomics per lane and ITEM
//   K(b,c), K(b,d)    b changes every iteration                        -> NB(NC+ND) atomics per lane and quartet
//   J(a,b)            warp-wide sum over the 32 kets                   -> NAB atomics per WARP and bra pair
// la >= lb, so the larger K blocks are the ones that stay in registers: ps|ss 9 -> 2, ds|ps 31 -> 4, fs|ps 47 -> 4
// atomics per quartet.  D(c,d) and (small classes) D(a,c), D(a,d) are loaded once per item as well.
//   work item = (chunk of <= TPQA_CHUNK bra pairs of one a-group, sorted by Schwarz bound desc) x (one aligned block
//               of 32 consecutive ket pairs); items are warp-private.  `border` lists the bra pairs (positions in the
//               class arrays) grouped by shell a.  Triangular tasks: quartet canonical iff ket position <= bra
//               position; bra pairs entirely below the ket block are skipped warp-uniformly.
// The set of fixed-point adds is the same for every rank count (items are dealt round-robin), so results stay
// bit-identical for any number of GPUs.
// Replaces libint2's engine.compute + the reference's stored-integral digestion
// (src/Integral/Int4C2E.cpp:233-302 and :601-671).
#pragma once
#include "cf_common.cuh"
#include "eri_generic.cuh"
#include "eri_tpq.cuh"

// developer experiment (session r2u): L1 prefetch hints for the three places where the (H2O)64 profile shows long-scoreboard
// stalls -- bit 0: the chunk's bra records at item start, bit 1: the item's ket primitive rows at item start, bit 2: the
// bra-dependent density lines of the digestion at the top of a bra-pair iteration.  0 = off (the default build).
#ifndef CF_TPQA_PF
#define CF_TPQA_PF 0
#endif
#ifndef CF_HAVE_PREFETCH_L1
#define CF_HAVE_PREFETCH_L1
__device__ __forceinline__ void cf_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

#ifndef TPQA_ACCK_MAX
#define TPQA_ACCK_MAX 44      // K(a,c)/K(a,d) register accumulators: at most this many doubles per thread
#endif
#ifndef TPQA_KEEPD_MAX
#define TPQA_KEEPD_MAX 16     // D(a,c)/D(a,d) kept in registers for the whole item up to this many doubles
#endif


__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// resident CTAs per SM the register allocation is held to: the light classes are latency-bound and want warps
// Measured (profiles/r2e_*): holding the smallest classes to 80-96 registers for 5-6 resident CTAs makes them SLOWER on
// (H2O)64 (ps|ps +18 %, pp|ss +8 %, ss|ss +3 %: the extra warps do not pay for the spills), so round 1's rule stays.
__host__ __device__ constexpr int tpqa_minb(int nout, int nkmax) {
#ifdef TPQA_MINB
    return TPQA_MINB;
#else
    return nout * nkmax <= 9 ? 4 : nout * nkmax <= 18 ? 3 : 2;
#endif
}

// NKMAX / NJMAX: exchange / Coulomb densities the register accumulators are sized for (NJMAX = 3 only for the
// multi-density build, Int4C2E.cpp:685-745)
template <int LA, int LB, int LC, int LD, int NKMAX, int NJMAX = 1>
__global__ void __launch_bounds__(TPQ_THREADS, tpqa_minb(cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD), NKMAX * NJMAX))
eri_jk_tpqa(const QuartetTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND, NOUT = NAB * NCD;
    constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    constexpr int GSZ = (LA + 1) * (LB + 1) * (LC + 1) * (LD + 1);
    constexpr int TABLEN = tpq_table_len(NROOTS);
    constexpr int MAXBP = TPQ_WBP;
    constexpr int NKA = NA * (NC + ND);                       // K(a,c) | K(a,d) elements per density
    constexpr bool ACCK = NKMAX * NKA <= TPQA_ACCK_MAX;       // keep them in registers over the bra loop
    constexpr bool KEEPD = ACCK && NKMAX * NKA <= TPQA_KEEPD_MAX;
    constexpr int NKACC = ACCK ? NKMAX * NKA : 1;
    constexpr int NKEEP = KEEPD ? NKMAX * NKA : 1;
    extern __shared__ double smem[];
    const double scaleJ = __ldg(t.scales), scaleK = __ldg(t.scales + 1);
    const double thr = __ldg(t.scales + 4);   // effective Schwarz threshold of this build (scales_kernel)
    const long long jlo = __ldg(t.scales + 6) != 0.0 ? t.jlo_off : 0;
    unsigned cnt_q = 0, cnt_p = 0;            // this lane's evaluated shell quartets / executed primitive quartets
    double* tab = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sbra = smem + TABLEN + warp * (TPQ_NBRA * MAXBP);   // this warp's [TPQ_NBRA][MAXBP]

    tpq_stage_tables<NROOTS>(tab, t.rys, threadIdx.x, TPQ_THREADS);
    __syncthreads();

    const size_t ld = (size_t)t.ncart;
    const bool same = t.same_class != 0;
    const long long nitem_local = (t.nitem - t.rank + t.world - 1) / t.world;
    const long long gw = (long long)blockIdx.x * (TPQ_THREADS / 32) + warp, nw = (long long)gridDim.x * (TPQ_THREADS / 32);
    int4 it_next = make_int4(0, 0, 0, 0);
    if (gw < nitem_local) it_next = __ldg(t.items + gw * t.world + t.rank);
    for (long long li = gw; li < nitem_local; li += nw) {
        const int4 it = it_next;                      // descriptor of the next item is fetched one item ahead
        if (li + nw < nitem_local) it_next = __ldg(t.items + (li + nw) * t.world + t.rank);
        const int i0 = it.x, k0 = it.y, nbra = it.w;
        const int ik = k0 + lane;
        const bool lane_ok = lane < it.z;

        // ---- ket pair of this lane: fixed for the whole item -------------------------------------------------
        int sc = 0, sd = 0, pcd0 = 0, npcd = 0, cc0 = 0, cd0 = 0;
        double Cx = 0, Cy = 0, Cz = 0, CDx = 0, CDy = 0, CDz = 0, Qk = 0;
        if (lane_ok) {
            sc = t.ket.sa[ik]; sd = t.ket.sb[ik];
            Cx = t.ket.A[3 * ik]; Cy = t.ket.A[3 * ik + 1]; Cz = t.ket.A[3 * ik + 2];
            CDx = t.ket.AB[3 * ik]; CDy = t.ket.AB[3 * ik + 1]; CDz = t.ket.AB[3 * ik + 2];
            pcd0 = t.ket.pbase[ik]; npcd = t.ket.nprim[ik];
            cc0 = t.ket.cao_a[ik]; cd0 = t.ket.cao_b[ik];
            Qk = t.ket.Q[ik];
        }
        const double qrun = thr > 0.0 ? warp_max(Qk) : 0.0;
        const double wcd = (sc == sd) ? 1.0 : 2.0;
#if CF_TPQA_PF & 1
        {   // the chunk's records are read one after the other by the bra loop: their lines, up front
            const char* pi = reinterpret_cast<const char*>(t.brec_i + i0);
            const char* pd = reinterpret_cast<const char*>(t.brec_d + 4 * (size_t)i0);
            if (lane * 128 < nbra * 16) cf_prefetch_l1(pi + lane * 128);
            if (lane * 128 < nbra * 32) cf_prefetch_l1(pd + lane * 128);
        }
#endif
#if CF_TPQA_PF & 2
        if (lane_ok && (lane & 15) == 0) {   // ket primitive rows of the block (32 pairs interleaved: 256 bytes per array and primitive)
            const int np = min(npcd, 6);
            for (int icd = 0; icd < np; icd++) {
                const int scd = pcd0 + icd * CF_PSTRIDE;
                cf_prefetch_l1(t.ket.p + scd); cf_prefetch_l1(t.ket.hp + scd); cf_prefetch_l1(t.ket.c + scd);
                cf_prefetch_l1(t.ket.Px + scd); cf_prefetch_l1(t.ket.Py + scd); cf_prefetch_l1(t.ket.Pz + scd);
            }
        }
#endif

        // ---- shell a: common to the chunk --------------------------------------------------------------------
        // Bra pairs come from the bra-role copy of the class (records + CONTIGUOUS primitives in `border` order), so the
        // record of pair ii+2 and the primitives of pair ii+1 are in flight while pair ii is computed: no dependent
        // global-load latency is exposed inside the loop (low-contraction systems are latency-, not FP64-bound).
        int4 ri0 = __ldg(t.brec_i + i0);
        const int ibf = ri0.x;
        const int sa = t.bra.sa[ibf], ca = t.bra.cao_a[ibf];
        const double Ax = t.bra.A[3 * ibf], Ay = t.bra.A[3 * ibf + 1], Az = t.bra.A[3 * ibf + 2];
        double2 rq0 = __ldg(reinterpret_cast<const double2*>(t.brec_d + 4 * (size_t)i0));
        double2 rz0 = __ldg(reinterpret_cast<const double2*>(t.brec_d + 4 * (size_t)i0) + 1);
        int4 ri1 = ri0; double2 rq1 = rq0, rz1 = rz0;
        if (nbra > 1) {
            ri1 = __ldg(t.brec_i + i0 + 1);
            rq1 = __ldg(reinterpret_cast<const double2*>(t.brec_d + 4 * (size_t)(i0 + 1)));
            rz1 = __ldg(reinterpret_cast<const double2*>(t.brec_d + 4 * (size_t)(i0 + 1)) + 1);
        }
        const size_t bps = (size_t)t.bprim_stride;
        auto stage = [&](int poff, int nb) {      // primitives poff .. poff+nb-1 of a bra pair -> this warp's shared block
            if (lane < nb) {
                const double* bp = t.bprim + poff + lane;
                const double px = bp[2 * bps], py = bp[3 * bps], pz = bp[4 * bps];
                sbra[0 * MAXBP + lane] = bp[0];
                sbra[1 * MAXBP + lane] = bp[bps];
                sbra[2 * MAXBP + lane] = px;
                sbra[3 * MAXBP + lane] = py;
                sbra[4 * MAXBP + lane] = pz;
                sbra[5 * MAXBP + lane] = bp[5 * bps];
                sbra[6 * MAXBP + lane] = px - Ax;
                sbra[7 * MAXBP + lane] = py - Ay;
                sbra[8 * MAXBP + lane] = pz - Az;
            }
        };
        __syncwarp();
        stage(ri0.w, min(MAXBP, ri0.y >> 16));
        __syncwarp();

        double dcd[NJMAX * NCD], jcd[NJMAX * NCD];
#pragma unroll
        for (int xj = 0; xj < NJMAX; xj++)
#pragma unroll
            for (int kl = 0; kl < NCD; kl++) {
                dcd[xj * NCD + kl] = (lane_ok && xj < t.nj) ? t.Dj[xj][(cd0 + kl % ND) * ld + cc0 + kl / ND] : 0.0;
                jcd[xj * NCD + kl] = 0.0;
            }
        double kacc[NKACC];      // [x][ K(a,c): i*NC+k | K(a,d): NA*NC + i*ND+l ]
#pragma unroll
        for (int e = 0; e < NKACC; e++) kacc[e] = 0.0;
        double dkeep[NKEEP];     // [x][ D(a,c): i*NC+k | D(a,d): NA*NC + i*ND+l ]
        if constexpr (KEEPD) {
#pragma unroll
            for (int x = 0; x < NKMAX; x++)
                if (x < t.nk) {
                    const double* __restrict__ D = t.Dk[x];
#pragma unroll
                    for (int e = 0; e < NA * NC; e++) dkeep[x * NKA + e] = lane_ok ? D[(ca + e / NC) * ld + cc0 + e % NC] : 0.0;
#pragma unroll
                    for (int e = 0; e < NA * ND; e++) dkeep[x * NKA + NA * NC + e] = lane_ok ? D[(ca + e / ND) * ld + cd0 + e % ND] : 0.0;
                }
        }

        int nact = 0;
        for (int ii = 0; ii < nbra; ii++) {
            // ---- pipeline: record of pair ii+2, primitives of pair ii+1 ------------------------------------------
            int4 ri2 = ri1; double2 rq2 = rq1, rz2 = rz1;
            if (ii + 2 < nbra) {
                ri2 = __ldg(t.brec_i + i0 + ii + 2);
                rq2 = __ldg(reinterpret_cast<const double2*>(t.brec_d + 4 * (size_t)(i0 + ii + 2)));
                rz2 = __ldg(reinterpret_cast<const double2*>(t.brec_d + 4 * (size_t)(i0 + ii + 2)) + 1);
            }
            const int nb1 = (ii + 1 < nbra) ? min(MAXBP, ri1.y >> 16) : 0;
            double n_p = 1.0, n_hp = 0.5, n_px = 0.0, n_py = 0.0, n_pz = 0.0, n_c = 0.0;
            if (lane < nb1) {
                const double* bp = t.bprim + ri1.w + lane;
                n_p = bp[0]; n_hp = bp[bps]; n_px = bp[2 * bps]; n_py = bp[3 * bps]; n_pz = bp[4 * bps]; n_c = bp[5 * bps];
            }
            auto rotate = [&]() {                 // pair ii+1 becomes current: its primitives go to shared memory
                __syncwarp();
                if (lane < nb1) {
                    sbra[0 * MAXBP + lane] = n_p; sbra[1 * MAXBP + lane] = n_hp;
                    sbra[2 * MAXBP + lane] = n_px; sbra[3 * MAXBP + lane] = n_py; sbra[4 * MAXBP + lane] = n_pz;
                    sbra[5 * MAXBP + lane] = n_c;
                    sbra[6 * MAXBP + lane] = n_px - Ax; sbra[7 * MAXBP + lane] = n_py - Ay; sbra[8 * MAXBP + lane] = n_pz - Az;
                }
                __syncwarp();
                ri0 = ri1; rq0 = rq1; rz0 = rz1;
                ri1 = ri2; rq1 = rq2; rz1 = rz2;
            };

            const int ib = ri0.x;
            const double Qb = rq0.x;
            if (thr > 0.0 && !(Qb * qrun > thr)) break;         // the chunk is sorted by Q descending
            if (same && ib < k0) { rotate(); continue; }        // every ket of the block lies above this bra pair
            bool act = lane_ok && (!same || ik <= ib);
            if (act && thr > 0.0) act = Qb * Qk > thr;
            const unsigned amask = __ballot_sync(0xffffffffu, act);
            if (!amask) { rotate(); continue; }
            nact += __popc(amask);
            cnt_q += act ? 1u : 0u;

            const int sb = ri0.y & 0xffff, cb = ri0.z;
#if CF_TPQA_PF & 4
            if (act) {   // density lines of this bra pair's digestion: rows of shell b
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    for (int x = 0; x < t.nk; x++) {
                        cf_prefetch_l1(t.Dk[x] + (size_t)(cb + j) * ld + cd0);
                        cf_prefetch_l1(t.Dk[x] + (size_t)(cb + j) * ld + cc0);
                    }
                    if (lane == 0) cf_prefetch_l1(t.Dj[0] + (size_t)(cb + j) * ld + ca);
                }
            }
#endif
            const double ABx = rq0.y, ABy = rz0.x, ABz = rz0.y;
            const int pab0 = ri0.w, npab = ri0.y >> 16;
            const double wgt = (sa == sb ? 1.0 : 2.0) * wcd * ((same && ib == ik) ? 1.0 : 2.0);

            double gout[NOUT];
#pragma unroll
            for (int n = 0; n < NOUT; n++) gout[n] = 0.0;

            for (int b0 = 0; b0 < npab; b0 += MAXBP) {
                const int nb = min(MAXBP, npab - b0);
                if (b0 > 0) {      // pairs with more than MAXBP primitives: further passes are staged on the spot
                    __syncwarp();
                    stage(pab0 + b0, nb);
                    __syncwarp();
                }
                if (!act) continue;
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
            rotate();       // shared block now holds pair ii+1; the digestion below works from registers only

            // ---- digestion of this bra pair (inactive lanes hold gout == 0 and only join the warp-wide J(a,b) sums)
#pragma unroll
            for (int xj = 0; xj < NJMAX; xj++) {
                if (xj >= t.nj) break;
                const double* __restrict__ DJ = t.Dj[xj];
                long long* aJ = t.accJm[xj];
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
                        s = fma(gout[ij * NCD + kl], dcd[xj * NCD + kl], s);
                        jcd[xj * NCD + kl] = fma(gout[ij * NCD + kl], dab, jcd[xj * NCD + kl]);
                    }
#ifdef CF_WARP_MULTI_SUM
                    sab[ij] = s;
#else
                    s = warp_sum_fixed(s);
                    if (lane == 0) fixed_add_j(aJ + off, jlo, s, scaleJ);
#endif
                }
#ifdef CF_WARP_MULTI_SUM
                {   // all NAB warp sums by recursive halving (eri_tpq.cuh); the lanes holding totals add them
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
            }
            if (!act) continue;
#pragma unroll
            for (int x = 0; x < NKMAX; x++) {
                if (x >= t.nk) break;
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
                            if constexpr (ACCK) kacc[x * NKA + i * NC + k] += s;
                            else fixed_add(acc + (ca + i) * ld + cc0 + k, s, scaleK);
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
                            if constexpr (ACCK) kacc[x * NKA + NA * NC + i * ND + l] += s;
                            else fixed_add(acc + (ca + i) * ld + cd0 + l, s, scaleK);
                        }
                }
                {   // K(b,c) += sum_ad V D(a,d)
                    double d[NA * ND];
#pragma unroll
                    for (int e = 0; e < NA * ND; e++) {
                        if constexpr (KEEPD) d[e] = dkeep[x * NKA + NA * NC + e];
                        else d[e] = D[(ca + e / ND) * ld + cd0 + e % ND];
                    }
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
                    for (int e = 0; e < NA * NC; e++) {
                        if constexpr (KEEPD) d[e] = dkeep[x * NKA + e];
                        else d[e] = D[(ca + e / NC) * ld + cc0 + e % NC];
                    }
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

        // ---- flush what accumulated over the chunk ---------------------------------------------------------------
        if (nact == 0) continue;   // warp-uniform: nothing was evaluated
        if (lane_ok) {
#pragma unroll
            for (int xj = 0; xj < NJMAX; xj++) {
                if (xj >= t.nj) break;
#pragma unroll
                for (int kl = 0; kl < NCD; kl++) fixed_add_j(t.accJm[xj] + (cd0 + kl % ND) * ld + cc0 + kl / ND, jlo, jcd[xj * NCD + kl], scaleJ);
            }
            if constexpr (ACCK) {
#pragma unroll
                for (int x = 0; x < NKMAX; x++) {
                    if (x >= t.nk) break;
                    long long* acc = t.accK[x];
#pragma unroll
                    for (int e = 0; e < NA * NC; e++) fixed_add(acc + (ca + e / NC) * ld + cc0 + e % NC, kacc[x * NKA + e], scaleK);
#pragma unroll
                    for (int e = 0; e < NA * ND; e++) fixed_add(acc + (ca + e / ND) * ld + cd0 + e % ND, kacc[x * NKA + NA * NC + e], scaleK);
                }
            }
        }
    }
    cf_cnt_flush(t.cnt, cnt_q, cnt_p);
}
