// eri_generic.cuh -- generic Rys-quadrature shell-quartet kernel with fused J/K digestion.
//
// One CTA of G threads works on one contracted shell quartet (ab|cd) at a time:
//   A  lanes < 2*NROOTS evaluate the Rys roots/weights of the primitive quartet (Chebyshev tables)
//   B  lanes < 3*NROOTS run the 2-D recurrences (VRR + both HRRs) for one (root, x|y|z) each and
//      leave I_x, I_y, I_z in shared memory
//   C  every lane assembles its share of the NOUT Cartesian integrals, accumulating over primitive
//      quartets in registers
//   D  the contracted block goes to shared memory once and is digested owner-computes: each lane
//      owns target elements of the six J/K blocks, forms its dot products in a fixed order and
//      adds them to the global accumulators as 64-bit FIXED-POINT integers (order-independent,
//      hence bit-stable for any schedule, grid size or number of GPUs).
// Replaces libint2's engine.compute + the stored-integral digestion of the reference
// (src/Integral/Int4C2E.cpp:233-302 and :601-671).
#pragma once
#include "cf_common.cuh"

#define RYS_NC 14
// table geometry: must match tools/gen_rys_tables.py (checked against the generated header in engine.cu)
__host__ __device__ constexpr int rys_tmax(int n) { return (32 + 7 * n) % 2 == 0 ? 32 + 7 * n : 33 + 7 * n; }
__host__ __device__ constexpr int rys_off(int n) {
    int off = 0;
    for (int m = 1; m < n; m++) off += (rys_tmax(m) / 2) * 2 * m * RYS_NC;
    return off;
}
__host__ __device__ constexpr int rys_asym_off(int n) { return n * (n - 1); }

// roots x_r (= t^2) -> out[0..n-1], weights -> out[n..2n-1]; value index v handled by the caller's lane
template <int NROOTS>
__device__ __forceinline__ double rys_value(const RysTablesDev& c_rys, double T, int v) {
    if (T >= (double)rys_tmax(NROOTS)) {
        const double a = c_rys.asym[rys_asym_off(NROOTS) + v];
        return v < NROOTS ? a / T : a * rsqrt(T);
    }
    const int it = (int)(T * 0.5);
    const double u = T - (2.0 * it + 1.0);
    const double* cs = c_rys.table + rys_off(NROOTS) + (size_t)(it * 2 * NROOTS + v) * RYS_NC;
    double b1 = 0.0, b2 = 0.0;
    const double u2 = u + u;
#pragma unroll
    for (int k = RYS_NC - 1; k >= 1; k--) {
        const double t = fma(u2, b1, __ldg(cs + k) - b2);
        b2 = b1;
        b1 = t;
    }
    return fma(u, b1, __ldg(cs) - b2);
}

// n-th Cartesian component of angular momentum l in the order lx descending, ly descending
__device__ __forceinline__ void cart_comp(int l, int n, int& lx, int& ly, int& lz) {
    int x = l;
    while (n > l - x) { n -= (l - x + 1); x--; }
    lx = x; ly = (l - x) - n; lz = n;
}

// 2-D Rys integrals of one root and one Cartesian direction -> g[idx(i,j,k,l)]
template <int LA, int LB, int LC, int LD>
__device__ __forceinline__ void rys_2d(double w0, double c00, double c00p, double b10, double b01, double b00,
                                       double ab, double cd, double* __restrict__ g) {
    constexpr int LAB = LA + LB, LCD = LC + LD;
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
    // bra transfer i -> j, then ket transfer k -> l
    double b[LB + 1][LA + 1][LCD + 1];
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
#pragma unroll
    for (int i = 0; i <= LA; i++) {
#pragma unroll
        for (int j = 0; j <= LB; j++) {
            double c[LCD + 1];
#pragma unroll
            for (int k = 0; k <= LCD; k++) c[k] = b[j][i][k];
#pragma unroll
            for (int l = 0; l <= LD; l++) {
                if (l > 0) {
#pragma unroll
                    for (int k = 0; k <= LCD - l; k++) c[k] = fma(cd, c[k], c[k + 1]);
                }
#pragma unroll
                for (int k = 0; k <= LC; k++) g[((i * (LB + 1) + j) * (LC + 1) + k) * (LD + 1) + l] = c[k];
            }
        }
    }
}

__device__ __forceinline__ void fixed_add(long long* addr, double v, double scale) {
    const long long q = __double2ll_rn(v * scale);
    atomicAdd(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(q));
}
// J adds: high limb as above; with jlo != 0 the rounding residual (exact in FP64: |x| < 2^52 -> x - rn(x) is
// representable, larger x are integers) is added, scaled by 2^31, to the low limb `jlo` words further on.
// Partial sums may wrap around 2^64 in either limb: integer addition is exact modulo 2^64 and only the FINAL value has
// to fit, which the scale guarantees (scales_kernel in engine.cu).
__device__ __forceinline__ void fixed_add_j(long long* addr, long long jlo, double v, double scale) {
    const double x = v * scale;
    const long long q = __double2ll_rn(x);
    atomicAdd(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(q));
    if (jlo) {
        const long long ql = __double2ll_rn((x - (double)q) * 0x1p31);
        atomicAdd(reinterpret_cast<unsigned long long*>(addr + jlo), static_cast<unsigned long long>(ql));
    }
}

template <int LA, int LB, int LC, int LD, int G, bool STORE>
__global__ void __launch_bounds__(G) eri_jk_generic(const QuartetTask t) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NAB = NA * NB, NCD = NC * ND, NOUT = NAB * NCD;
    constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    constexpr int GSZ = (LA + 1) * (LB + 1) * (LC + 1) * (LD + 1);
    constexpr int NPL = (NOUT + G - 1) / G;   // outputs per lane
    // shared: rw[2*NROOTS] | g[NROOTS*3*GSZ] | V[NOUT] | D blocks
    extern __shared__ double smem[];
    const double scaleJ = STORE ? 1.0 : __ldg(t.scales), scaleK = STORE ? 1.0 : __ldg(t.scales + 1);
    const double thr = STORE ? t.thr : __ldg(t.scales + 4);   // effective Schwarz threshold of this build (scales_kernel)
    const long long jlo = (!STORE && __ldg(t.scales + 6) != 0.0) ? t.jlo_off : 0;
    unsigned cnt_q = 0, cnt_p = 0;          // evaluated shell quartets / executed primitive quartets (lane 0 counts)
    double* rw = smem;
    double* g = rw + 2 * NROOTS;
    double* V = g + NROOTS * 3 * GSZ;
    double* Dab = V + NOUT;                 // [NAB]  Dtot(a,b)
    double* Dcd = Dab + NAB;                // [NCD]
    double* Dx = Dcd + NCD;                 // per K density: ac[NA*NC] ad[NA*ND] bc[NB*NC] bd[NB*ND]
    constexpr int NDX = NA * NC + NA * ND + NB * NC + NB * ND;

    const int lane = threadIdx.x;

    // this lane's outputs n = lane + m*G and their packed (ix,iy,iz) table indices
    int idx3[NPL];
#pragma unroll
    for (int m = 0; m < NPL; m++) {
        const int n = lane + m * G;
        int packed = 0;
        if (n < NOUT) {
            const int id = n % ND, ic = (n / ND) % NC, ib = (n / (ND * NC)) % NB, ia = n / (ND * NC * NB);
            int ax, ay, az, bx, by, bz, cx, cy, cz, dx, dy, dz;
            cart_comp(LA, ia, ax, ay, az); cart_comp(LB, ib, bx, by, bz);
            cart_comp(LC, ic, cx, cy, cz); cart_comp(LD, id, dx, dy, dz);
            const int ix = ((ax * (LB + 1) + bx) * (LC + 1) + cx) * (LD + 1) + dx;
            const int iy = ((ay * (LB + 1) + by) * (LC + 1) + cy) * (LD + 1) + dy;
            const int iz = ((az * (LB + 1) + bz) * (LC + 1) + cz) * (LD + 1) + dz;
            packed = ix | (iy << 10) | (iz << 20);
        }
        idx3[m] = packed;
    }

    const long long nchunk_total = (t.nquartet + t.chunk - 1) / t.chunk;
    const long long nchunk_local = (nchunk_total - t.rank + t.world - 1) / t.world;

    for (long long lc = blockIdx.x; lc < nchunk_local; lc += gridDim.x) {
        const long long chunk = lc * t.world + t.rank;
        long long q = chunk * t.chunk;
        const long long q_end = min(q + (long long)t.chunk, t.nquartet);
        // bra pair of the first quartet: largest ib with qoff[ib] <= q
        int lo = 0, hi = t.diag ? 1 : t.bra.npair;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (t.qoff[mid] <= q) lo = mid; else hi = mid;
        }
        int ib = lo;
        for (; q < q_end; q++) {
            int ik;
            if (t.diag) { ib = (int)q; ik = (int)q; }
            else { while (t.qoff[ib + 1] <= q) ib++; ik = (int)(q - t.qoff[ib]); }

            const int sa = t.bra.sa[ib], sb = t.bra.sb[ib], sc = t.ket.sa[ik], sd = t.ket.sb[ik];
            const double Ax = t.bra.A[3 * ib], Ay = t.bra.A[3 * ib + 1], Az = t.bra.A[3 * ib + 2];
            const double ABx = t.bra.AB[3 * ib], ABy = t.bra.AB[3 * ib + 1], ABz = t.bra.AB[3 * ib + 2];
            const double Cx = t.ket.A[3 * ik], Cy = t.ket.A[3 * ik + 1], Cz = t.ket.A[3 * ik + 2];
            const double CDx = t.ket.AB[3 * ik], CDy = t.ket.AB[3 * ik + 1], CDz = t.ket.AB[3 * ik + 2];
            if (thr > 0.0 && !(t.bra.Q[ib] * t.ket.Q[ik] > thr)) continue;   // uniform across the CTA
            if (lane == 0) cnt_q++;
            const int pab0 = t.bra.pbase[ib], npab = t.bra.nprim[ib];
            const int pcd0 = t.ket.pbase[ik], npcd = t.ket.nprim[ik];
            double wgt = (sa == sb ? 1.0 : 2.0) * (sc == sd ? 1.0 : 2.0);
            wgt *= (t.same_class && ib == ik) ? 1.0 : 2.0;

            double gout[NPL];
#pragma unroll
            for (int m = 0; m < NPL; m++) gout[m] = 0.0;

            for (int iab = 0; iab < npab; iab++) {
                const int sab = pab0 + iab * CF_PSTRIDE;
                const double p = t.bra.p[sab], cab = t.bra.c[sab];
                const double Px = t.bra.Px[sab], Py = t.bra.Py[sab], Pz = t.bra.Pz[sab];
                for (int icd = 0; icd < npcd; icd++) {
                    const int scd = pcd0 + icd * CF_PSTRIDE;
                    const double ccd = t.ket.c[scd];
                    if (fabs(cab * ccd) < t.prim_cut) continue;   // uniform across the CTA
                    if (lane == 0) cnt_p++;
                    const double qe = t.ket.p[scd];
                    const double Qx = t.ket.Px[scd], Qy = t.ket.Py[scd], Qz = t.ket.Pz[scd];
                    const double pq = p + qe;
                    const double rho = p * qe / pq;
                    const double PQx = Px - Qx, PQy = Py - Qy, PQz = Pz - Qz;
                    const double T = rho * (PQx * PQx + PQy * PQy + PQz * PQz);
                    if (lane < 2 * NROOTS) rw[lane] = rys_value<NROOTS>(t.rys, T, lane);
                    __syncthreads();
                    if (lane < 3 * NROOTS) {
                        const int r = lane / 3, dim = lane - 3 * r;
                        const double x = rw[r];
                        const double rx_p = rho * x / p;       // rho t^2 / p
                        const double rx_q = rho * x / qe;
                        const double b00 = 0.5 * x / pq;
                        const double b10 = (1.0 - rx_p) * (0.5 / p);
                        const double b01 = (1.0 - rx_q) * (0.5 / qe);
                        double PA, PQ, QC, ab, cd, w0;
                        if (dim == 0) { PA = Px - Ax; PQ = PQx; QC = Qx - Cx; ab = ABx; cd = CDx; w0 = 1.0; }
                        else if (dim == 1) { PA = Py - Ay; PQ = PQy; QC = Qy - Cy; ab = ABy; cd = CDy; w0 = 1.0; }
                        else { PA = Pz - Az; PQ = PQz; QC = Qz - Cz; ab = ABz; cd = CDz;
                               w0 = rw[NROOTS + r] * cab * ccd * rsqrt(pq) * wgt; }
                        const double c00 = PA - rx_p * PQ;
                        const double c00p = QC + rx_q * PQ;
                        rys_2d<LA, LB, LC, LD>(w0, c00, c00p, b10, b01, b00, ab, cd, g + (size_t)lane * GSZ);
                    }
                    __syncthreads();
#pragma unroll
                    for (int m = 0; m < NPL; m++) {
                        const int ix = idx3[m] & 1023, iy = (idx3[m] >> 10) & 1023, iz = idx3[m] >> 20;
                        double s = 0.0;
#pragma unroll
                        for (int r = 0; r < NROOTS; r++)
                            s = fma(g[(3 * r) * GSZ + ix] * g[(3 * r + 1) * GSZ + iy], g[(3 * r + 2) * GSZ + iz], s);
                        gout[m] += s;
                    }
                    // the next primitive quartet overwrites rw/g: the barrier after phase A orders rw,
                    // and g is only rewritten after that barrier as well
                }
            }

            if (STORE) {
#pragma unroll
                for (int m = 0; m < NPL; m++) {
                    const int n = lane + m * G;
                    if (n < NOUT) t.store[(size_t)q * NOUT + n] = gout[m] / wgt;
                }
                __syncthreads();
                continue;
            }

            // ---- phase D: digestion ------------------------------------------------------------
            const int ca = t.bra.cao_a[ib], cb = t.bra.cao_b[ib], cc = t.ket.cao_a[ik], cdd = t.ket.cao_b[ik];
            const int ld = t.ncart;
#pragma unroll
            for (int m = 0; m < NPL; m++) {
                const int n = lane + m * G;
                if (n < NOUT) V[n] = gout[m];
            }
            for (int e = lane; e < NAB; e += G) Dab[e] = t.Dtot[(size_t)(cb + e % NB) * ld + ca + e / NB];
            for (int e = lane; e < NCD; e += G) Dcd[e] = t.Dtot[(size_t)(cdd + e % ND) * ld + cc + e / ND];
            for (int x = 0; x < t.nk; x++) {
                const double* D = t.Dk[x];
                double* d = Dx + x * NDX;
                for (int e = lane; e < NA * NC; e += G) d[e] = D[(size_t)(ca + e / NC) * ld + cc + e % NC];
                d += NA * NC;
                for (int e = lane; e < NA * ND; e += G) d[e] = D[(size_t)(ca + e / ND) * ld + cdd + e % ND];
                d += NA * ND;
                for (int e = lane; e < NB * NC; e += G) d[e] = D[(size_t)(cb + e / NC) * ld + cc + e % NC];
                d += NB * NC;
                for (int e = lane; e < NB * ND; e += G) d[e] = D[(size_t)(cb + e / ND) * ld + cdd + e % ND];
            }
            __syncthreads();
            // J targets: NAB (bra block) + NCD (ket block)
            for (int e = lane; e < NAB + NCD; e += G) {
                double s = 0.0;
                if (e < NAB) {
                    for (int kl = 0; kl < NCD; kl++) s = fma(V[e * NCD + kl], Dcd[kl], s);
                    fixed_add_j(t.accJ + (size_t)(cb + e % NB) * ld + ca + e / NB, jlo, s, scaleJ);
                } else {
                    const int kl = e - NAB;
                    for (int ij = 0; ij < NAB; ij++) s = fma(V[ij * NCD + kl], Dab[ij], s);
                    fixed_add_j(t.accJ + (size_t)(cdd + kl % ND) * ld + cc + kl / ND, jlo, s, scaleJ);
                }
            }
            for (int xj = 1; xj < t.nj; xj++) {   // further Coulomb densities of the multi-density build: D read from global memory
                const double* DJ = t.Dj[xj];
                long long* aJ = t.accJm[xj];
                for (int e = lane; e < NAB + NCD; e += G) {
                    double s = 0.0;
                    if (e < NAB) {
                        for (int kl = 0; kl < NCD; kl++) s = fma(V[e * NCD + kl], DJ[(size_t)(cdd + kl % ND) * ld + cc + kl / ND], s);
                        fixed_add_j(aJ + (size_t)(cb + e % NB) * ld + ca + e / NB, jlo, s, scaleJ);
                    } else {
                        const int kl = e - NAB;
                        for (int ij = 0; ij < NAB; ij++) s = fma(V[ij * NCD + kl], DJ[(size_t)(cb + ij % NB) * ld + ca + ij / NB], s);
                        fixed_add_j(aJ + (size_t)(cdd + kl % ND) * ld + cc + kl / ND, jlo, s, scaleJ);
                    }
                }
            }
            // K targets, per density: ac, ad, bc, bd
            for (int x = 0; x < t.nk; x++) {
                const double* dac = Dx + x * NDX;
                const double* dad = dac + NA * NC;
                const double* dbc = dad + NA * ND;
                const double* dbd = dbc + NB * NC;
                long long* acc = t.accK[x];
                for (int e = lane; e < NDX; e += G) {
                    double s = 0.0;
                    if (e < NA * NC) {                       // K(i,k) += sum_jl V[i,j,k,l] D(j,l)
                        const int i = e / NC, k = e % NC;
                        for (int j = 0; j < NB; j++)
                            for (int l = 0; l < ND; l++) s = fma(V[((i * NB + j) * NC + k) * ND + l], dbd[j * ND + l], s);
                        fixed_add(acc + (size_t)(ca + i) * ld + cc + k, s, scaleK);
                    } else if (e < NA * NC + NA * ND) {      // K(i,l) += sum_jk V D(j,k)
                        const int f = e - NA * NC, i = f / ND, l = f % ND;
                        for (int j = 0; j < NB; j++)
                            for (int k = 0; k < NC; k++) s = fma(V[((i * NB + j) * NC + k) * ND + l], dbc[j * NC + k], s);
                        fixed_add(acc + (size_t)(ca + i) * ld + cdd + l, s, scaleK);
                    } else if (e < NA * NC + NA * ND + NB * NC) {   // K(j,k) += sum_il V D(i,l)
                        const int f = e - NA * NC - NA * ND, j = f / NC, k = f % NC;
                        for (int i = 0; i < NA; i++)
                            for (int l = 0; l < ND; l++) s = fma(V[((i * NB + j) * NC + k) * ND + l], dad[i * ND + l], s);
                        fixed_add(acc + (size_t)(cb + j) * ld + cc + k, s, scaleK);
                    } else {                                  // K(j,l) += sum_ik V D(i,k)
                        const int f = e - NA * NC - NA * ND - NB * NC, j = f / ND, l = f % ND;
                        for (int i = 0; i < NA; i++)
                            for (int k = 0; k < NC; k++) s = fma(V[((i * NB + j) * NC + k) * ND + l], dac[i * NC + k], s);
                        fixed_add(acc + (size_t)(cb + j) * ld + cdd + l, s, scaleK);
                    }
                }
            }
            __syncthreads();   // V / D blocks are reused by the next quartet
        }
    }
    if (!STORE && threadIdx.x < 32) cf_cnt_flush(t.cnt, cnt_q, cnt_p);
}

template <int LA, int LB, int LC, int LD>
constexpr size_t eri_generic_smem(int nk) {
    constexpr int NA = cf_ncart(LA), NB = cf_ncart(LB), NC = cf_ncart(LC), ND = cf_ncart(LD);
    constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    constexpr int GSZ = (LA + 1) * (LB + 1) * (LC + 1) * (LD + 1);
    return sizeof(double) * (size_t)(2 * NROOTS + NROOTS * 3 * GSZ + NA * NB * NC * ND + NA * NB + NC * ND +
                                     (nk > 0 ? nk : 0) * (NA * NC + NA * ND + NB * NC + NB * ND));
}
