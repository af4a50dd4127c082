// engine.cu -- host side of the B200 J/K engine + the C ABI of include/chinium_fock.h.
//
// Replaces the reference's Int4C2E setup pipeline and ContractInts
// (src/Integral/Int4C2E.cpp:494-587, :673-683).  Design:
//   * shell pairs grouped by angular class (la>=lb), primitive-pair data SoA in HBM
//   * Schwarz bounds from the SAME quartet kernels run on the (ab|ab) diagonal
//   * J/K digested in the CARTESIAN working basis: D_cart = C^T D C on the way in,
//     J = C J_cart C^T on the way out (C = Racah solid-harmonic coefficients), so the ERI
//     kernels never do a 4-index cart->pure transform
//   * accumulators are 64-bit fixed point with a power-of-two scale derived on the device from a
//     rigorous Schwarz bound -> integer adds commute -> results are bit-identical for any launch
//     geometry, stream interleaving and number of GPUs
// No CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <dlfcn.h>
#include <cub/cub.cuh>         // device radix sort / scan of the setup (pair order, bra-role order)
#include <nccl.h>              // types and prototypes only: libnccl.so.2 is bound with dlopen when a multi-device handle is created
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/chinium_fock.h"
#include "cf_common.cuh"
#include "eri_generic.cuh"     // rys_tmax/rys_off (host constexpr) -- no kernels instantiated here
#include "eri_tpq.cuh"         // TPQ_THREADS
#include "eri_grad.cuh"        // GradTask (kernels are instantiated in eri_inst.cu)
#include "oneint.cuh"          // one-electron integrals S, T, V
#include "rys_tables_data.h"

#ifndef CF_BRALOOP_MAXK
#define CF_BRALOOP_MAXK 8.0    // mean primitive quartets per shell quartet below which the bra-loop kernel is used
#endif

#define CUDA_TRY(x)                                                                             \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            set_error(h, std::string(#x) + ": " + cudaGetErrorString(e_));                      \
            return CF_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

// launchers from eri_inst.cu (one translation unit per bra class)
#define DECL_BRA(n) cudaError_t cf_launch_bra##n(int, const QuartetTask&, int, int, cudaStream_t, int*, size_t*, int*);
DECL_BRA(0) DECL_BRA(1) DECL_BRA(2) DECL_BRA(3) DECL_BRA(4) DECL_BRA(5) DECL_BRA(6) DECL_BRA(7) DECL_BRA(8) DECL_BRA(9)
typedef cudaError_t (*bra_launch_fn)(int, const QuartetTask&, int, int, cudaStream_t, int*, size_t*, int*);
#define DECL_GRAD(n) cudaError_t cf_launch_grad_bra##n(int, const GradTask&, int, cudaStream_t, int*, size_t*);
DECL_GRAD(0) DECL_GRAD(1) DECL_GRAD(2) DECL_GRAD(3) DECL_GRAD(4) DECL_GRAD(5) DECL_GRAD(6) DECL_GRAD(7) DECL_GRAD(8) DECL_GRAD(9)
typedef cudaError_t (*grad_launch_fn)(int, const GradTask&, int, cudaStream_t, int*, size_t*);
#define DECL_GRADMAT(n) cudaError_t cf_launch_gradmat_bra##n(int, const GradTask&, int, cudaStream_t, int*, size_t*);
DECL_GRADMAT(0) DECL_GRADMAT(1) DECL_GRADMAT(2) DECL_GRADMAT(3) DECL_GRADMAT(4) DECL_GRADMAT(5) DECL_GRADMAT(6) DECL_GRADMAT(7) DECL_GRADMAT(8) DECL_GRADMAT(9)
static grad_launch_fn g_gradmat_launch[CF_NCLS] = {cf_launch_gradmat_bra0, cf_launch_gradmat_bra1, cf_launch_gradmat_bra2, cf_launch_gradmat_bra3, cf_launch_gradmat_bra4,
                                                   cf_launch_gradmat_bra5, cf_launch_gradmat_bra6, cf_launch_gradmat_bra7, cf_launch_gradmat_bra8, cf_launch_gradmat_bra9};
#define DECL_HESS(n) cudaError_t cf_launch_hess_bra##n(int, const GradTask&, int, cudaStream_t, int*, size_t*);
DECL_HESS(0) DECL_HESS(1) DECL_HESS(2) DECL_HESS(3) DECL_HESS(4) DECL_HESS(5) DECL_HESS(6) DECL_HESS(7) DECL_HESS(8) DECL_HESS(9)
static grad_launch_fn g_hess_launch[CF_NCLS] = {cf_launch_hess_bra0, cf_launch_hess_bra1, cf_launch_hess_bra2, cf_launch_hess_bra3, cf_launch_hess_bra4,
                                                cf_launch_hess_bra5, cf_launch_hess_bra6, cf_launch_hess_bra7, cf_launch_hess_bra8, cf_launch_hess_bra9};
static grad_launch_fn g_grad_launch[CF_NCLS] = {cf_launch_grad_bra0, cf_launch_grad_bra1, cf_launch_grad_bra2, cf_launch_grad_bra3, cf_launch_grad_bra4,
                                                cf_launch_grad_bra5, cf_launch_grad_bra6, cf_launch_grad_bra7, cf_launch_grad_bra8, cf_launch_grad_bra9};
static bra_launch_fn g_bra_launch[CF_NCLS] = {cf_launch_bra0, cf_launch_bra1, cf_launch_bra2, cf_launch_bra3, cf_launch_bra4,
                                              cf_launch_bra5, cf_launch_bra6, cf_launch_bra7, cf_launch_bra8, cf_launch_bra9};

static std::string g_last_error;

// primitive-quartet cutoff of the J/K kernels: |c_ab c_cd wgt| below it is skipped.  Developer knob CF_PRIM_CUT for the
// accuracy/time scan of tools/prim_cut_scan.py (read once); the default is what every parity test runs with.
// 1e-18 from the scan of session r2l (profiles/r2l_primcut.txt; densities with O(1) entries, |J|max 110 - 620): against 1e-22
// max|dJ| 5.1e-13 / max|dK| 1.7e-13 on c18, 2.4e-14 / 3.3e-14 on fe4s4, 1.4e-14 / 1.7e-13 on (H2O)64 -- 200x below the 1e-10 bar --
// for 17 % / 10 % / 2 % fewer primitive quartets (c18 -7.5 %, fe4s4 -4 %); 1e-16 would leave only a factor 3 (2.9e-11).
#define CF_PRIM_CUT_DEFAULT 1e-18
static double cf_prim_cut() {
    static double v = -1.0;
    if (v < 0.0) { const char* e = getenv("CF_PRIM_CUT"); v = e ? atof(e) : CF_PRIM_CUT_DEFAULT; if (!(v >= 0.0)) v = CF_PRIM_CUT_DEFAULT; }
    return v;
}

// Device buffer.  Setup allocates a few hundred arrays (per class, per task): they come from the device's stream-ordered
// memory pool (cudaMallocAsync on the default stream; the pool keeps freed blocks, so repeated handles cost microseconds
// per allocation instead of a cudaMalloc each).  `pooled = false` (plain cudaMalloc) is used for the accumulators that
// NCCL reads and writes over NVLink.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool pooled = true;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        if (!pooled) return cudaMalloc(&p, count * sizeof(T));
        cudaError_t e = cudaMallocAsync(&p, count * sizeof(T), 0);
        if (e != cudaSuccess) { (void)cudaGetLastError(); pooled = false; e = cudaMalloc(&p, count * sizeof(T)); }
        return e;
    }
    cudaError_t upload(const std::vector<T>& v) {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release() {
        if (p) { if (pooled) cudaFreeAsync(p, 0); else cudaFree(p); }
        p = nullptr; n = 0;
    }
};

// basis arrays on the device (pair build and primitive screening run there, cf_create)
struct BasisDev {
    const int* l; const int* nprim; const int* prim_off; const int* cao_off;
    const double* exps; const double* coefs; const double* xyz;
};

// One thread per canonical shell pair e = s1 (s1+1)/2 + s2, s2 <= s1 (replaces the reference's pair loops,
// Int4C2E.cpp:79-231, and round 1's host loop): angular class and number of primitive pairs whose prefactor
// ca cb exp(-ab/p |AB|^2) sqrt(2) pi^(5/4) / p survives the cutoff.  cls 255 = no primitive left (pair dropped).
__global__ void pair_classify_kernel(int ns, long long npairs, BasisDev b, double cutoff, unsigned char* __restrict__ cls, int* __restrict__ kept) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= npairs) return;
    long long s1 = (long long)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while (s1 * (s1 + 1) / 2 > e) s1--;
    while ((s1 + 1) * (s1 + 2) / 2 <= e) s1++;
    int a = (int)s1, c = (int)(e - s1 * (s1 + 1) / 2);
    if (b.l[c] > b.l[a]) { const int t = a; a = c; c = t; }
    const double dx = b.xyz[3 * a] - b.xyz[3 * c], dy = b.xyz[3 * a + 1] - b.xyz[3 * c + 1], dz = b.xyz[3 * a + 2] - b.xyz[3 * c + 2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double pref = 5.9149671727956128778;   // sqrt(2) pi^(5/4)
    int n = 0;
    for (int i = 0; i < b.nprim[a]; i++)
        for (int j = 0; j < b.nprim[c]; j++) {
            const double ea = b.exps[b.prim_off[a] + i], eb = b.exps[b.prim_off[c] + j], p = ea + eb;
            const double cc = b.coefs[b.prim_off[a] + i] * b.coefs[b.prim_off[c] + j] * exp(-ea * eb / p * r2) * pref / p;
            if (fabs(cc) >= cutoff) n++;
        }
    kept[e] = n;
    cls[e] = n ? (unsigned char)cf_pair_class(b.l[a], b.l[c]) : (unsigned char)255;
}

// Primitive-pair data of one class in device order: one thread per pair writes its SoA entries and its surviving primitives
// into the interleaved slots pbase[i] + k * CF_PSTRIDE (padding slots were pre-filled with p = 1, c = 0).
__global__ void pair_fill_kernel(int np, const int* __restrict__ sa, const int* __restrict__ sb, const int* __restrict__ pbase,
                                 const int* __restrict__ nkeep, BasisDev b,
                                 double cutoff, int* __restrict__ cao_a, int* __restrict__ cao_b, double* __restrict__ A, double* __restrict__ AB,
                                 double* __restrict__ p_, double* __restrict__ hp, double* __restrict__ Px, double* __restrict__ Py,
                                 double* __restrict__ Pz, double* __restrict__ c_, double* __restrict__ aexp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const int a = sa[i], c = sb[i];
    const double ax = b.xyz[3 * a], ay = b.xyz[3 * a + 1], az = b.xyz[3 * a + 2];
    const double bx = b.xyz[3 * c], by = b.xyz[3 * c + 1], bz = b.xyz[3 * c + 2];
    const double dx = ax - bx, dy = ay - by, dz = az - bz, r2 = dx * dx + dy * dy + dz * dz;
    cao_a[i] = b.cao_off[a]; cao_b[i] = b.cao_off[c];
    A[3 * (size_t)i] = ax; A[3 * (size_t)i + 1] = ay; A[3 * (size_t)i + 2] = az;
    AB[3 * (size_t)i] = dx; AB[3 * (size_t)i + 1] = dy; AB[3 * (size_t)i + 2] = dz;
    const double pref = 5.9149671727956128778;
    size_t slot = (size_t)pbase[i];
    int left = nkeep[i];       // the slots were sized by pair_classify_kernel's count: never write more (same arithmetic, same answer)
    for (int ia = 0; ia < b.nprim[a]; ia++)
        for (int jb = 0; jb < b.nprim[c]; jb++) {
            const double ea = b.exps[b.prim_off[a] + ia], eb = b.exps[b.prim_off[c] + jb], p = ea + eb;
            const double cc = b.coefs[b.prim_off[a] + ia] * b.coefs[b.prim_off[c] + jb] * exp(-ea * eb / p * r2) * pref / p;
            if (fabs(cc) < cutoff || left <= 0) continue;
            left--;
            p_[slot] = p; hp[slot] = 0.5 / p; c_[slot] = cc; aexp[slot] = ea;
            Px[slot] = (ea * ax + eb * bx) / p; Py[slot] = (ea * ay + eb * by) / p; Pz[slot] = (ea * az + eb * bz) / p;
            slot += CF_PSTRIDE;
        }
}
__global__ void fill_value_kernel(size_t n, double v, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}
// sort keys: pair order inside a class = (primitive count descending, Schwarz bound descending); bra-role order = (shell a
// ascending, Schwarz bound descending).  Q >= 0, so the bit pattern of the double orders like the value; its top bits
// are complemented for "descending".  Radix sort is stable and deterministic: ties keep creation order on every rank.
__global__ void pair_sort_key_kernel(int np, const int* __restrict__ nprim, const double* __restrict__ Q, unsigned long long* __restrict__ key, int* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const unsigned long long qb = (unsigned long long)__double_as_longlong(fmax(Q[i], 0.0));
    const unsigned long long npk = (unsigned long long)(4095 - min(nprim[i], 4095));
    key[i] = (npk << 52) | ((~qb >> 12) & ((1ull << 52) - 1));
    val[i] = i;
}
__global__ void role_sort_key_kernel(int np, const int* __restrict__ sa, const double* __restrict__ Q, unsigned long long* __restrict__ key, int* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const unsigned long long qb = (unsigned long long)__double_as_longlong(fmax(Q[i], 0.0));
    key[i] = ((unsigned long long)sa[i] << 48) | ((~qb >> 16) & ((1ull << 48) - 1));
    val[i] = i;
}
__global__ void desc_key_kernel(int np, const double* __restrict__ Q, unsigned long long* __restrict__ key, int* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    key[i] = ~(unsigned long long)__double_as_longlong(fmax(Q[i], 0.0));
    val[i] = i;
}
template <class T>
__global__ void gather_kernel(int n, const int* __restrict__ idx, const T* __restrict__ in, T* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}
// bra-role copy of a class (eri_tpqa.cuh): records + CONTIGUOUS primitives in `border` order, gathered from the slot arrays
__global__ void role_fill_kernel(int np, const int* __restrict__ border, const int* __restrict__ off, long long tot, const int* __restrict__ sb,
                                 const int* __restrict__ cao_b, const int* __restrict__ nprim, const int* __restrict__ pbase,
                                 const double* __restrict__ Q, const double* __restrict__ AB, const double* __restrict__ p,
                                 const double* __restrict__ hp, const double* __restrict__ Px, const double* __restrict__ Py,
                                 const double* __restrict__ Pz, const double* __restrict__ c, int4* __restrict__ ri, double* __restrict__ rd,
                                 double* __restrict__ bp) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= np) return;
    const int pos = border[e], n = nprim[pos], o = off[e];
    ri[e] = make_int4(pos, sb[pos] | (n << 16), cao_b[pos], o);
    rd[4 * (size_t)e] = Q[pos];
    for (int x = 0; x < 3; x++) rd[4 * (size_t)e + 1 + x] = AB[3 * (size_t)pos + x];
    for (int k = 0; k < n; k++) {
        const size_t src = (size_t)pbase[pos] + (size_t)k * CF_PSTRIDE;
        bp[o + k] = p[src]; bp[tot + o + k] = hp[src]; bp[2 * tot + o + k] = Px[src]; bp[3 * tot + o + k] = Py[src];
        bp[4 * tot + o + k] = Pz[src]; bp[5 * tot + o + k] = c[src];
    }
}

// stable device radix sort of (key, value) pairs; returns the sorted values in `vals` (keys are scratch)
static cudaError_t device_sort_pairs(int n, unsigned long long* keys, int* vals) {
    if (n <= 1) return cudaSuccess;
    unsigned long long* k2 = nullptr; int* v2 = nullptr; void* tmp = nullptr; size_t bytes = 0;
    cudaError_t e;
    if ((e = cudaMalloc(&k2, sizeof(unsigned long long) * n)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&v2, sizeof(int) * n)) != cudaSuccess) { cudaFree(k2); return e; }
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, k2, vals, v2, n);
    if ((e = cudaMalloc(&tmp, bytes)) == cudaSuccess) {
        e = cub::DeviceRadixSort::SortPairs(tmp, bytes, keys, k2, vals, v2, n);
        if (e == cudaSuccess) e = cudaMemcpy(vals, v2, sizeof(int) * n, cudaMemcpyDeviceToDevice);
    }
    cudaFree(tmp); cudaFree(k2); cudaFree(v2);
    return e;
}

struct PairClassHost {
    int la = 0, lb = 0;
    // host copies in DEVICE order (pairs sorted by (primitive count desc, Schwarz bound desc) once the bounds exist)
    std::vector<int> sa, sb, nprim, nprim_full;
    std::vector<double> Q, Qpure;    // Cartesian / pure-function Schwarz bounds
    std::vector<int> order;          // identity (kept for the position -> pair indirection of the host helpers)
    std::vector<int> seg;            // device positions where the primitive count changes (+ npair): Q descends inside a segment
    DevBuf<int> d_sa, d_sb, d_cao_a, d_cao_b, d_pbase, d_nprim, d_seg;
    DevBuf<double> d_A, d_AB, d_Q, d_Qcart, d_p, d_hp, d_Px, d_Py, d_Pz, d_c, d_aexp;
    // bra role of the bra-loop kernels (eri_tpqa.cuh): device positions grouped by shell a (Q descending inside a
    // group; `group` holds the group boundaries in `border`); ket role: max Q of every aligned block of 32 positions
    int nblk = 0;
    std::vector<int> border, group;
    DevBuf<int> d_border;
    DevBuf<int4> d_brec_i;
    DevBuf<double> d_blkQ, d_brec_d, d_bprim;
    long long bprim_stride = 0;
    int npair() const { return (int)sa.size(); }
    PairClassDev dev() const {
        PairClassDev d;
        d.npair = npair();
        d.sa = d_sa.p; d.sb = d_sb.p; d.cao_a = d_cao_a.p; d.cao_b = d_cao_b.p;
        d.pbase = d_pbase.p; d.nprim = d_nprim.p; d.A = d_A.p; d.AB = d_AB.p; d.Q = d_Q.p;
        d.p = d_p.p; d.hp = d_hp.p; d.Px = d_Px.p; d.Py = d_Py.p; d.Pz = d_Pz.p; d.c = d_c.p;
        return d;
    }
    // device layout for the current host order: blocks of CF_PSTRIDE pairs with interleaved primitives.  The host only
    // lays out the slot bases (a prefix sum over blocks); every primitive is generated by pair_fill_kernel on the device.
    cudaError_t layout_and_fill(const BasisDev& b, double cutoff) {
        const int np = npair();
        order.resize(np); std::iota(order.begin(), order.end(), 0);
        std::vector<int> pbase(np);
        size_t nslot = 0;
        for (int b0 = 0; b0 < np; b0 += CF_PSTRIDE) {
            int mx = 0;
            for (int i = b0; i < std::min(np, b0 + CF_PSTRIDE); i++) mx = std::max(mx, nprim[i]);
            for (int i = b0; i < std::min(np, b0 + CF_PSTRIDE); i++) pbase[i] = (int)(nslot + (i - b0));
            nslot += (size_t)mx * CF_PSTRIDE;
        }
        if (nslot > 0x7fffffffull) return cudaErrorInvalidValue;
        cudaError_t e;
        if ((e = d_sa.upload(sa)) != cudaSuccess) return e;
        if ((e = d_sb.upload(sb)) != cudaSuccess) return e;
        if ((e = d_pbase.upload(pbase)) != cudaSuccess) return e;
        if ((e = d_nprim.upload(nprim)) != cudaSuccess) return e;
        if ((e = d_cao_a.alloc(np)) != cudaSuccess || (e = d_cao_b.alloc(np)) != cudaSuccess) return e;
        if ((e = d_A.alloc(3 * (size_t)np)) != cudaSuccess || (e = d_AB.alloc(3 * (size_t)np)) != cudaSuccess) return e;
        if (d_Q.n != (size_t)np && (e = d_Q.alloc(np)) != cudaSuccess) return e;
        DevBuf<double>* slots[7] = {&d_p, &d_hp, &d_Px, &d_Py, &d_Pz, &d_c, &d_aexp};
        const double fillv[7] = {1.0, 0.5, 0.0, 0.0, 0.0, 0.0, 0.5};
        for (int k = 0; k < 7; k++) {
            if ((e = slots[k]->alloc(nslot)) != cudaSuccess) return e;
            if (nslot) fill_value_kernel<<<(unsigned)((nslot + 255) / 256), 256>>>(nslot, fillv[k], slots[k]->p);
        }
        seg.clear();
        for (int i = 0; i < np; i++) if (i == 0 || nprim[i] != nprim[i - 1]) seg.push_back(i);
        seg.push_back(np);
        if ((e = d_seg.upload(seg)) != cudaSuccess) return e;
        if (np) pair_fill_kernel<<<(np + 127) / 128, 128>>>(np, d_sa.p, d_sb.p, d_pbase.p, d_nprim.p, b, cutoff, d_cao_a.p, d_cao_b.p, d_A.p, d_AB.p,
                                                             d_p.p, d_hp.p, d_Px.p, d_Py.p, d_Pz.p, d_c.p, d_aexp.p);
        return cudaGetLastError();
    }
    // after the Schwarz kernels: order = device radix sort by (primitive count desc, pure Schwarz bound desc); the host
    // copies and the device bounds are permuted accordingly, then the class is laid out again in that order
    cudaError_t sort_by_bounds(const BasisDev& b, double cutoff, const double* d_qpure) {
        const int np = npair();
        if (np == 0) return cudaSuccess;
        DevBuf<unsigned long long> key; DevBuf<int> val; DevBuf<double> q2;
        cudaError_t e;
        if ((e = key.alloc(np)) != cudaSuccess || (e = val.alloc(np)) != cudaSuccess || (e = q2.alloc(np)) != cudaSuccess) return e;
        pair_sort_key_kernel<<<(np + 255) / 256, 256>>>(np, d_nprim.p, d_qpure, key.p, val.p);
        if ((e = device_sort_pairs(np, key.p, val.p)) != cudaSuccess) return e;
        std::vector<int> perm(np);
        std::vector<double> qc(np), qp(np);
        gather_kernel<double><<<(np + 255) / 256, 256>>>(np, val.p, d_Qcart.p, q2.p);
        if ((e = cudaMemcpy(qc.data(), q2.p, sizeof(double) * np, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
        gather_kernel<double><<<(np + 255) / 256, 256>>>(np, val.p, d_qpure, q2.p);
        if ((e = cudaMemcpy(qp.data(), q2.p, sizeof(double) * np, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
        if ((e = cudaMemcpy(perm.data(), val.p, sizeof(int) * np, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
        auto permute = [&](std::vector<int>& v) { std::vector<int> t(np); for (int i = 0; i < np; i++) t[i] = v[perm[i]]; v.swap(t); };
        permute(sa); permute(sb); permute(nprim); permute(nprim_full);
        Q = qc; Qpure = qp;
        if ((e = layout_and_fill(b, cutoff)) != cudaSuccess) return e;
        if ((e = d_Qcart.upload(Q)) != cudaSuccess) return e;
        e = d_Q.upload(Qpure);
        key.release(); val.release(); q2.release();
        return e;
    }
    double qpos(int pos) const { return Qpure[pos]; }
    // bra / ket roles of the bra-loop kernels; needs the sorted layout and Qpure
    cudaError_t build_roles() {
        const int np = npair();
        nblk = (np + 31) / 32;
        border.resize(np); group.clear();
        if (np == 0) return cudaSuccess;
        DevBuf<unsigned long long> key; DevBuf<int> val, d_off;
        cudaError_t e;
        if ((e = key.alloc(np)) != cudaSuccess || (e = val.alloc(np)) != cudaSuccess) return e;
        role_sort_key_kernel<<<(np + 255) / 256, 256>>>(np, d_sa.p, d_Q.p, key.p, val.p);
        if ((e = device_sort_pairs(np, key.p, val.p)) != cudaSuccess) return e;
        if ((e = cudaMemcpy(border.data(), val.p, sizeof(int) * np, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
        for (int i = 0; i < np; i++) if (i == 0 || sa[border[i]] != sa[border[i - 1]]) group.push_back(i);
        group.push_back(np);
        std::vector<double> bq(nblk, 0.0);
        for (int i = 0; i < np; i++) bq[i / 32] = std::max(bq[i / 32], qpos(i));
        std::vector<int> off(np);
        size_t tot = 0;
        for (int i = 0; i < np; i++) {
            if (nprim[border[i]] >= 32768 || tot > 0x7fffffffull) return cudaErrorInvalidValue;
            off[i] = (int)tot; tot += nprim[border[i]];
        }
        bprim_stride = (long long)tot;
        if ((e = d_border.upload(border)) != cudaSuccess || (e = d_off.upload(off)) != cudaSuccess) return e;
        if ((e = d_brec_i.alloc(np)) != cudaSuccess || (e = d_brec_d.alloc(4 * (size_t)np)) != cudaSuccess || (e = d_bprim.alloc(6 * tot)) != cudaSuccess) return e;
        role_fill_kernel<<<(np + 127) / 128, 128>>>(np, d_border.p, d_off.p, (long long)tot, d_sb.p, d_cao_b.p, d_nprim.p, d_pbase.p, d_Q.p, d_AB.p,
                                                     d_p.p, d_hp.p, d_Px.p, d_Py.p, d_Pz.p, d_c.p, d_brec_i.p, d_brec_d.p, d_bprim.p);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        e = d_blkQ.upload(bq);
        key.release(); val.release(); d_off.release();
        return e;
    }
    // chunks of consecutive entries of `border` inside one a-group: at most lch pairs and (after the first pair) at most
    // pcap primitive pairs, so that the cost of a work item is bounded; heaviest chunks first (static schedule, short tail)
    void make_chunks(int lch, double pcap, std::vector<int>& cstart, std::vector<int>& ccnt, std::vector<int>& cmax, std::vector<double>& cq) const {
        struct Ch { int start, cnt, maxpos; double q, cost; };
        std::vector<Ch> ch;
        for (size_t g = 0; g + 1 < group.size(); g++)
            for (int c0 = group[g]; c0 < group[g + 1];) {
                int n = 0, mx = 0; double cost = 0;
                while (c0 + n < group[g + 1] && n < lch) {
                    const double np1 = nprim[border[c0 + n]];
                    if (n > 0 && cost + np1 > pcap) break;
                    cost += np1; mx = std::max(mx, border[c0 + n]); n++;
                }
                ch.push_back({c0, n, mx, qpos(border[c0]), cost});
                c0 += n;
            }
        std::stable_sort(ch.begin(), ch.end(), [](const Ch& x, const Ch& y) { return x.cost > y.cost; });
        for (const Ch& c : ch) { cstart.push_back(c.start); ccnt.push_back(c.cnt); cmax.push_back(c.maxpos); cq.push_back(c.q); }
    }
    void release() {
        d_border.release(); d_blkQ.release(); d_brec_i.release(); d_brec_d.release(); d_bprim.release();
        d_sa.release(); d_sb.release(); d_cao_a.release(); d_cao_b.release(); d_pbase.release(); d_nprim.release(); d_seg.release();
        d_A.release(); d_AB.release(); d_Q.release(); d_Qcart.release(); d_p.release(); d_hp.release(); d_Px.release(); d_Py.release();
        d_Pz.release(); d_c.release(); d_aexp.release();
    }
};

struct ClassPairTask {
    int bra = 0, ket = 0;
    long long nquartet = 0;          // all canonical quartets of the class pair (before Schwarz screening)
    DevBuf<long long> d_qoff;        // generic kernels: first quartet of each bra pair
    DevBuf<int4> d_items;            // kinds 1,2,3: work items (uniform pair, first spread pair, count), built on the device
    long long nitem = 0;
    int kind = 0;                    // 0 generic (CTA per quartet), 1 thread per quartet, 2 sliced thread per quartet
    int nq_item = 0;                 // kinds 1,2,3: ket pairs per work item
    int swap = 0;                    // kind 3: the LOWER class is handed to the kernel as the CTA-uniform pair
    int braloop = 0;                 // kind 1: bra-loop kernel (eri_tpqa.cuh) with chunk items
    int G = 32;
    size_t smem[4] = {0, 0, 0, 0};   // by nk
    double flops_eri = 0;            // F_alg without the digestion term (NOMINAL contraction depths)
    double nfun_sum = 0;             // sum over quartets of N_s (pure functions) for the digestion term
    double per_prim = 0;             // model flops of ONE primitive quartet of this class pair
    double nfun_q = 0;               // N_s of one quartet
    int index = 0;                   // position in cf_handle::tasks = slot of the task's device counters
    int owner = -1;                  // multi-GPU: -1 = items dealt round-robin over all ranks; r >= 0 = rank r runs the whole (small) task
};

struct cf_handle {
    int device = 0;
    cf_options opt{};
    int nshell = 0, nbf = 0, ncart = 0;
    std::vector<int> type, l, nprim, prim_off, bf_off, cao_off, nfun;
    std::vector<double> exps, coefs, xyz;
    std::vector<int> shell2atom;           // empty when the caller gave none (gradients then refuse)
    DevBuf<int> d_shell2atom, d_l, d_nprim_sh, d_prim_off_sh;
    DevBuf<double> d_exps, d_coefs, d_xyz, d_atomZ, d_atomxyz;
    DevBuf<double> d_gpart, d_grad, d_hess;
    // per-shell transformation (function x cartesian), pooled by type
    std::vector<double> ctrans;            // pool
    std::vector<int> ct_off;               // [nshell] offset into pool
    DevBuf<double> d_ctrans;
    DevBuf<int> d_ct_off, d_bf_off, d_cao_off, d_nfun, d_ncartsh;
    PairClassHost cls[CF_NCLS];
    std::vector<ClassPairTask*> tasks;
    DevBuf<double> d_rys_table, d_rys_asym, d_boys;
    // per-build work space
    DevBuf<double> d_Dpure[3], d_Dcart[4] /* 0: Dtot, 1..3: Dk */, d_out[4] /* pure J,Kd,Ka,Kb */, d_partial, d_diag,
        d_QS /* [nshell^2] Cartesian Schwarz bound per shell pair */, d_B /* [4][nshell^2] block 1-norms of |D_cart| */, d_Bmax /* same, max-norms */, d_rwork;
    DevBuf<long long> d_acc;
    DevBuf<unsigned long long> d_cnt;        // [ntask][CF_CNT_WORDS] work counters of the last build (cf_common.cuh)
    std::vector<unsigned long long> cnt_host;
    double qmax_cart = 0;
    double scales_host[8] = {1, 1, 0, 0, 0, 0, 0, 0};   // copy of the scale tail of the last synchronised build
    const long long* last_tail = nullptr;    // device address of that tail (the accumulator of the last *_device call)
    int last_nk = 0;                         // exchange densities of the last build (flops_executed_last)
    double density_threshold = 0.0;          // > 0: density-weighted screening (cf_set_density_threshold)
    cudaEvent_t ev[4];
    cudaStream_t side[3];
    cudaEvent_t ev_fork, ev_join[3];
    cf_stats stats{};
    long long ref_counts[2] = {0, 0};        // the reference's RepulsionLength / ShellQuartetLength
    std::string err;
    bool diag_ready = false;
    struct MultiCtx* multi = nullptr;        // != nullptr: this handle fronts one partition handle per device (cf_create_multi)
};

static void set_error(cf_handle* h, const std::string& s) {
    if (h) h->err = s;
    g_last_error = s;
}

static int multi_build_jk(cf_handle* h, int nbf, const double* Dd, const double* Da, const double* Db, double exx,
                          double* J, double* Kd, double* Ka, double* Kb);
static int multi_build_g(cf_handle* h, int nbf, int nmat, const double* Ds, double exx, double* Gs);
static int multi_contract_grads(cf_handle* h, int nbf, const double* D1, const double* D2, double exx, int natom, double* grad);
static int multi_contract_hess(cf_handle* h, int nbf, const double* D, double exx, int natom, double* hess);
static void multi_destroy(cf_handle* h);
static cf_handle* multi_part0(const cf_handle* h);
static void multi_set_density_threshold(cf_handle* h, double dthr);

// every entry point works on the handle's device and leaves the caller's current device as it found it
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
// accumulator layout (int64 words): [J | K_0 .. K_{nk-1} | J low limb | tail], tail = CF_ACC_TAIL words holding the
// fixed-point scales of THAT build as doubles -- scales travel with the accumulator they belong to
#define CF_ACC_TAIL 8

// ------------------------------------------------------------------------------------------------
// small device kernels
// ------------------------------------------------------------------------------------------------
// D_cart(block sa,sb) = C_a^T Dsym(block) C_b, with Dsym = (D + D^T)/2 ; optionally Dtot = 2Dd + Da + Db
__global__ void pure_to_cart_kernel(int nshell, int nbf, int ncart, const double* __restrict__ Dd, const double* __restrict__ Da,
                                    const double* __restrict__ Db, double fd, double fa, double fb,
                                    const double* __restrict__ ctrans, const int* __restrict__ ct_off, const int* __restrict__ bf_off,
                                    const int* __restrict__ cao_off, const int* __restrict__ nfun, const int* __restrict__ ncsh,
                                    double* __restrict__ out, double* __restrict__ bnorm, double* __restrict__ bmax) {
    const int sa = blockIdx.x, sb = blockIdx.y;
    __shared__ double sh[64], shm[64];
    double asum = 0.0, amax = 0.0;
    const int na = nfun[sa], nb = nfun[sb], nca = ncsh[sa], ncb = ncsh[sb];
    const double* Ca = ctrans + ct_off[sa];
    const double* Cb = ctrans + ct_off[sb];
    for (int e = threadIdx.x; e < nca * ncb; e += blockDim.x) {
        const int x = e / ncb, y = e % ncb;
        double s = 0.0;
        for (int m = 0; m < na; m++) {
            const double cam = Ca[m * nca + x];
            if (cam == 0.0) continue;
            for (int n = 0; n < nb; n++) {
                const size_t i = bf_off[sa] + m, j = bf_off[sb] + n;
                double d = 0.0;
                if (Dd) d += fd * 0.5 * (Dd[j * nbf + i] + Dd[i * nbf + j]);
                if (Da) d += fa * 0.5 * (Da[j * nbf + i] + Da[i * nbf + j]);
                if (Db) d += fb * 0.5 * (Db[j * nbf + i] + Db[i * nbf + j]);
                s = fma(cam * Cb[n * ncb + y], d, s);
            }
        }
        out[(size_t)(cao_off[sb] + y) * ncart + cao_off[sa] + x] = s;
        asum += fabs(s);
        amax = fmax(amax, fabs(s));
    }
    // entrywise 1-norm of the Cartesian shell block (fixed-order tree: deterministic), used by the fixed-point scale bound
    sh[threadIdx.x] = asum; shm[threadIdx.x] = amax;
    __syncthreads();
    for (int w = 32; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sh[threadIdx.x] += sh[threadIdx.x + w]; shm[threadIdx.x] = fmax(shm[threadIdx.x], shm[threadIdx.x + w]); }
        __syncthreads();
    }
    // max-norm of the block: density-weighted screening (effective threshold of the build, scales_kernel)
    if (threadIdx.x == 0) { bnorm[(size_t)sb * nshell + sa] = sh[0]; bmax[(size_t)sb * nshell + sa] = shm[0]; }
}

// out_pure(block) = factor * C_a [ (acc + acc^T) / scale ] C_b^T ; integer sum first (exact), one conversion.
// The integer sum is taken modulo 2^64 (unsigned arithmetic): partial sums of the build may have wrapped around, the
// final value fits by construction of the scale (scales_kernel), so the wrapped sum IS the exact value.
// acc_lo != nullptr and scales[6] != 0: the J low limb (rounding residuals scaled by 2^31) is added.
// Only blocks sb <= sa (and m >= n inside diagonal blocks) are computed and mirrored, so the output is EXACTLY
// symmetric like the reference's 1/4 (raw + raw^T) (Int4C2E.cpp:661-664).
__global__ void finalize_kernel(int nbf, int ncart, const long long* __restrict__ acc, const long long* __restrict__ acc_lo,
                                const double* __restrict__ scales, int which_scale,
                                double factor, const double* __restrict__ ctrans, const int* __restrict__ ct_off,
                                const int* __restrict__ bf_off, const int* __restrict__ cao_off, const int* __restrict__ nfun,
                                const int* __restrict__ ncsh, double* __restrict__ out) {
    const int sa = blockIdx.x, sb = blockIdx.y;
    if (sb > sa) return;
    const int na = nfun[sa], nb = nfun[sb], nca = ncsh[sa], ncb = ncsh[sb];
    const double* Ca = ctrans + ct_off[sa];
    const double* Cb = ctrans + ct_off[sb];
    const double f = factor / scales[which_scale];
    const bool lo_on = acc_lo != nullptr && scales[6] != 0.0;
    for (int e = threadIdx.x; e < na * nb; e += blockDim.x) {
        const int m = e / nb, n = e % nb;
        if (sa == sb && n > m) continue;
        double s = 0.0;
        for (int x = 0; x < nca; x++) {
            const double cam = Ca[m * nca + x];
            if (cam == 0.0) continue;
            for (int y = 0; y < ncb; y++) {
                const size_t i = cao_off[sa] + x, j = cao_off[sb] + y;
                const long long v = (long long)((unsigned long long)acc[j * ncart + i] + (unsigned long long)acc[i * ncart + j]);
                double dv = (double)v;
                if (lo_on) {
                    const long long vl = (long long)((unsigned long long)acc_lo[j * ncart + i] + (unsigned long long)acc_lo[i * ncart + j]);
                    dv += (double)vl * 0x1p-31;
                }
                s = fma(cam * Cb[n * ncb + y], dv, s);
            }
        }
        s *= f;
        out[(size_t)(bf_off[sb] + n) * nbf + bf_off[sa] + m] = s;
        out[(size_t)(bf_off[sa] + m) * nbf + bf_off[sb] + n] = s;
    }
}

// Rigorous bounds on every partial sum the accumulators can hold, from the Cartesian Schwarz matrix
// QS[a,b] >= sqrt((ij|ij)) for all i in a, j in b (0 for dropped pairs) and the block 1-norms B[c,d] of |D_cart|:
//   block 0 (J):  sum_cd QS[c,d] B0[c,d]                    |rawJ_ij|  <= 8 QS_max * that
//   block x (K):  max_a sum_b QS[a,b] r_b, r_b = sum_d Bx[b,d]     |rawK_ik|  <= 8 QS_max * that
// All reductions run in a FIXED order (strided per thread, then a shared-memory tree): every rank derives bit-identical
// scales from the same density.
__global__ void bounds_kernel(int ns, const double* __restrict__ QS, const double* __restrict__ B, const double* __restrict__ Bmax,
                              double* __restrict__ r_work, double* __restrict__ bounds) {
    __shared__ double sh[1024];
    const int x = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* Bx = B + (size_t)x * ns * ns;
    double v = 0.0;
    if (x == 0) {
        for (size_t i = tid; i < (size_t)ns * ns; i += nt) v = fma(QS[i], Bx[i], v);
    } else {
        double* r = r_work + (size_t)x * ns;
        for (int b = tid; b < ns; b += nt) {          // B is symmetric: column sums read contiguously
            double sum = 0.0;
            for (int d = 0; d < ns; d++) sum += Bx[(size_t)b * ns + d];
            r[b] = sum;
        }
        __syncthreads();
        for (int a = tid; a < ns; a += nt) {
            double sum = 0.0;
            for (int b = 0; b < ns; b++) sum = fma(QS[(size_t)a * ns + b], r[b], sum);
            v = fmax(v, sum);
        }
    }
    sh[tid] = v;
    __syncthreads();
    for (int w = nt / 2; w > 0; w >>= 1) {
        if (tid < w) sh[tid] = (x == 0) ? sh[tid] + sh[tid + w] : fmax(sh[tid], sh[tid + w]);
        __syncthreads();
    }
    if (tid == 0) bounds[x] = sh[0];
    __syncthreads();
    // largest |D_cart| element of density x (0: total density): bounds[4 + x]
    const double* Mx = Bmax + (size_t)x * ns * ns;
    double m = 0.0;
    for (size_t i = tid; i < (size_t)ns * ns; i += nt) m = fmax(m, Mx[i]);
    sh[tid] = m;
    __syncthreads();
    for (int w = nt / 2; w > 0; w >>= 1) {
        if (tid < w) sh[tid] = fmax(sh[tid], sh[tid + w]);
        __syncthreads();
    }
    if (tid == 0) bounds[4 + x] = sh[0];
}
// scales[0] = J scale, scales[1] = K scale (powers of two), scales[2..3] = the bounds themselves,
// scales[4] = effective Schwarz threshold of this build: a quartet is evaluated iff Q_ab Q_cd > thr (the reference's
// test, Int4C2E.cpp:108-113) AND Q_ab Q_cd max|D| > dthr (density-weighted, off when dthr <= 0), scales[5] = max|D|,
// scales[6] = 1: the J adds feed the low limb as well, scales[7] = estimated worst rounding error of J without it.
//
// Range: the accumulators are summed modulo 2^64, so only the FINAL value of an element has to fit.  The value read by
// finalize_kernel is raw + raw^T = 4 J_cart (J) resp. 8 K_cart (K), and by Cauchy-Schwarz
//   |J_cart(i,j)| <= Q_ij sum_kl Q_kl |D_kl| <= Qmax * bounds[0],   |K_cart(i,k)| <= sum_jl Q_ij Q_kl |D_jl| <= Qmax * bounds[1+x]
// (every partial sum of any subset of the contributions obeys the same bound, so each double -> int64 conversion is in
// range as well).  A margin of 2^-10 covers the rounding of the bounds and of up to 2^52 individual conversions.
// Resolution: 1/scale per contribution; n roundings of +-1/2 unit pile up like sqrt(n/12).  For J on large systems
// (n ~ 10^6 shell pairs per element) that can approach the 1e-10 bar: the low limb is then switched on (jlo_mode 0:
// when the estimate exceeds 1e-11; 1: always; -1: never).
__global__ void scales_kernel(const double* __restrict__ bounds, int nk, double qmax, double thr, double dthr, double npair,
                              int jlo_mode, double* __restrict__ scales) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double nkmax = 0.0;
    for (int x = 0; x < nk; x++) nkmax = fmax(nkmax, bounds[1 + x]);
    const double margin = 1.0 + 0x1p-10;
    const double bj = 4.0 * qmax * bounds[0] * margin, bk = 8.0 * qmax * nkmax * margin;
    int ej, ek;
    frexp(fmax(bj, 1e-300), &ej);      // bj = m 2^ej, m in [1/2, 1)  ->  bj 2^(63-ej) < 2^63
    frexp(fmax(bk, 1e-300), &ek);
    // zero / tiny densities: cap the exponent so the scale stays finite (contributions are then exactly 0 anyway)
    scales[0] = ldexp(1.0, min(63 - ej, 512));
    scales[1] = ldexp(1.0, min(63 - ek, 512));
    scales[2] = bj;
    scales[3] = bk;
    double dmax = bounds[4];
    for (int x = 0; x < nk; x++) dmax = fmax(dmax, bounds[5 + x]);
    double eff = thr > 0.0 ? thr : 0.0;
    if (dthr > 0.0) eff = fmax(eff, dmax > 0.0 ? dthr / dmax : 1e300);
    scales[4] = eff;
    scales[5] = dmax;
    const double est = 0.25 / scales[0] * sqrt(npair / 6.0) * 6.0;     // J = raw / 4; ~6 sigma of sqrt(n/12)-type noise, two orientations
    scales[6] = (jlo_mode > 0 || (jlo_mode == 0 && est > 1e-11)) ? 1.0 : 0.0;
    scales[7] = est;
}

// Schwarz bounds of one pair class from the stored Cartesian (ab|ab) blocks: one CTA per pair
__global__ void schwarz_kernel(int npair, int nca, int ncb, const double* __restrict__ blocks, const int* __restrict__ sa,
                               const int* __restrict__ sb, const double* __restrict__ ctrans, const int* __restrict__ ct_off,
                               const int* __restrict__ bf_off, const int* __restrict__ nfun, int nbf, double* __restrict__ Qcart,
                               double* __restrict__ Qpure, double* __restrict__ diag) {
    const int ip = blockIdx.x;
    const int nab = nca * ncb;
    const double* V = blocks + (size_t)ip * nab * nab;
    const int a = sa[ip], b = sb[ip];
    const int na = nfun[a], nb = nfun[b];
    const double* Ca = ctrans + ct_off[a];
    const double* Cb = ctrans + ct_off[b];
    __shared__ double red[2][128];
    double qc = 0.0, qp = 0.0;
    for (int x = threadIdx.x; x < nab; x += blockDim.x) qc = fmax(qc, sqrt(fabs(V[(size_t)x * nab + x])));
    for (int m = threadIdx.x; m < na * nb; m += blockDim.x) {
        const int ma = m / nb, mb = m % nb;
        double s = 0.0;
        for (int x = 0; x < nab; x++) {
            const double mx = Ca[ma * nca + x / ncb] * Cb[mb * ncb + x % ncb];
            if (mx == 0.0) continue;
            double r = 0.0;
            for (int y = 0; y < nab; y++) r = fma(Ca[ma * nca + y / ncb] * Cb[mb * ncb + y % ncb], V[(size_t)x * nab + y], r);
            s = fma(mx, r, s);
        }
        qp = fmax(qp, sqrt(fabs(s)));
        if (diag) {
            diag[(size_t)(bf_off[b] + mb) * nbf + bf_off[a] + ma] = s;
            diag[(size_t)(bf_off[a] + ma) * nbf + bf_off[b] + mb] = s;
        }
    }
    red[0][threadIdx.x] = qc; red[1][threadIdx.x] = qp;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            red[0][threadIdx.x] = fmax(red[0][threadIdx.x], red[0][threadIdx.x + w]);
            red[1][threadIdx.x] = fmax(red[1][threadIdx.x], red[1][threadIdx.x + w]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { Qcart[ip] = red[0][0]; Qpure[ip] = red[1][0]; }
}

// Work-item lists (SURVEY 8a a4-a6 replaced): one thread per CTA-uniform pair ih walks the primitive-count segments
// of the spread class; inside a segment Q descends, so the pairs that can pass Q_H * Q_S > thr are a prefix found by
// bisection.  Prefixes of adjacent segments are merged into contiguous runs (no screening: ONE run), and the runs are
// cut into pieces of <= nq pairs = work items.  counts-only pass when items == nullptr.
__global__ void item_list_kernel(int nH, const double* __restrict__ QH, const double* __restrict__ QS, const int* __restrict__ seg,
                                 int nseg, int nq, double thr, int same, long long* __restrict__ counts,
                                 const long long* __restrict__ offs, int4* __restrict__ items) {
    const int ih = blockIdx.x * blockDim.x + threadIdx.x;
    if (ih >= nH) return;
    const double need = thr > 0.0 ? thr / fmax(QH[ih], 1e-300) : -1.0;
    long long n = 0;
    int4* out = items ? items + offs[ih] : nullptr;
    int run_lo = 0, run_hi = 0;                  // surviving prefixes of adjacent segments are merged into one run
    auto flush = [&]() {
        for (int k = run_lo; k < run_hi; k += nq) {
            if (out) out[n] = make_int4(ih, k, min(nq, run_hi - k), 0);
            n++;
        }
    };
    for (int s = 0; s < nseg; s++) {
        const int lo = seg[s];
        int hi = seg[s + 1];
        if (same) hi = min(hi, ih + 1);          // canonical quartets of a triangular task: spread pair <= uniform pair
        if (hi <= lo) continue;
        int len = hi - lo;
        if (thr > 0.0) {                         // largest prefix with Q > need
            int a = 0, b = len;
            while (a < b) {
                const int mid = (a + b) >> 1;
                if (QS[lo + mid] > need) a = mid + 1; else b = mid;
            }
            len = a;
        }
        if (len == 0) continue;
        if (run_hi == lo) run_hi = lo + len;
        else { flush(); run_lo = lo; run_hi = lo + len; }
    }
    flush();
    if (counts) counts[ih] = n;
}

// Work items of the bra-loop kernels (eri_tpqa.cuh): one thread per bra chunk walks the aligned ket blocks of 32.
// An item survives iff the best quartet of (chunk, block) can pass the Schwarz test and, for triangular tasks, the
// chunk reaches up to the block (canonical iff ket position <= bra position).  counts-only pass when items == nullptr.
__global__ void item_list_a_kernel(int nchunk, const int* __restrict__ cstart, const int* __restrict__ ccnt,
                                   const int* __restrict__ cmaxpos, const double* __restrict__ cq, int nblk, int npairK,
                                   const double* __restrict__ blkQ, double thr, int same, long long* __restrict__ counts,
                                   const long long* __restrict__ offs, int4* __restrict__ items) {
    const int ic = blockIdx.x * blockDim.x + threadIdx.x;
    if (ic >= nchunk) return;
    const double q = cq[ic];
    const int start = cstart[ic], cnt = ccnt[ic], maxpos = cmaxpos[ic];
    int4* out = items ? items + offs[ic] : nullptr;
    long long n = 0;
    for (int r = 0; r < nblk; r++) {
        const int k0 = r * 32;
        if (same && maxpos < k0) break;
        if (thr > 0.0 && !(q * blkQ[r] > thr)) continue;
        if (out) out[n] = make_int4(start, k0, min(32, npairK - k0), cnt);
        n++;
    }
    if (counts) counts[ic] = n;
}

// The reference's screening counts (getRepulsionLength, Int4C2E.cpp:79-128) for threshold > 0, part s3 == s1 of its loop
// nest s1; s2 <= s1; s3 <= s1; s4 <= max(s2, s3): the two shell pairs share shell s1, so the uniqueness predicate
// bf2 <= bf1 && bf3 <= bf1 && bf4 <= (bf1 == bf3 ? bf2 : bf3) couples them and the test has to be made function by
// function on Diag1212.  One CTA per s1, threads over (s2, s4); P = shell-pair maxima of sqrt|Diag| prune most pairs.
__global__ void ref_count_special_kernel(int ns, int nbf, const int* __restrict__ bf_off, const int* __restrict__ nfun,
                                         const double* __restrict__ diag, const double* __restrict__ P, double thr,
                                         unsigned long long* __restrict__ out) {
    const int s1 = blockIdx.x;
    const int n1 = nfun[s1], o1 = bf_off[s1];
    unsigned long long nq = 0, ni = 0;
    const long long npairs = (long long)(s1 + 1) * (s1 + 1);
    for (long long e = threadIdx.x; e < npairs; e += blockDim.x) {
        const int s2 = (int)(e / (s1 + 1)), s4 = (int)(e % (s1 + 1));
        if (!(P[(size_t)s1 * ns + s2] * P[(size_t)s1 * ns + s4] > thr)) continue;
        const int n2 = nfun[s2], o2 = bf_off[s2], n4 = nfun[s4], o4 = bf_off[s4];
        bool keep = false;
        for (int f1 = 0; f1 < n1 && !keep; f1++) {
            const int bf1 = o1 + f1;
            for (int f2 = 0; f2 < n2 && !keep; f2++) {
                const int bf2 = o2 + f2;
                if (bf2 > bf1) break;
                const double d12 = diag[(size_t)bf2 * nbf + bf1];
                for (int f3 = 0; f3 <= f1 && !keep; f3++) {
                    const int bf3 = o1 + f3;
                    const int lim = (bf1 == bf3) ? bf2 : bf3;
                    for (int f4 = 0; f4 < n4; f4++) {
                        const int bf4 = o4 + f4;
                        if (bf4 > lim) break;
                        if (sqrt(fabs(d12 * diag[(size_t)bf4 * nbf + bf3])) > thr) { keep = true; break; }
                    }
                }
            }
        }
        int uniq = 0;
        if (keep) {   // number of unique function quartets: a count over (f1, f3 <= f1) with closed forms for f2 and f4
            for (int f1 = 0; f1 < n1; f1++) {
                const int c2 = s2 < s1 ? n2 : f1 + 1;                       // f2 with bf2 <= bf1
                for (int f3 = 0; f3 < f1; f3++) uniq += c2 * (s4 < s1 ? n4 : f3 + 1);      // bf4 <= bf3
                // f3 == f1: bf4 <= bf2
                if (s4 < s2) uniq += c2 * n4;
                else if (s4 == s2) uniq += c2 * (c2 + 1) / 2;               // sum over f2 < c2 of (f2 + 1)
            }
        }
        if (keep) { nq++; ni += (unsigned long long)uniq; }
    }
    __shared__ unsigned long long sh[2][256];
    sh[0][threadIdx.x] = nq; sh[1][threadIdx.x] = ni;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sh[0][threadIdx.x] += sh[0][threadIdx.x + w]; sh[1][threadIdx.x] += sh[1][threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(out, sh[0][0]); atomicAdd(out + 1, sh[1][0]); }
}

// register-resident DFMA loop: the FP64 roofline denominator measured on the device itself
__global__ void dfma_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
static double fact_d(int n) { double r = 1; for (int i = 2; i <= n; i++) r *= i; return r; }
static double binom_d(int n, int k) { return (k < 0 || k > n) ? 0.0 : fact_d(n) / (fact_d(k) * fact_d(n - k)); }
static int cart_index_h(int l, int lx, int ly) {
    int n = 0;
    for (int x = l; x > lx; x--) n += l - x + 1;
    return n + (l - lx - ly);
}
// rows: functions of a shell of `type` in the reference's order; cols: Cartesian monomials (lx desc, ly desc)
static std::vector<double> shell_transform_h(int type) {
    const int l = std::abs(type), nc = cf_ncart(l);
    if (type >= 0) {
        std::vector<double> C((size_t)nc * nc, 0.0);
        for (int i = 0; i < nc; i++) C[(size_t)i * nc + i] = 1.0;
        return C;
    }
    std::vector<double> C((size_t)(2 * l + 1) * nc, 0.0);
    for (int m = -l; m <= l; m++) {
        const int am = std::abs(m);
        const double N = std::sqrt(2.0 * fact_d(l + am) * fact_d(l - am) / (m == 0 ? 2.0 : 1.0)) / (std::pow(2.0, am) * fact_d(l));
        for (int t = 0; t <= (l - am) / 2; t++)
            for (int u = 0; u <= t; u++)
                for (int v2 = (m >= 0 ? 0 : 1); v2 <= am; v2 += 2) {
                    const int k = (m >= 0) ? v2 / 2 : (v2 - 1) / 2;
                    const double c = (((t + k) & 1) ? -1.0 : 1.0) * std::pow(0.25, t) * binom_d(l, t) * binom_d(l - t, am + t) *
                                     binom_d(t, u) * binom_d(am, v2);
                    const int ex = 2 * t + am - 2 * u - v2, ey = 2 * u + v2;
                    if (ex < 0) continue;
                    C[(size_t)(m + l) * nc + cart_index_h(l, ex, ey)] += N * c;
                }
    }
    return C;
}

static int check_device(cf_handle* h, int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { set_error(h, "no CUDA device (this engine has no CPU fallback)"); return CF_ERR_NO_DEVICE; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { set_error(h, "device ordinal out of range"); return CF_ERR_NO_DEVICE; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { set_error(h, "cudaGetDeviceProperties failed"); return CF_ERR_NO_DEVICE; }
    if (p.major != 10) { set_error(h, "device is not sm_100 (kernels are built for sm_100a only)"); return CF_ERR_NO_DEVICE; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error(h, "cudaSetDevice failed"); return CF_ERR_NO_DEVICE; }
    {   // the default memory pool keeps what setup frees (DevBuf): later allocations are served from it
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) { unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
        (void)cudaGetLastError();
    }
    if (h) h->device = device;
    return CF_OK;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int launch_task(cf_handle* h, ClassPairTask* t, QuartetTask& qt, int store, cudaStream_t s) {
    const PairClassHost& B = h->cls[t->bra];
    const PairClassHost& K = h->cls[t->ket];
    qt.bra = B.dev(); qt.ket = K.dev();
    if (t->swap && !store) { qt.bra = K.dev(); qt.ket = B.dev(); }
    qt.qoff = t->d_qoff.p; qt.nquartet = t->nquartet;
    qt.items = t->d_items.p; qt.nitem = t->nitem;
    qt.braloop = t->braloop;
    qt.border = B.d_border.p; qt.brec_i = B.d_brec_i.p; qt.brec_d = B.d_brec_d.p; qt.bprim = B.d_bprim.p; qt.bprim_stride = B.bprim_stride;
    qt.same_class = (t->bra == t->ket);
    qt.thr = h->opt.threshold > 0 ? h->opt.threshold : 0.0;
    if (!store && t->owner >= 0) {            // small task owned by one rank
        if (qt.rank != t->owner) return CF_OK;
        qt.rank = 0; qt.world = 1;
    }
    const long long nq = qt.nquartet;
    if (nq == 0) return CF_OK;
    int grid;
    if (t->kind >= 1 && !store) {
        long long nlocal = (t->nitem - qt.rank + qt.world - 1) / qt.world;
        if (nlocal <= 0) return CF_OK;
        if (t->kind == 1) nlocal = (nlocal + TPQ_THREADS / 32 - 1) / (TPQ_THREADS / 32);   // warp-private items
        grid = (int)std::min<long long>(nlocal, 148LL * 16);
        // Every CTA stages its root tables (18-64 KB) before the first item: when a partition leaves only a few items
        // per CTA (many GPUs, small class pairs) that fixed cost and the ragged tail dominate (round 1: 0.83 efficiency
        // at 8 GPUs on c18).  Fewer, longer-running CTAs then: at least CF_MIN_ITEMS items per CTA as long as every SM
        // keeps 4 CTAs.  Launch geometry only -- the set of fixed-point adds, hence the result, is unchanged.
        {
            static int min_items = -1;
            if (min_items < 0) { const char* e = getenv("CF_MIN_ITEMS"); min_items = e ? atoi(e) : 8; }
            const long long want = std::max<long long>(148LL * 4, (nlocal + min_items - 1) / std::max(1, min_items));
            grid = (int)std::min<long long>(grid, std::max<long long>(1, std::min<long long>(nlocal, want)));
        }
    } else {
        // chunk: enough chunks to fill the machine ~8x over, at most 64 quartets each
        long long chunk = nq / (148LL * 32 * std::max(1, qt.world));
        chunk = std::max(1LL, std::min(64LL, chunk));
        qt.chunk = (int)chunk;
        const long long nchunk_total = (nq + chunk - 1) / chunk;
        const long long nchunk_local = (nchunk_total - qt.rank + qt.world - 1) / qt.world;
        if (nchunk_local <= 0) return CF_OK;
        grid = (int)std::min<long long>(nchunk_local, 148LL * 32);
    }
    cudaError_t e = g_bra_launch[t->bra](t->ket, qt, store, grid, s, nullptr, nullptr, nullptr);
    if (e != cudaSuccess) { set_error(h, std::string("ERI kernel launch failed: ") + cudaGetErrorString(e)); return CF_ERR_CUDA; }
    h->stats.n_launches_last++;
    return CF_OK;
}

// F_m(T) for m = 0..mmax by the (all-positive) series e^-T sum_k (2T)^k / ((2m+1)(2m+3)...(2m+2k+1)), long double
static void boys_reference(double T, int mmax, double* out) {
    const long double t = T, e = expl(-t);
    for (int m = 0; m <= mmax; m++) {
        long double term = 1.0L / (2 * m + 1), sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2.0L * t / (2 * m + 2 * k + 1);
            sum += term;
            if (term < 1e-22L * sum) break;
        }
        out[m] = (double)(e * sum);
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" const char* cf_last_error(const cf_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

extern "C" int cf_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { set_error(nullptr, "no CUDA device"); return CF_ERR_NO_DEVICE; }
    if (device < 0) cudaGetDevice(&device);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return CF_ERR_CUDA;
    if (name && name_len > 0) { std::strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return CF_OK;
}

extern "C" int cf_measure_fp64_peak(int device, double* tflops) {
    cf_handle* h = nullptr;
    int rc = check_device(nullptr, device);
    if (rc != CF_OK) return rc;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device < 0 ? 0 : device);
    const int threads = 512, blocks = sms * 4, iters = 1 << 15;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * iters * (double)threads * blocks;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (tflops) *tflops = best;
    return CF_OK;
}

extern "C" void cf_destroy(cf_handle* h) {
    if (!h) return;
    if (h->multi) { multi_destroy(h); delete h; return; }
    DeviceGuard guard(h->device);
    for (auto& c : h->cls) c.release();
    for (auto* t : h->tasks) { t->d_qoff.release(); t->d_items.release(); delete t; }
    h->d_ctrans.release(); h->d_ct_off.release(); h->d_bf_off.release(); h->d_cao_off.release(); h->d_nfun.release(); h->d_ncartsh.release();
    h->d_rys_table.release(); h->d_rys_asym.release(); h->d_boys.release();
    for (auto& b : h->d_Dpure) b.release();
    for (auto& b : h->d_Dcart) b.release();
    for (auto& b : h->d_out) b.release();
    h->d_partial.release(); h->d_diag.release(); h->d_acc.release(); h->d_cnt.release();
    h->d_QS.release(); h->d_B.release(); h->d_Bmax.release(); h->d_rwork.release();
    h->d_shell2atom.release(); h->d_gpart.release(); h->d_grad.release(); h->d_hess.release();
    h->d_atomZ.release(); h->d_atomxyz.release();
    h->d_l.release(); h->d_nprim_sh.release(); h->d_prim_off_sh.release(); h->d_exps.release(); h->d_coefs.release(); h->d_xyz.release();
    for (auto& e : h->ev) cudaEventDestroy(e);
    for (int i = 0; i < 3; i++) { cudaStreamDestroy(h->side[i]); cudaEventDestroy(h->ev_join[i]); }
    cudaEventDestroy(h->ev_fork);
    delete h;
}

static void fill_rys_tables(RysTablesDev& r, const cf_handle* h) {
    r.table = h->d_rys_table.p;
    r.asym = h->d_rys_asym.p;
    r.boys = h->d_boys.p;
}
static void fill_rys(QuartetTask& qt, const cf_handle* h) { fill_rys_tables(qt.rys, h); }

extern "C" cf_handle* cf_create(const cf_basis* basis, const cf_options* opts) {
    if (!basis || basis->nshell <= 0 || !basis->type || !basis->nprim || !basis->prim_offset || !basis->exps ||
        !basis->coefs_normalized || !basis->center_xyz) {
        set_error(nullptr, "cf_create: null or empty basis");
        return nullptr;
    }
    struct RestoreDevice { int prev = -1; RestoreDevice() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; } ~RestoreDevice() { if (prev >= 0) cudaSetDevice(prev); } } restore_device;
    cf_handle* h = new cf_handle();
    if (opts) h->opt = *opts;
    if (h->opt.world_size <= 0) { h->opt.world_size = 1; h->opt.rank = 0; }
    if (h->opt.rank < 0 || h->opt.rank >= h->opt.world_size) { set_error(nullptr, "cf_create: rank out of range"); delete h; return nullptr; }
    if (h->opt.pair_cutoff <= 0) h->opt.pair_cutoff = 1e-20;
    if (check_device(h, h->opt.device) != CF_OK) { g_last_error = h->err; delete h; return nullptr; }
    const double t_start = now_s();
    const bool timing = h->opt.verbose >= 2 || getenv("CF_SETUP_TIMING") != nullptr;
    double t_mark = t_start;
    auto mark = [&](const char* what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        const double t = now_s();
        std::fprintf(stderr, "  [cf_create] %-28s %8.3f ms\n", what, (t - t_mark) * 1e3);
        t_mark = t;
    };
    auto fail = [&](const std::string& s) -> cf_handle* { set_error(nullptr, s); cf_destroy(h); return nullptr; };
    for (auto& e : h->ev) cudaEventCreate(&e);
    for (int i = 0; i < 3; i++) { cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking); cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming); }
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);

    // static check of the table geometry assumed by the kernels
    for (int n = 1; n <= RYS_NMAX; n++)
        if (rys_tmax(n) != rys_tmax_h[n] || rys_off(n) != rys_off_h[n] || rys_asym_off(n) != rys_asym_off_h[n])
            return fail("Rys table geometry mismatch between generator and kernels");

    // ---- shells
    const int ns = basis->nshell;
    h->nshell = ns;
    h->type.assign(basis->type, basis->type + ns);
    h->nprim.assign(basis->nprim, basis->nprim + ns);
    h->prim_off.assign(basis->prim_offset, basis->prim_offset + ns);
    h->xyz.assign(basis->center_xyz, basis->center_xyz + 3 * ns);
    if (basis->shell2atom) h->shell2atom.assign(basis->shell2atom, basis->shell2atom + ns);
    int nptot = 0;
    for (int s = 0; s < ns; s++) nptot = std::max(nptot, h->prim_off[s] + h->nprim[s]);
    h->exps.assign(basis->exps, basis->exps + nptot);
    h->coefs.assign(basis->coefs_normalized, basis->coefs_normalized + nptot);
    h->l.resize(ns); h->bf_off.resize(ns); h->cao_off.resize(ns); h->nfun.resize(ns); h->ct_off.resize(ns);
    std::vector<int> ncsh(ns);
    std::vector<int> pool_off(32, -1);   // by type + 16
    int nbf = 0, ncart = 0;
    for (int s = 0; s < ns; s++) {
        const int t = h->type[s], l = std::abs(t);
        if (l > CF_LMAX_DEV) return fail("cf_create: shell angular momentum above CF_MAX_L is not supported by the device kernels");
        if (t > 1) return fail("cf_create: Cartesian shells with l >= 2 are not produced by the reference's basis reader and are not supported");
        if (h->nprim[s] <= 0) return fail("cf_create: shell without primitives");
        h->l[s] = l;
        h->nfun[s] = t < 0 ? 2 * l + 1 : cf_ncart(l);
        ncsh[s] = cf_ncart(l);
        h->bf_off[s] = nbf; h->cao_off[s] = ncart;
        nbf += h->nfun[s]; ncart += ncsh[s];
        if (pool_off[t + 16] < 0) {
            pool_off[t + 16] = (int)h->ctrans.size();
            auto C = shell_transform_h(t);
            h->ctrans.insert(h->ctrans.end(), C.begin(), C.end());
        }
        h->ct_off[s] = pool_off[t + 16];
    }
    if (nbf >= 32768) return fail("cf_create: nbf >= 32768 (the reference's short int indices, Int4C2E.h:19-29)");
    h->nbf = nbf; h->ncart = ncart;

#define UP(buf, vec) if ((buf).upload(vec) != cudaSuccess) return fail("cudaMalloc/cudaMemcpy failed in cf_create")
    UP(h->d_ctrans, h->ctrans); UP(h->d_ct_off, h->ct_off); UP(h->d_bf_off, h->bf_off); UP(h->d_cao_off, h->cao_off);
    UP(h->d_nfun, h->nfun); UP(h->d_ncartsh, ncsh);
    {
        std::vector<double> tab(rys_table_h, rys_table_h + RYS_TABLE_LEN), asym(rys_asym_h, rys_asym_h + RYS_NMAX * (RYS_NMAX + 1));
        UP(h->d_rys_table, tab); UP(h->d_rys_asym, asym);
        std::vector<double> boys((size_t)BOYS_NROW * BOYS_NCOL);
        for (int i = 0; i < BOYS_NROW; i++) boys_reference(i * BOYS_DT, BOYS_NCOL - 1, &boys[(size_t)i * BOYS_NCOL]);
        UP(h->d_boys, boys);
    }

    mark("shells + tables");
    // ---- shell pairs: built, primitive-screened, Schwarz-bounded and sorted ON THE DEVICE (north_star; replaces the
    // reference's serial pair loops, Int4C2E.cpp:19-231).  The host only buckets pair ids by class and lays out slot bases.
    for (int la = 0; la <= CF_LMAX_DEV; la++)
        for (int lb = 0; lb <= la; lb++) { auto& c = h->cls[cf_pair_class(la, lb)]; c.la = la; c.lb = lb; }
    UP(h->d_l, h->l); UP(h->d_nprim_sh, h->nprim); UP(h->d_prim_off_sh, h->prim_off); UP(h->d_exps, h->exps); UP(h->d_coefs, h->coefs); UP(h->d_xyz, h->xyz);
    BasisDev bd;
    bd.l = h->d_l.p; bd.nprim = h->d_nprim_sh.p; bd.prim_off = h->d_prim_off_sh.p; bd.cao_off = h->d_cao_off.p;
    bd.exps = h->d_exps.p; bd.coefs = h->d_coefs.p; bd.xyz = h->d_xyz.p;
    const double cutoff = h->opt.pair_cutoff;
    long long pairs_kept = 0;
    {
        const long long npairs_tot = (long long)ns * (ns + 1) / 2;
        DevBuf<unsigned char> d_cls; DevBuf<int> d_kept;
        if (d_cls.alloc(npairs_tot) != cudaSuccess || d_kept.alloc(npairs_tot) != cudaSuccess) return fail("cudaMalloc failed (pair classification)");
        pair_classify_kernel<<<(unsigned)((npairs_tot + 127) / 128), 128>>>(ns, npairs_tot, bd, cutoff, d_cls.p, d_kept.p);
        std::vector<unsigned char> cls_h(npairs_tot); std::vector<int> kept_h(npairs_tot);
        if (cudaMemcpy(cls_h.data(), d_cls.p, npairs_tot, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(kept_h.data(), d_kept.p, sizeof(int) * npairs_tot, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(std::string("pair classification kernel failed: ") + cudaGetErrorString(cudaGetLastError()));
        d_cls.release(); d_kept.release();
        long long e = 0;
        for (int s1 = 0; s1 < ns; s1++)
            for (int s2 = 0; s2 <= s1; s2++, e++) {
                if (cls_h[e] == 255) continue;
                int a = s1, b = s2;
                if (h->l[b] > h->l[a]) std::swap(a, b);
                PairClassHost& c = h->cls[cls_h[e]];
                c.sa.push_back(a); c.sb.push_back(b);
                c.nprim.push_back(kept_h[e]); c.nprim_full.push_back(h->nprim[a] * h->nprim[b]);
                c.Q.push_back(0.0); c.Qpure.push_back(0.0);
                pairs_kept++;
            }
    }
    mark("pair classify + bucket");
    for (auto& c : h->cls) if (c.layout_and_fill(bd, cutoff) != cudaSuccess) return fail("pair layout failed");
    mark("pair fill (device)");

    // ---- Schwarz bounds: (ab|ab) Cartesian blocks from the quartet kernel in STORE/diag mode; then the class is sorted
    // by (primitive count desc, pure-function Schwarz bound desc) with a device radix sort and laid out in that order.
    // Equal primitive counts inside a warp keep the per-thread primitive loops of the thread-per-quartet kernels
    // convergent; heavy pairs first gives the static schedule a short tail.
    if (h->d_diag.alloc((size_t)nbf * nbf) != cudaSuccess) return fail("cudaMalloc failed (diag)");
    cudaMemset(h->d_diag.p, 0, sizeof(double) * (size_t)nbf * nbf);
    h->qmax_cart = 0;
    DevBuf<double> blocks;          // (ab|ab) store buffer, one allocation shared by all classes (slabs of at most ~1 GiB)
    blocks.pooled = false;
    {
        size_t need = 0;
        for (int ci = 0; ci < CF_NCLS; ci++) {
            const PairClassHost& c = h->cls[ci];
            if (c.npair() == 0) continue;
            const size_t nout = (size_t)cf_ncart(c.la) * cf_ncart(c.lb) * cf_ncart(c.la) * cf_ncart(c.lb);
            need = std::max(need, nout * std::max<size_t>(1, std::min<size_t>(c.npair(), (1ull << 27) / nout)));
        }
        if (blocks.alloc(need) != cudaSuccess) return fail("cudaMalloc failed (Schwarz store buffer)");
    }
    for (int ci = 0; ci < CF_NCLS; ci++) {
        PairClassHost& c = h->cls[ci];
        const int np = c.npair();
        if (np == 0) continue;
        const int nca = cf_ncart(c.la), ncb = cf_ncart(c.lb);
        const size_t nout = (size_t)nca * ncb * nca * ncb;
        // in slabs so the store buffer stays below ~1 GiB
        const int slab = (int)std::max<size_t>(1, std::min<size_t>(np, (1ull << 27) / nout));
        DevBuf<double> qpure;
        if (qpure.alloc(np) != cudaSuccess || c.d_Qcart.alloc(np) != cudaSuccess) return fail("cudaMalloc failed (Schwarz)");
        for (int p0 = 0; p0 < np; p0 += slab) {
            const int cnt = std::min(slab, np - p0);
            QuartetTask qt{};
            PairClassDev d = c.dev();
            // shift the pair arrays so that local pair q maps to global pair p0+q
            d.npair = cnt; d.sa += p0; d.sb += p0; d.cao_a += p0; d.cao_b += p0; d.pbase += p0; d.nprim += p0; d.A += 3 * p0; d.AB += 3 * p0; d.Q += p0;
            qt.bra = d; qt.ket = d; qt.qoff = nullptr; qt.nquartet = cnt; qt.chunk = 1; qt.rank = 0; qt.world = 1;
            qt.same_class = 1; qt.ncart = ncart; qt.nk = 0; qt.store = blocks.p; qt.diag = 1; qt.prim_cut = 0.0; qt.thr = 0.0;
            fill_rys(qt, h);
            cudaError_t e = g_bra_launch[ci](ci, qt, 1, cnt, 0, nullptr, nullptr, nullptr);
            if (e != cudaSuccess) return fail(std::string("Schwarz launch failed: ") + cudaGetErrorString(e));
            schwarz_kernel<<<cnt, 128>>>(cnt, nca, ncb, blocks.p, c.d_sa.p + p0, c.d_sb.p + p0, h->d_ctrans.p, h->d_ct_off.p, h->d_bf_off.p,
                                         h->d_nfun.p, nbf, c.d_Qcart.p + p0, qpure.p + p0, h->d_diag.p);
        }
        if (c.sort_by_bounds(bd, cutoff, qpure.p) != cudaSuccess)
            return fail(std::string("Schwarz kernels / class sort failed: ") + cudaGetErrorString(cudaGetLastError()));
        qpure.release();
        for (double q : c.Q) h->qmax_cart = std::max(h->qmax_cart, q);
        if (c.build_roles() != cudaSuccess) return fail("bra/ket role build failed");
    }
    blocks.release();
    h->diag_ready = true;
    mark("Schwarz + device sort + roles");
    {   // dense Cartesian Schwarz matrix over shell pairs (0 for pairs dropped by the primitive cutoff)
        std::vector<double> QS((size_t)ns * ns, 0.0);
        for (auto& c : h->cls)
            for (int i = 0; i < c.npair(); i++) {
                QS[(size_t)c.sa[i] * ns + c.sb[i]] = c.Q[i];
                QS[(size_t)c.sb[i] * ns + c.sa[i]] = c.Q[i];
            }
        if (h->d_QS.upload(QS) != cudaSuccess || h->d_B.alloc(4 * (size_t)ns * ns) != cudaSuccess || h->d_Bmax.alloc(4 * (size_t)ns * ns) != cudaSuccess || h->d_rwork.alloc(4 * (size_t)ns) != cudaSuccess)
            return fail("cudaMalloc failed (Schwarz matrix)");
    }
    {   // ---- the reference's own counts RepulsionLength / ShellQuartetLength (getRepulsionLength, Int4C2E.cpp:79-128):
        // loop nest s1; s2 <= s1; s3 <= s1; s4 <= max(s2,s3), a shell quartet counts iff one of its UNIQUE function
        // quartets has sqrt|Diag(bf1,bf2) Diag(bf3,bf4)| > threshold, and then contributes all its unique function quartets.
        const double thr0 = h->opt.threshold;
        long long nq_ref = 0, ni_ref = 0;
        if (!(thr0 > 0.0)) {   // nothing is screened (the reference's default -1): closed forms
            const long long N = (long long)nbf * (nbf + 1) / 2;
            ni_ref = N * (N + 1) / 2;
            for (long long s1 = 0; s1 < ns; s1++)   // s3 < s1: every s4 <= s3 holds unique functions; s3 == s1: s4 <= s1 (s4 <= s2 for one-function shells)
                nq_ref += (s1 + 1) * (s1 + 1) * s1 / 2 + (h->nfun[s1] >= 2 ? (s1 + 1) * (s1 + 1) : (s1 + 1) * (s1 + 2) / 2);
        } else {
            // P[a,b] = max over the shell pair's functions of sqrt|Diag1212| (0 for pairs without surviving primitives)
            std::vector<double> P((size_t)ns * ns, 0.0);
            for (auto& c : h->cls)
                for (int i = 0; i < c.npair(); i++) { P[(size_t)c.sa[i] * ns + c.sb[i]] = c.Qpure[i]; P[(size_t)c.sb[i] * ns + c.sa[i]] = c.Qpure[i]; }
            // part s3 < s1 (then only s4 <= s3 holds unique functions, and every function quartet with bf2 <= bf1, bf4 <= bf3
            // is unique): kept iff P12 P34 > thr.  Pairs (s3,s4) enter a Fenwick tree over their rank in P as s1 advances.
            const long long npr = (long long)ns * (ns + 1) / 2;
            std::vector<int> rank_of(npr);
            {
                std::vector<long long> idx(npr);
                std::iota(idx.begin(), idx.end(), 0LL);
                auto pv = [&](long long e) { long long a = (long long)((std::sqrt(8.0 * e + 1.0) - 1.0) / 2.0); while (a * (a + 1) / 2 > e) a--; while ((a + 1) * (a + 2) / 2 <= e) a++; return P[(size_t)a * ns + (e - a * (a + 1) / 2)]; };
                std::vector<double> pvals(npr);
                for (long long e = 0; e < npr; e++) pvals[e] = pv(e);
                std::stable_sort(idx.begin(), idx.end(), [&](long long x, long long y) { return pvals[x] > pvals[y]; });
                std::vector<double> sorted(npr);
                for (long long r = 0; r < npr; r++) { rank_of[idx[r]] = (int)r; sorted[r] = pvals[idx[r]]; }
                std::vector<long long> fen_n(npr + 1, 0), fen_u(npr + 1, 0);
                auto add = [&](int r, long long u) { for (int i = r + 1; i <= npr; i += i & -i) { fen_n[i]++; fen_u[i] += u; } };
                auto query = [&](int cnt, long long& n, long long& u) { n = 0; u = 0; for (int i = cnt; i > 0; i -= i & -i) { n += fen_n[i]; u += fen_u[i]; } };
                auto upair = [&](int a, int b) { return a == b ? (long long)h->nfun[a] * (h->nfun[a] + 1) / 2 : (long long)h->nfun[a] * h->nfun[b]; };
                for (int s1 = 0; s1 < ns; s1++) {
                    if (s1 > 0) for (int s4 = 0; s4 <= s1 - 1; s4++) add(rank_of[(long long)(s1 - 1) * s1 / 2 + s4], upair(s1 - 1, s4));   // pairs with s3 = s1 - 1
                    for (int s2 = 0; s2 <= s1; s2++) {
                        const double p12 = P[(size_t)s1 * ns + s2];
                        if (!(p12 > 0.0)) continue;
                        // number of ranks r with sorted[r] * p12 > thr (sorted descends)
                        const int cnt = (int)(std::partition_point(sorted.begin(), sorted.end(), [&](double q) { return q * p12 > thr0; }) - sorted.begin());
                        long long n, u;
                        query(cnt, n, u);
                        nq_ref += n; ni_ref += u * upair(s1, s2);
                    }
                }
            }
            // part s3 == s1 on the device
            DevBuf<double> dP; DevBuf<unsigned long long> dout;
            if (dP.upload(P) != cudaSuccess || dout.alloc(2) != cudaSuccess) return fail("cudaMalloc failed (reference counts)");
            cudaMemset(dout.p, 0, 2 * sizeof(unsigned long long));
            ref_count_special_kernel<<<ns, 256>>>(ns, nbf, h->d_bf_off.p, h->d_nfun.p, h->d_diag.p, dP.p, thr0, dout.p);
            unsigned long long o2[2] = {0, 0};
            if (cudaMemcpy(o2, dout.p, sizeof(o2), cudaMemcpyDeviceToHost) != cudaSuccess) return fail("reference count kernel failed");
            nq_ref += (long long)o2[0]; ni_ref += (long long)o2[1];
            dP.release(); dout.release();
        }
        h->ref_counts[0] = ni_ref; h->ref_counts[1] = nq_ref;
    }
    mark("QS matrix + reference counts");

    // ---- class-pair tasks.  Quartet (ib, ik) of a task is canonical iff ik <= ib when bra class == ket class.
    // Schwarz screening (threshold > 0, Int4C2E.cpp:108-113) is a per-quartet test Q_b * Q_k > thr inside the kernels;
    // the statistics below count exactly the quartets that pass it.
    const double thr = h->opt.threshold;
    cf_stats& st = h->stats;
    st = cf_stats{};
    st.nshell = ns; st.nbf = nbf; st.ncart = ncart;
    st.ref_repulsion_length = h->ref_counts[0]; st.ref_shell_quartet_length = h->ref_counts[1];
    st.shell_pairs_total = (long long)ns * (ns + 1) / 2;
    st.shell_pairs_kept = pairs_kept;
    struct KetByQ { std::vector<double> qs, pre_prim, pre_primfull, pre_uniq; };
    std::vector<KetByQ> ketq(CF_NCLS);
    for (int cb = 0; cb < CF_NCLS; cb++)
        for (int ck = 0; ck <= cb; ck++) {
            const PairClassHost& B = h->cls[cb];
            const PairClassHost& K = h->cls[ck];
            if (B.npair() == 0 || K.npair() == 0) continue;
            ClassPairTask* t = new ClassPairTask();
            t->bra = cb; t->ket = ck;
            const int nb = B.npair(), nk = K.npair();
            const bool same = (cb == ck);
            const int nfa = h->nfun[B.sa[0]], nfb = h->nfun[B.sb[0]], nfc = h->nfun[K.sa[0]], nfd = h->nfun[K.sb[0]];
            // kets by descending Q with prefix sums of the per-pair weights: a property of the ket class, built once per class
            auto nfun_pair = [](bool diag, int n1, int n2) { return diag ? n1 * (n1 + 1) / 2.0 : (double)n1 * n2; };
            KetByQ& kb = ketq[ck];
            if (kb.qs.empty()) {
                std::vector<int> kq(nk);
                {   // order by descending Q: device radix sort on the class's bound array (ties in creation order)
                    DevBuf<unsigned long long> key; DevBuf<int> val;
                    if (key.alloc(nk) != cudaSuccess || val.alloc(nk) != cudaSuccess) { delete t; return fail("cudaMalloc failed (ket order)"); }
                    desc_key_kernel<<<(nk + 255) / 256, 256>>>(nk, K.d_Q.p, key.p, val.p);
                    if (device_sort_pairs(nk, key.p, val.p) != cudaSuccess ||
                        cudaMemcpy(kq.data(), val.p, sizeof(int) * nk, cudaMemcpyDeviceToHost) != cudaSuccess) { delete t; return fail("ket order sort failed"); }
                    key.release(); val.release();
                }
                kb.qs.resize(nk); kb.pre_prim.assign(nk + 1, 0.0); kb.pre_primfull.assign(nk + 1, 0.0); kb.pre_uniq.assign(nk + 1, 0.0);
                for (int k = 0; k < nk; k++) {
                    const int s = kq[k];
                    kb.qs[k] = K.Qpure[s];
                    kb.pre_prim[k + 1] = kb.pre_prim[k] + K.nprim[s];
                    kb.pre_primfull[k + 1] = kb.pre_primfull[k] + K.nprim_full[s];
                    kb.pre_uniq[k + 1] = kb.pre_uniq[k] + nfun_pair(K.sa[s] == K.sb[s], nfc, nfd);
                }
            }
            const std::vector<double>&qs = kb.qs, &pre_prim = kb.pre_prim, &pre_primfull = kb.pre_primfull, &pre_uniq = kb.pre_uniq;
            const int L = B.la + B.lb + K.la + K.lb, nr = L / 2 + 1;
            const double per_prim = nr * (40.0 + 12.0 * (B.la + B.lb + 1) * (K.la + K.lb + 1) +
                                          3.0 * cf_ncart(B.la) * cf_ncart(B.lb) * cf_ncart(K.la) * cf_ncart(K.lb));
            double nq = 0, uniq = 0, primq = 0, primq_full = 0;        // over ordered (bra, ket) combinations
            double dq = 0, duniq = 0, dprimq = 0, dprimq_full = 0;     // diagonal (i, i) part, same class only
            double uniq_fix = 0;
#pragma omp parallel for reduction(+ : nq, uniq, primq, primq_full, dq, duniq, dprimq, dprimq_full, uniq_fix) schedule(static)
            for (int i = 0; i < nb; i++) {
                int cnt = nk;
                if (thr > 0) {
                    const double need = thr / std::max(B.Qpure[i], 1e-300);
                    cnt = (int)(std::partition_point(qs.begin(), qs.end(), [&](double q) { return q > need; }) - qs.begin());
                }
                const double nab = nfun_pair(B.sa[i] == B.sb[i], nfa, nfb);
                nq += cnt; uniq += nab * pre_uniq[cnt];
                primq += (double)B.nprim[i] * pre_prim[cnt]; primq_full += (double)B.nprim_full[i] * pre_primfull[cnt];
                if (same && (thr <= 0 || B.Qpure[i] * B.Qpure[i] > thr)) {
                    dq += 1; duniq += nab * nab; dprimq += (double)B.nprim[i] * B.nprim[i];
                    dprimq_full += (double)B.nprim_full[i] * B.nprim_full[i];
                    uniq_fix += nab * nab - nab * (nab + 1) / 2.0;   // (ab|ab): only the upper triangle of functions is unique
                }
            }
            if (same) {
                nq = 0.5 * (nq + dq); uniq = 0.5 * (uniq + duniq) - uniq_fix;
                primq = 0.5 * (primq + dprimq); primq_full = 0.5 * (primq_full + dprimq_full);
            }
            t->nquartet = same ? (long long)nb * (nb + 1) / 2 : (long long)nb * nk;
            const long long nq_kept = (long long)(nq + 0.5);
            if (nq_kept == 0) { delete t; continue; }
            QuartetTask dummy{};
            for (int k = 0; k < 4; k++) { dummy.nk = k; g_bra_launch[cb](ck, dummy, 0, 0, 0, &t->G, &t->smem[k], &t->kind); }
            t->nq_item = t->kind >> 4; t->swap = (t->kind >> 3) & 1; t->kind &= 7;
            // thread-per-quartet classes: tasks with few primitive quartets per shell quartet are bound by the digestion
            // (atomics, latency) -> bra-loop kernel; deeply contracted ones are FP64-bound -> one bra pair per item
            {   // the larger register-resident classes gain less from the bra loop and pay for its extra registers sooner
                const int nout = cf_ncart(B.la) * cf_ncart(B.lb) * cf_ncart(K.la) * cf_ncart(K.lb);
                const double kmax = nout <= 18 ? CF_BRALOOP_MAXK : 0.375 * CF_BRALOOP_MAXK;
                t->braloop = (t->kind == 1 && primq < kmax * nq) ? 1 : 0;
            }
            if (const char* e = getenv("CF_BRALOOP")) t->braloop = (t->kind == 1 && atoi(e) != 0) ? 1 : 0;   // developer override (A/B)
            if (t->braloop) {   // bra-loop kernels: (bra chunk of one a-group) x (aligned block of 32 kets)
                // chunk length: long chunks amortise the J(c,d)/K(a,.) flushes, but the static schedule wants >= ~32 items
                // per resident warp and no item heavier than ~1/16 of a warp's share (cost ~ bra primitives x ket primitives)
                const double nbatch = (double)nb * K.nblk * (same ? 0.5 : 1.0);
                const int lch = (int)std::max(1.0, std::min(64.0, std::floor(nbatch / (148.0 * 8 * 32))));
                double sumB = 0, sumK = 0; int maxK = 1;
                for (int i = 0; i < nb; i++) sumB += B.nprim[i];
                for (int i = 0; i < nk; i++) { sumK += K.nprim[i]; maxK = std::max(maxK, K.nprim[i]); }
                const double pcap = std::max(1.0, sumB * sumK * (same ? 0.5 : 1.0) / (148.0 * 8 * 16 * 32 * maxK));
                std::vector<int> cstart, ccnt, cmax; std::vector<double> cq;
                B.make_chunks(lch, pcap, cstart, ccnt, cmax, cq);
                const int nH = (int)cstart.size();
                DevBuf<int> d_cs, d_cc, d_cm; DevBuf<double> d_cq;
                DevBuf<long long> d_cnt, d_off;
                if (d_cs.upload(cstart) != cudaSuccess || d_cc.upload(ccnt) != cudaSuccess || d_cm.upload(cmax) != cudaSuccess || d_cq.upload(cq) != cudaSuccess ||
                    d_cnt.alloc(nH) != cudaSuccess || d_off.alloc(nH) != cudaSuccess) { delete t; return fail("cudaMalloc failed (item counts)"); }
                const int tb = 128, gb = (nH + tb - 1) / tb;
                item_list_a_kernel<<<gb, tb>>>(nH, d_cs.p, d_cc.p, d_cm.p, d_cq.p, K.nblk, nk, K.d_blkQ.p,
                                               thr, same ? 1 : 0, d_cnt.p, nullptr, nullptr);
                std::vector<long long> cnt(nH), off(nH);
                if (cudaMemcpy(cnt.data(), d_cnt.p, sizeof(long long) * nH, cudaMemcpyDeviceToHost) != cudaSuccess) { delete t; return fail("item count kernel failed"); }
                long long tot = 0;
                for (int i = 0; i < nH; i++) { off[i] = tot; tot += cnt[i]; }
                t->nitem = tot;
                if (tot > 0) {
                    if (cudaMemcpy(d_off.p, off.data(), sizeof(long long) * nH, cudaMemcpyHostToDevice) != cudaSuccess ||
                        t->d_items.alloc((size_t)tot) != cudaSuccess) { delete t; return fail("cudaMalloc failed (work items)"); }
                    item_list_a_kernel<<<gb, tb>>>(nH, d_cs.p, d_cc.p, d_cm.p, d_cq.p, K.nblk, nk, K.d_blkQ.p,
                                                   thr, same ? 1 : 0, nullptr, d_off.p, t->d_items.p);
                    if (cudaDeviceSynchronize() != cudaSuccess) { delete t; return fail("item list kernel failed"); }
                }
                d_cnt.release(); d_off.release(); d_cs.release(); d_cc.release(); d_cm.release(); d_cq.release();
            } else if (t->kind >= 1) {   // work items: (CTA-uniform pair, run of <= nq_item spread pairs); swap: roles exchanged
                const PairClassHost& Hc = t->swap ? K : B;
                const PairClassHost& Sc = t->swap ? B : K;
                const int nH = Hc.npair(), nseg = (int)Sc.seg.size() - 1;
                DevBuf<long long> d_cnt, d_off;
                if (d_cnt.alloc(nH) != cudaSuccess || d_off.alloc(nH) != cudaSuccess) { delete t; return fail("cudaMalloc failed (item counts)"); }
                const int tb = 128, gb = (nH + tb - 1) / tb;
                item_list_kernel<<<gb, tb>>>(nH, Hc.d_Q.p, Sc.d_Q.p, Sc.d_seg.p, nseg, t->nq_item, thr, same ? 1 : 0, d_cnt.p, nullptr, nullptr);
                std::vector<long long> cnt(nH), off(nH);
                if (cudaMemcpy(cnt.data(), d_cnt.p, sizeof(long long) * nH, cudaMemcpyDeviceToHost) != cudaSuccess) { delete t; return fail("item count kernel failed"); }
                long long tot = 0;
                for (int i = 0; i < nH; i++) { off[i] = tot; tot += cnt[i]; }
                t->nitem = tot;
                if (tot > 0) {
                    if (cudaMemcpy(d_off.p, off.data(), sizeof(long long) * nH, cudaMemcpyHostToDevice) != cudaSuccess ||
                        t->d_items.alloc((size_t)tot) != cudaSuccess) { delete t; return fail("cudaMalloc failed (work items)"); }
                    item_list_kernel<<<gb, tb>>>(nH, Hc.d_Q.p, Sc.d_Q.p, Sc.d_seg.p, nseg, t->nq_item, thr, same ? 1 : 0, nullptr, d_off.p, t->d_items.p);
                    if (cudaDeviceSynchronize() != cudaSuccess) { delete t; return fail("item list kernel failed"); }
                }
                d_cnt.release(); d_off.release();
            } else {
                std::vector<long long> qoff(nb + 1, 0);
                for (int i = 0; i < nb; i++) qoff[i + 1] = qoff[i] + (same ? i + 1 : nk);
                if (t->d_qoff.upload(qoff) != cudaSuccess) { delete t; return fail("qoff upload failed"); }
            }
            t->flops_eri = primq_full * per_prim;
            t->per_prim = per_prim;
            t->nfun_q = (double)nfa * nfb * nfc * nfd;
            t->nfun_sum = (double)nq_kept * t->nfun_q;
            st.canonical_quartets += nq_kept;
            st.unique_integrals += (long long)(uniq + 0.5);
            st.primitive_quartets += (long long)(primq + 0.5);
            for (int k = 0; k < 4; k++) st.flops_alg_jk[k] += t->flops_eri + 2.0 * (2 + 4 * k) * t->nfun_sum;
            {   // gradient model: K n_r' [40 + 12 (L_ab+2)(L_cd+2) + 48 N_c] + 32 N_c per quartet
                const int nr1 = (L + 1) / 2 + 1;
                const double ncq = (double)cf_ncart(B.la) * cf_ncart(B.lb) * cf_ncart(K.la) * cf_ncart(K.lb);
                st.flops_alg_grad += primq_full * nr1 * (40.0 + 12.0 * (B.la + B.lb + 2) * (K.la + K.lb + 2) + 48.0 * ncq) + 32.0 * ncq * (double)nq_kept;
            }
            h->tasks.push_back(t);
        }
    mark("tasks: stats + item lists");
    st.canonical_quartets_local = st.canonical_quartets / h->opt.world_size;
    if (h->opt.world_size > 1) { for (int k = 0; k < 4; k++) st.flops_alg_jk[k] /= h->opt.world_size; st.flops_alg_grad /= h->opt.world_size; }
    // heavy tasks first so the tail of the build is made of small kernels
    std::sort(h->tasks.begin(), h->tasks.end(), [](const ClassPairTask* a, const ClassPairTask* b) { return a->flops_eri > b->flops_eri; });
    for (size_t i = 0; i < h->tasks.size(); i++) h->tasks[i]->index = (int)i;
    if (h->opt.world_size > 1) {
        // Static, cost-balanced partition (north_star): large tasks deal their work items round-robin over the ranks; a task
        // whose share per rank would be less than ~2 waves of CTAs cannot get faster by splitting (its duration is one
        // item chain), so it is given WHOLE to the least-loaded rank (longest-processing-time-first over the modelled
        // cost).  Every rank derives the same assignment from the same setup; the integer accumulators make the sum
        // independent of who computed what.
        const int world = h->opt.world_size;
        std::vector<double> load(world, 0.0);
        for (ClassPairTask* t : h->tasks) {          // already sorted by descending cost
            const double cost = t->flops_eri + 12.0 * t->nfun_sum;
            const long long units = t->kind == 0 ? t->nquartet / 16 : t->nitem;
            const long long two_waves = t->kind == 1 ? 148LL * 4 * 4 * 2 : 148LL * 2 * 2;
            const char* e = getenv("CF_TASK_OWNER");
            if ((e ? atoi(e) != 0 : true) && units / world < two_waves) {
                int best = 0;
                for (int r = 1; r < world; r++) if (load[r] < load[best]) best = r;
                t->owner = best; load[best] += cost;
            } else {
                for (int r = 0; r < world; r++) load[r] += cost / world;
            }
        }
    }
    h->cnt_host.assign(h->tasks.size() * CF_CNT_WORDS, 0ull);

    // ---- work space
    const size_t n2p = (size_t)nbf * nbf, n2c = (size_t)ncart * ncart;
    bool ok = true;
    for (auto& b : h->d_Dpure) ok = ok && b.alloc(n2p) == cudaSuccess;
    for (auto& b : h->d_Dcart) ok = ok && b.alloc(n2c) == cudaSuccess;
    for (auto& b : h->d_out) ok = ok && b.alloc(n2p) == cudaSuccess;
    // accumulator of the host calls: up to [J_0..J_2 | K_0..K_2 | Jlo_0..Jlo_2 | tail] (multi-density build)
    h->d_acc.pooled = false;      // NCCL works on it (cf_create_multi)
    ok = ok && h->d_acc.alloc(9 * n2c + CF_ACC_TAIL) == cudaSuccess && h->d_partial.alloc(4 * 256) == cudaSuccess &&
         h->d_cnt.alloc(std::max<size_t>(1, h->tasks.size()) * CF_CNT_WORDS) == cudaSuccess;
    if (!ok) return fail("cudaMalloc failed (work space)");
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(std::string("setup kernels failed: ") + cudaGetErrorString(cudaGetLastError()));
    if (h->opt.verbose > 0) {
        // the reference prints one line per setup stage (Int4C2E.cpp:500-587); the stages are fused here
        std::printf("Calculating diagonal elements of repulsion integrals ... Done in %f s\n", now_s() - t_start);
        std::printf("After screening: %lld integrals and %lld shell quartets\n", (long long)st.ref_repulsion_length, (long long)st.ref_shell_quartet_length);
    }
    return h;
}

extern "C" int cf_nbf(const cf_handle* h) { return h ? h->nbf : -1; }

extern "C" int cf_set_density_threshold(cf_handle* h, double dthr) {
    if (!h) return CF_ERR_BAD_ARGUMENT;
    h->density_threshold = dthr > 0.0 ? dthr : 0.0;
    if (h->multi) multi_set_density_threshold(h, h->density_threshold);
    return CF_OK;
}

extern "C" int cf_get_stats(const cf_handle* h, cf_stats* out) {
    if (!h || !out) return CF_ERR_BAD_ARGUMENT;
    *out = h->stats;
    return CF_OK;
}
static int multi_only_host_calls(cf_handle* h) {
    set_error(h, "this is a multi-device handle (cf_create_multi): use the host calls cf_build_jk / cf_build_g_multi / "
                 "cf_contract_grads; the *_device entry points work on single-device handles");
    return CF_ERR_BAD_ARGUMENT;
}


extern "C" int cf_get_repulsion_diag(cf_handle* h, double* diag1212) {
    if (!h || !diag1212) return CF_ERR_BAD_ARGUMENT;
    if (h->multi) return cf_get_repulsion_diag(multi_part0(h), diag1212);
    if (!h->diag_ready) { set_error(h, "Diagonal elements of repulsion integrals are missing!"); return CF_ERR_STATE; }
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaMemcpy(diag1212, h->d_diag.p, sizeof(double) * (size_t)h->nbf * h->nbf, cudaMemcpyDeviceToHost));
    return CF_OK;
}

// words to ALLOCATE for an accumulator of nk exchange densities / the leading words a multi-GPU caller all-reduces
extern "C" size_t cf_acc_reduce_len(const cf_handle* h, int nk) {
    if (!h || nk < 0 || nk > 3) return 0;
    return (size_t)(2 + nk) * h->ncart * h->ncart;
}
extern "C" size_t cf_acc_len(const cf_handle* h, int nk) {
    const size_t n = cf_acc_reduce_len(h, nk);
    return n ? n + CF_ACC_TAIL : 0;
}

// densities present -> compact list of exchange densities
static int exchange_list(const double* Dd, const double* Da, const double* Db, double exx, const double* out[3], int slot[3]) {
    int nk = 0;
    if (exx > 0.0) {
        if (Dd) { out[nk] = Dd; slot[nk++] = 0; }
        if (Da) { out[nk] = Da; slot[nk++] = 1; }
        if (Db) { out[nk] = Db; slot[nk++] = 2; }
    }
    return nk;
}

extern "C" int cf_accumulate_device(cf_handle* h, int nbf, const double* Dd, const double* Da, const double* Db, double exx,
                                    int64_t* acc, void* stream) {
    if (!h) return CF_ERR_BAD_ARGUMENT;
    if (nbf != h->nbf) { set_error(h, "nbf does not match the basis of this handle"); return CF_ERR_BAD_ARGUMENT; }
    if (!Dd && !Da && !Db) { set_error(h, "at least one density is required"); return CF_ERR_BAD_ARGUMENT; }
    if (!acc) { set_error(h, "null accumulator"); return CF_ERR_BAD_ARGUMENT; }
    if (h->multi) return multi_only_host_calls(h);
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const double* dk[3]; int slot[3];
    const int nk = exchange_list(Dd, Da, Db, exx, dk, slot);
    const int ns = h->nshell, ncart = h->ncart;
    const size_t n2c = (size_t)ncart * ncart;
    double* tail = reinterpret_cast<double*>((long long*)acc + (size_t)(2 + nk) * n2c);     // scales of THIS build
    h->last_tail = (const long long*)tail; h->last_nk = nk;
    h->stats.n_launches_last = 0;
    CUDA_TRY(cudaEventRecord(h->ev[0], s));
    dim3 grid2(ns, ns);
    // total density 2Dd + Da + Db (Int4C2E.cpp:612-615) and the exchange densities, in the Cartesian working basis
    const size_t ns2 = (size_t)ns * ns;
    pure_to_cart_kernel<<<grid2, 64, 0, s>>>(ns, nbf, ncart, Dd, Da, Db, 2.0, 1.0, 1.0, h->d_ctrans.p, h->d_ct_off.p, h->d_bf_off.p,
                                             h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, h->d_Dcart[0].p, h->d_B.p, h->d_Bmax.p);
    h->stats.n_launches_last++;
    for (int x = 0; x < nk; x++) {
        pure_to_cart_kernel<<<grid2, 64, 0, s>>>(ns, nbf, ncart, dk[x], nullptr, nullptr, 1.0, 0.0, 0.0, h->d_ctrans.p, h->d_ct_off.p,
                                                 h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, h->d_Dcart[1 + x].p,
                                                 h->d_B.p + (size_t)(1 + x) * ns2, h->d_Bmax.p + (size_t)(1 + x) * ns2);
        h->stats.n_launches_last++;
    }
    bounds_kernel<<<1 + nk, 1024, 0, s>>>(ns, h->d_QS.p, h->d_B.p, h->d_Bmax.p, h->d_rwork.p, h->d_partial.p);
    scales_kernel<<<1, 32, 0, s>>>(h->d_partial.p, nk, h->qmax_cart, h->opt.threshold, h->density_threshold,
                                   (double)h->stats.shell_pairs_kept, h->opt.j_two_limb, tail);
    h->stats.n_launches_last += 2;
    CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(long long) * (2 + nk) * n2c, s));
    CUDA_TRY(cudaMemsetAsync(h->d_cnt.p, 0, sizeof(unsigned long long) * h->d_cnt.n, s));
    // the scales stay on the device (kernels read them through QuartetTask::scales): no host synchronisation inside the
    // build; the range check happens after the caller's synchronisation (check_scales)
    CUDA_TRY(cudaEventRecord(h->ev[1], s));
    // fan the class-pair kernels out over the caller's stream + 3 side streams (small launches overlap)
    CUDA_TRY(cudaEventRecord(h->ev_fork, s));
    for (int i = 0; i < 3; i++) CUDA_TRY(cudaStreamWaitEvent(h->side[i], h->ev_fork, 0));
    int it = 0;
    for (ClassPairTask* t : h->tasks) {
        QuartetTask qt{};
        qt.rank = h->opt.rank; qt.world = h->opt.world_size;
        qt.ncart = ncart; qt.nk = nk;
        qt.Dtot = h->d_Dcart[0].p;
        for (int x = 0; x < nk; x++) { qt.Dk[x] = h->d_Dcart[1 + x].p; qt.accK[x] = (long long*)acc + (size_t)(1 + x) * n2c; }
        qt.accJ = (long long*)acc;
        qt.nj = 1; qt.Dj[0] = qt.Dtot; qt.accJm[0] = qt.accJ;
        qt.jlo_off = (long long)((size_t)(1 + nk) * n2c);
        qt.scales = tail; qt.cnt = h->d_cnt.p + (size_t)t->index * CF_CNT_WORDS;
        qt.store = nullptr; qt.diag = 0; qt.prim_cut = cf_prim_cut();
        fill_rys(qt, h);
        cudaStream_t ts = (it % 4 == 0) ? s : h->side[it % 4 - 1];
        it++;
        int rc = launch_task(h, t, qt, 0, ts);
        if (rc != CF_OK) return rc;
    }
    for (int i = 0; i < 3; i++) { CUDA_TRY(cudaEventRecord(h->ev_join[i], h->side[i])); CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join[i], 0)); }
    CUDA_TRY(cudaEventRecord(h->ev[2], s));
    return CF_OK;
}

extern "C" int cf_finalize_device(cf_handle* h, int nbf, const int64_t* acc, double exx, int has_d, int has_a, int has_b,
                                  double* J, double* Kd, double* Ka, double* Kb, void* stream) {
    if (!h || !acc || !J) return CF_ERR_BAD_ARGUMENT;
    if (nbf != h->nbf) { set_error(h, "nbf does not match the basis of this handle"); return CF_ERR_BAD_ARGUMENT; }
    if (h->multi) return multi_only_host_calls(h);
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int ns = h->nshell, ncart = h->ncart;
    const size_t n2c = (size_t)ncart * ncart, n2p = (size_t)nbf * nbf;
    const int nk = exx > 0.0 ? (has_d != 0) + (has_a != 0) + (has_b != 0) : 0;      // layout of THIS accumulator
    const long long* a = (const long long*)acc;
    const double* tail = reinterpret_cast<const double*>(a + (size_t)(2 + nk) * n2c);
    dim3 grid2(ns, ns);
    // J = 1/4 (raw + raw^T), K = 1/8 (raw + raw^T) * EXX   (Int4C2E.cpp:661-670)
    finalize_kernel<<<grid2, 64, 0, s>>>(nbf, ncart, a, a + (size_t)(1 + nk) * n2c, tail, 0, 0.25, h->d_ctrans.p, h->d_ct_off.p, h->d_bf_off.p,
                                         h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, J);
    h->stats.n_launches_last++;
    double* outs[3] = {Kd, Ka, Kb};
    const int has[3] = {has_d, has_a, has_b};
    int x = 0;
    for (int k = 0; k < 3; k++) {
        if (!has[k]) continue;
        if (!outs[k]) { set_error(h, "K output missing for a density that was given"); return CF_ERR_BAD_ARGUMENT; }
        if (exx > 0.0) {
            finalize_kernel<<<grid2, 64, 0, s>>>(nbf, ncart, a + (size_t)(1 + x) * n2c, nullptr, tail, 1, 0.125 * exx,
                                                 h->d_ctrans.p, h->d_ct_off.p, h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, outs[k]);
            h->stats.n_launches_last++;
            x++;
        } else {
            CUDA_TRY(cudaMemsetAsync(outs[k], 0, sizeof(double) * n2p, s));   // EXX <= 0: K returned as zeros (Int4C2E.cpp:638)
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev[3], s));
    return CF_OK;
}

extern "C" int cf_build_jk_device(cf_handle* h, int nbf, const double* Dd, const double* Da, const double* Db, double exx,
                                  double* J, double* Kd, double* Ka, double* Kb, void* stream) {
    if (!h) return CF_ERR_BAD_ARGUMENT;
    if (h->multi) return multi_only_host_calls(h);
    DeviceGuard guard(h->device);
    int rc = cf_accumulate_device(h, nbf, Dd, Da, Db, exx, (int64_t*)h->d_acc.p, stream);
    if (rc != CF_OK) return rc;
    rc = cf_finalize_device(h, nbf, (const int64_t*)h->d_acc.p, exx, Dd != nullptr, Da != nullptr, Db != nullptr, J, Kd, Ka, Kb, stream);
    return rc;
}

// after a build has been synchronised: J/K scales and bounds of that build -> stats, the work counters, and the range
// check.  The scale leaves 2^-10 of head room above the rigorous bound of the final values; the only way to leave the
// representable range is a non-finite or astronomically large density
static int check_scales(cf_handle* h) {
    const double* sc = h->scales_host;
    h->stats.fixedpoint_scale_log2[0] = std::log2(sc[0]);
    h->stats.fixedpoint_scale_log2[1] = std::log2(sc[1]);
    h->stats.threshold_effective_last = sc[4];
    h->stats.j_two_limb_last = sc[6] != 0.0 ? 1 : 0;
    h->stats.j_rounding_estimate_last = sc[7];
    {   // per-task counters -> evaluated shell quartets, EXECUTED primitive quartets and the model flops they stand for
        unsigned long long nq = 0, np = 0;
        double fl = 0.0;
        for (const ClassPairTask* t : h->tasks) {
            const unsigned long long* c = h->cnt_host.data() + (size_t)t->index * CF_CNT_WORDS;
            unsigned long long q = 0, pq = 0;
            for (int i = 0; i < CF_CNT_SLOTS; i++) { q += c[i]; pq += c[CF_CNT_SLOTS + i]; }
            nq += q; np += pq;
            fl += (double)pq * t->per_prim + 2.0 * (2 + 4 * h->last_nk) * (double)q * t->nfun_q;
        }
        h->stats.quartets_evaluated_last = (int64_t)nq;
        h->stats.primitive_quartets_executed_last = (int64_t)np;
        h->stats.flops_executed_last = fl;
    }
    if (!(sc[0] > 0x1p-900) || !(sc[1] > 0x1p-900) || !std::isfinite(sc[2]) || !std::isfinite(sc[3])) {
        set_error(h, "fixed-point accumulator range exceeded: density contains non-finite or astronomically large entries");
        return CF_ERR_RANGE;
    }
    return CF_OK;
}

static int fetch_times(cf_handle* h) {
    float a = 0, b = 0;
    if (cudaEventElapsedTime(&a, h->ev[0], h->ev[3]) == cudaSuccess) h->stats.ms_device_last = a;
    if (cudaEventElapsedTime(&b, h->ev[1], h->ev[2]) == cudaSuccess) h->stats.ms_eri_last = b;
    (void)cudaGetLastError();   // events not recorded yet are not an error worth keeping
    return CF_OK;
}

// scale tail + work counters of the last build -> host (enqueued on stream s)
static int fetch_build_info(cf_handle* h, cudaStream_t s) {
    if (!h->last_tail) return CF_OK;
    CUDA_TRY(cudaMemcpyAsync(h->scales_host, h->last_tail, sizeof(h->scales_host), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h->cnt_host.data(), h->d_cnt.p, sizeof(unsigned long long) * h->cnt_host.size(), cudaMemcpyDeviceToHost, s));
    return CF_OK;
}

extern "C" int cf_sync_stats(cf_handle* h) {   // after a *_device call has been synchronised by the caller
    if (!h) return CF_ERR_BAD_ARGUMENT;
    if (h->multi) return CF_OK;                // the host calls of a multi-device handle refresh the statistics themselves
    DeviceGuard guard(h->device);
    fetch_times(h);
    int rc = fetch_build_info(h, 0);
    if (rc != CF_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(0));
    return check_scales(h);
}

extern "C" int cf_build_jk(cf_handle* h, int nbf, const double* Dd, const double* Da, const double* Db, double exx,
                           double* J, double* Kd, double* Ka, double* Kb) {
    if (!h) return CF_ERR_BAD_ARGUMENT;
    if (nbf != h->nbf) { set_error(h, "nbf does not match the basis of this handle"); return CF_ERR_BAD_ARGUMENT; }
    if (!J) { set_error(h, "J output is required"); return CF_ERR_BAD_ARGUMENT; }
    if (!Dd && !Da && !Db) { set_error(h, "at least one density is required"); return CF_ERR_BAD_ARGUMENT; }
    if ((Dd && !Kd) || (Da && !Ka) || (Db && !Kb)) { set_error(h, "K output missing for a density that was given"); return CF_ERR_BAD_ARGUMENT; }
    if (h->multi) return multi_build_jk(h, nbf, Dd, Da, Db, exx, J, Kd, Ka, Kb);
    DeviceGuard guard(h->device);
    const double t0 = now_s();
    if (h->opt.verbose > 0) std::printf("Contracting 4c-2e repulsion integrals with 1 matrix ... ");
    const size_t bytes = sizeof(double) * (size_t)nbf * nbf;
    const double* src[3] = {Dd, Da, Db};
    const double* dev[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; k++)
        if (src[k]) { CUDA_TRY(cudaMemcpyAsync(h->d_Dpure[k].p, src[k], bytes, cudaMemcpyHostToDevice, 0)); dev[k] = h->d_Dpure[k].p; }
    int rc = cf_build_jk_device(h, nbf, dev[0], dev[1], dev[2], exx, h->d_out[0].p, h->d_out[1].p, h->d_out[2].p, h->d_out[3].p, nullptr);
    if (rc != CF_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(J, h->d_out[0].p, bytes, cudaMemcpyDeviceToHost, 0));
    double* outs[3] = {Kd, Ka, Kb};
    for (int k = 0; k < 3; k++)
        if (src[k]) CUDA_TRY(cudaMemcpyAsync(outs[k], h->d_out[1 + k].p, bytes, cudaMemcpyDeviceToHost, 0));
    rc = fetch_build_info(h, 0);
    if (rc != CF_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(0));
    CUDA_TRY(cudaGetLastError());
    fetch_times(h);
    rc = check_scales(h);
    if (h->opt.verbose > 0) std::printf("Done in %f s\n", now_s() - t0);
    return rc;
}

// Per-class-pair timing of the last densities' build (serialised, CUDA events): the measurement behind the
// per-kernel roofline table.  rows of 8 doubles: bra class, ket class, quartets, ms, F_alg at nominal contraction (for nk),
// group size, primitive quartets EXECUTED by the launch, model flops of those (+ digestion of the evaluated quartets)
extern "C" int cf_profile_tasks(cf_handle* h, int nbf, const double* Dd_dev, const double* Da_dev, const double* Db_dev, double exx,
                                double* rows, int max_rows, int* nrows) {
    if (!h || !rows || !nrows) return CF_ERR_BAD_ARGUMENT;
    if (h->multi) return multi_only_host_calls(h);
    DeviceGuard guard(h->device);
    int rc = cf_accumulate_device(h, nbf, Dd_dev, Da_dev, Db_dev, exx, (int64_t*)h->d_acc.p, nullptr);   // sets densities + scales
    if (rc != CF_OK) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    const double* dk[3]; int slot[3];
    const int nk = exchange_list(Dd_dev, Da_dev, Db_dev, exx, dk, slot);
    const size_t n2c = (size_t)h->ncart * h->ncart;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int n = 0;
    std::vector<unsigned long long> c(CF_CNT_WORDS);
    for (ClassPairTask* t : h->tasks) {
        if (n >= max_rows) break;
        QuartetTask qt{};
        qt.rank = h->opt.rank; qt.world = h->opt.world_size; qt.ncart = h->ncart; qt.nk = nk;
        qt.Dtot = h->d_Dcart[0].p;
        for (int x = 0; x < nk; x++) { qt.Dk[x] = h->d_Dcart[1 + x].p; qt.accK[x] = h->d_acc.p + (size_t)(1 + x) * n2c; }
        qt.accJ = h->d_acc.p; qt.scales = reinterpret_cast<const double*>(h->d_acc.p + (size_t)(2 + nk) * n2c); qt.prim_cut = cf_prim_cut();
        qt.nj = 1; qt.Dj[0] = qt.Dtot; qt.accJm[0] = qt.accJ; qt.jlo_off = (long long)((size_t)(1 + nk) * n2c);
        qt.cnt = h->d_cnt.p + (size_t)t->index * CF_CNT_WORDS;
        fill_rys(qt, h);
        float best = 1e30f;
        for (int rep = 0; rep < 2; rep++) {
            CUDA_TRY(cudaMemsetAsync(qt.cnt, 0, sizeof(unsigned long long) * CF_CNT_WORDS, 0));
            cudaEventRecord(e0, 0);
            rc = launch_task(h, t, qt, 0, 0);
            cudaEventRecord(e1, 0);
            if (rc != CF_OK) return rc;
            CUDA_TRY(cudaEventSynchronize(e1));
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        CUDA_TRY(cudaMemcpy(c.data(), qt.cnt, sizeof(unsigned long long) * CF_CNT_WORDS, cudaMemcpyDeviceToHost));
        unsigned long long q = 0, pq = 0;
        for (int i = 0; i < CF_CNT_SLOTS; i++) { q += c[i]; pq += c[CF_CNT_SLOTS + i]; }
        double* r = rows + 8 * n;
        r[0] = t->bra; r[1] = t->ket; r[2] = (double)t->nquartet / h->opt.world_size; r[3] = best;
        r[4] = (t->flops_eri + 2.0 * (2 + 4 * nk) * t->nfun_sum) / h->opt.world_size; r[5] = t->G;
        r[6] = (double)pq; r[7] = (double)pq * t->per_prim + 2.0 * (2 + 4 * nk) * (double)q * t->nfun_q;
        n++;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *nrows = n;
    return CF_OK;
}

// fixed-order sum of the per-CTA rows of the gradient kernels
__global__ void grad_reduce_kernel(const double* __restrict__ gpart, int nrow, int ngrad, double* __restrict__ grad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ngrad) return;
    double s = 0.0;
    for (int r = 0; r < nrow; r++) s += gpart[(size_t)r * ngrad + j];
    grad[j] = s;
}

// Int4C2E::ContractGrads(D1, D2, output) (Int4C2E.cpp:747-763) on top of getRepulsion1 (:312-408); this partition's share
extern "C" int cf_contract_grads(cf_handle* h, int nbf, const double* D1, const double* D2, double exx, int natom, double* grad) {
    if (!h || !D1 || !D2 || !grad || natom <= 0) return CF_ERR_BAD_ARGUMENT;
    if (nbf != h->nbf) { set_error(h, "nbf does not match the basis of this handle"); return CF_ERR_BAD_ARGUMENT; }
    if (h->multi) return multi_contract_grads(h, nbf, D1, D2, exx, natom, grad);
    if (h->shell2atom.empty()) { set_error(h, "cf_contract_grads needs cf_basis.shell2atom"); return CF_ERR_BAD_ARGUMENT; }
    for (int a : h->shell2atom) if (a < 0 || a >= natom) { set_error(h, "shell2atom entry outside [0, natom)"); return CF_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    const int ns = h->nshell, ncart = h->ncart, ngrad = 3 * natom;
    const size_t bytes = sizeof(double) * (size_t)nbf * nbf, ns2 = (size_t)ns * ns;
    const int max_grid = 148 * 8;
    if (h->d_shell2atom.n == 0 && h->d_shell2atom.upload(h->shell2atom) != cudaSuccess) { set_error(h, "cudaMalloc failed (shell2atom)"); return CF_ERR_CUDA; }
    if (h->d_gpart.n < (size_t)max_grid * ngrad && h->d_gpart.alloc((size_t)max_grid * ngrad) != cudaSuccess) { set_error(h, "cudaMalloc failed (gradient rows)"); return CF_ERR_CUDA; }
    if (h->d_grad.n < (size_t)ngrad && h->d_grad.alloc(ngrad) != cudaSuccess) { set_error(h, "cudaMalloc failed (gradient)"); return CF_ERR_CUDA; }
    CUDA_TRY(cudaMemcpyAsync(h->d_Dpure[0].p, D1, bytes, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemcpyAsync(h->d_Dpure[1].p, D2, bytes, cudaMemcpyHostToDevice, 0));
    dim3 grid2(ns, ns);
    for (int k = 0; k < 2; k++)
        pure_to_cart_kernel<<<grid2, 64>>>(ns, nbf, ncart, h->d_Dpure[k].p, nullptr, nullptr, 1.0, 0.0, 0.0, h->d_ctrans.p, h->d_ct_off.p,
                                            h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, h->d_Dcart[1 + k].p,
                                            h->d_B.p + (size_t)k * ns2, h->d_Bmax.p + (size_t)k * ns2);
    CUDA_TRY(cudaMemsetAsync(h->d_gpart.p, 0, sizeof(double) * (size_t)max_grid * ngrad, 0));
    CUDA_TRY(cudaEventRecord(h->ev[1], 0));
    int nlaunch = 2;
    for (ClassPairTask* t : h->tasks) {
        const PairClassHost& B = h->cls[t->bra];
        const PairClassHost& K = h->cls[t->ket];
        if (t->nquartet == 0) continue;
        if (t->d_qoff.n == 0) {       // first gradient call: quartet offsets of the CTA-per-quartet enumeration
            const int nb = B.npair(), nk = K.npair();
            std::vector<long long> qoff(nb + 1, 0);
            for (int i = 0; i < nb; i++) qoff[i + 1] = qoff[i] + (t->bra == t->ket ? i + 1 : nk);
            if (t->d_qoff.upload(qoff) != cudaSuccess) { set_error(h, "qoff upload failed"); return CF_ERR_CUDA; }
        }
        GradTask gt{};
        gt.bra = B.dev(); gt.ket = K.dev();
        gt.bra_aexp = B.d_aexp.p; gt.ket_aexp = K.d_aexp.p;
        gt.qoff = t->d_qoff.p; gt.nquartet = t->nquartet;
        gt.rank = h->opt.rank; gt.world = h->opt.world_size;
        gt.same_class = (t->bra == t->ket); gt.ncart = ncart;
        gt.D1 = h->d_Dcart[1].p; gt.D2 = h->d_Dcart[2].p; gt.exx = exx;
        gt.shell2atom = h->d_shell2atom.p; gt.gpart = h->d_gpart.p; gt.ngrad = ngrad;
        gt.prim_cut = 1e-22; gt.thr = h->opt.threshold > 0 ? h->opt.threshold : 0.0;
        fill_rys_tables(gt.rys, h);
        long long chunk = t->nquartet / ((long long)max_grid * 4 * std::max(1, gt.world));
        chunk = std::max(1LL, std::min(64LL, chunk));
        gt.chunk = (int)chunk;
        int G = 0;
        g_grad_launch[t->bra](t->ket, gt, 0, 0, &G, nullptr);       // query: G < 0 = thread-per-quartet kernel, |G| quartets per CTA block
        const long long per = G < 0 ? -G : chunk;
        const long long nchunk_total = (t->nquartet + per - 1) / per;
        const long long nchunk_local = (nchunk_total - gt.rank + gt.world - 1) / gt.world;
        if (nchunk_local <= 0) continue;
        const int grid = (int)std::min<long long>(nchunk_local, max_grid);
        cudaError_t e = g_grad_launch[t->bra](t->ket, gt, grid, 0, nullptr, nullptr);
        if (e != cudaSuccess) { set_error(h, std::string("gradient kernel launch failed: ") + cudaGetErrorString(e)); return CF_ERR_CUDA; }
        nlaunch++;
    }
    CUDA_TRY(cudaEventRecord(h->ev[2], 0));
    grad_reduce_kernel<<<(ngrad + 127) / 128, 128>>>(h->d_gpart.p, max_grid, ngrad, h->d_grad.p);
    CUDA_TRY(cudaMemcpyAsync(grad, h->d_grad.p, sizeof(double) * ngrad, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    CUDA_TRY(cudaGetLastError());
    {   // CUDA-event time of the gradient kernels of this call
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->stats.ms_grad_last = ms;
        h->stats.n_launches_last = nlaunch + 1;
    }
    return CF_OK;
}

// Int4C2E::ContractHesss(D1, D2, output) (Int4C2E.cpp:792-811) on top of getRepulsion2 (:410-492); this partition's share.
// hess: [3*natom][3*natom] col-major (symmetric).  The second-derivative integrals are contracted with the two-particle
// density on the fly (eri_hess.cuh); FP64 atomics into the small matrix (not the SCF loop: one call per `derivative 2` job).
extern "C" int cf_contract_hess(cf_handle* h, int nbf, const double* D, double exx, int natom, double* hess) {
    if (!h || !D || !hess || natom <= 0) return CF_ERR_BAD_ARGUMENT;
    if (nbf != h->nbf) { set_error(h, "nbf does not match the basis of this handle"); return CF_ERR_BAD_ARGUMENT; }
    if (h->multi) return multi_contract_hess(h, nbf, D, exx, natom, hess);
    if (h->shell2atom.empty()) { set_error(h, "cf_contract_hess needs cf_basis.shell2atom"); return CF_ERR_BAD_ARGUMENT; }
    for (int a : h->shell2atom) if (a < 0 || a >= natom) { set_error(h, "shell2atom entry outside [0, natom)"); return CF_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    const int ns = h->nshell, ncart = h->ncart, nh = 3 * natom;
    const size_t bytes = sizeof(double) * (size_t)nbf * nbf, ns2 = (size_t)ns * ns, nh2 = (size_t)nh * nh;
    const int max_grid = 148 * 8;
    if (h->d_shell2atom.n == 0 && h->d_shell2atom.upload(h->shell2atom) != cudaSuccess) { set_error(h, "cudaMalloc failed (shell2atom)"); return CF_ERR_CUDA; }
    if (h->d_hess.n < nh2 && h->d_hess.alloc(nh2) != cudaSuccess) { set_error(h, "cudaMalloc failed (hessian)"); return CF_ERR_CUDA; }
    CUDA_TRY(cudaMemcpyAsync(h->d_Dpure[0].p, D, bytes, cudaMemcpyHostToDevice, 0));
    dim3 grid2(ns, ns);
    pure_to_cart_kernel<<<grid2, 64>>>(ns, nbf, ncart, h->d_Dpure[0].p, nullptr, nullptr, 1.0, 0.0, 0.0, h->d_ctrans.p, h->d_ct_off.p,
                                        h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, h->d_Dcart[1].p, h->d_B.p, h->d_Bmax.p);
    CUDA_TRY(cudaMemsetAsync(h->d_hess.p, 0, sizeof(double) * nh2, 0));
    CUDA_TRY(cudaEventRecord(h->ev[1], 0));
    int nlaunch = 1;
    for (ClassPairTask* t : h->tasks) {
        const PairClassHost& B = h->cls[t->bra];
        const PairClassHost& K = h->cls[t->ket];
        if (t->nquartet == 0) continue;
        if (t->d_qoff.n == 0) {       // quartet offsets of the CTA-per-quartet enumeration (shared with the gradient calls)
            const int nb = B.npair(), nk = K.npair();
            std::vector<long long> qoff(nb + 1, 0);
            for (int i = 0; i < nb; i++) qoff[i + 1] = qoff[i] + (t->bra == t->ket ? i + 1 : nk);
            if (t->d_qoff.upload(qoff) != cudaSuccess) { set_error(h, "qoff upload failed"); return CF_ERR_CUDA; }
        }
        GradTask gt{};
        gt.bra = B.dev(); gt.ket = K.dev();
        gt.bra_aexp = B.d_aexp.p; gt.ket_aexp = K.d_aexp.p;
        gt.qoff = t->d_qoff.p; gt.nquartet = t->nquartet;
        gt.rank = h->opt.rank; gt.world = h->opt.world_size;
        gt.same_class = (t->bra == t->ket); gt.ncart = ncart;
        gt.D1 = h->d_Dcart[1].p; gt.D2 = h->d_Dcart[1].p; gt.exx = exx;
        gt.shell2atom = h->d_shell2atom.p; gt.gmat = h->d_hess.p; gt.ngrad = nh;
        gt.prim_cut = 1e-22; gt.thr = h->opt.threshold > 0 ? h->opt.threshold : 0.0;
        fill_rys_tables(gt.rys, h);
        long long chunk = t->nquartet / ((long long)max_grid * 4 * std::max(1, gt.world));
        chunk = std::max(1LL, std::min(64LL, chunk));
        gt.chunk = (int)chunk;
        const long long nchunk_total = (t->nquartet + chunk - 1) / chunk;
        const long long nchunk_local = (nchunk_total - gt.rank + gt.world - 1) / gt.world;
        if (nchunk_local <= 0) continue;
        const int grid = (int)std::min<long long>(nchunk_local, max_grid);
        cudaError_t e = g_hess_launch[t->bra](t->ket, gt, grid, 0, nullptr, nullptr);
        if (e != cudaSuccess) { set_error(h, std::string("hessian kernel launch failed: ") + cudaGetErrorString(e)); return CF_ERR_CUDA; }
        nlaunch++;
    }
    CUDA_TRY(cudaEventRecord(h->ev[2], 0));
    CUDA_TRY(cudaMemcpyAsync(hess, h->d_hess.p, sizeof(double) * nh2, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    CUDA_TRY(cudaGetLastError());
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->stats.ms_grad_last = ms;
        h->stats.n_launches_last = nlaunch;
    }
    // exact symmetry (the two triangles saw the same terms in different atomic orders)
    for (int x = 0; x < nh; x++)
        for (int y = 0; y < x; y++) {
            const double v = 0.5 * (hess[(size_t)y * nh + x] + hess[(size_t)x * nh + y]);
            hess[(size_t)y * nh + x] = v; hess[(size_t)x * nh + y] = v;
        }
    return CF_OK;
}

// G_pure(block) = C_a (raw + raw^T) C_b^T for one (atom, direction) matrix of the matrix-form gradient (doubles)
__global__ void finalize_gradmat_kernel(int nbf, int ncart, const double* __restrict__ raw, const double* __restrict__ ctrans,
                                        const int* __restrict__ ct_off, const int* __restrict__ bf_off, const int* __restrict__ cao_off,
                                        const int* __restrict__ nfun, const int* __restrict__ ncsh, double* __restrict__ out) {
    const int sa = blockIdx.x, sb = blockIdx.y;
    if (sb > sa) return;
    const size_t n2c = (size_t)ncart * ncart, n2p = (size_t)nbf * nbf;
    raw += (size_t)blockIdx.z * n2c; out += (size_t)blockIdx.z * n2p;
    const int na = nfun[sa], nb = nfun[sb], nca = ncsh[sa], ncb = ncsh[sb];
    const double* Ca = ctrans + ct_off[sa];
    const double* Cb = ctrans + ct_off[sb];
    for (int e = threadIdx.x; e < na * nb; e += blockDim.x) {
        const int m = e / nb, n = e % nb;
        if (sa == sb && n > m) continue;
        double s = 0.0;
        for (int x = 0; x < nca; x++) {
            const double cam = Ca[m * nca + x];
            if (cam == 0.0) continue;
            for (int y = 0; y < ncb; y++) {
                const size_t i = cao_off[sa] + x, j = cao_off[sb] + y;
                s = fma(cam * Cb[n * ncb + y], raw[j * ncart + i] + raw[i * ncart + j], s);
            }
        }
        out[(size_t)(bf_off[sb] + n) * nbf + bf_off[sa] + m] = s;
        out[(size_t)(bf_off[sa] + m) * nbf + bf_off[sb] + n] = s;
    }
}

// Int4C2E::ContractGrads(D, output) (Int4C2E.cpp:766-790): the 3*natom matrices G^(atom,xyz)[D] = d/dR (J[2D] - exx K[D]),
// HOST pointers, G = [3*natom][nbf*nbf] col-major (index 3*atom + xyz), each exactly symmetric.  A handle with
// world_size > 1 returns its partition's share (sum over ranks).
extern "C" int cf_contract_grads_matrices(cf_handle* h, int nbf, const double* D, double exx, int natom, double* G) {
    if (!h || !D || !G || natom <= 0) return CF_ERR_BAD_ARGUMENT;
    if (nbf != h->nbf) { set_error(h, "nbf does not match the basis of this handle"); return CF_ERR_BAD_ARGUMENT; }
    if (h->multi) { set_error(h, "cf_contract_grads_matrices works on single-device handles"); return CF_ERR_BAD_ARGUMENT; }
    if (h->shell2atom.empty()) { set_error(h, "cf_contract_grads_matrices needs cf_basis.shell2atom"); return CF_ERR_BAD_ARGUMENT; }
    for (int a : h->shell2atom) if (a < 0 || a >= natom) { set_error(h, "shell2atom entry outside [0, natom)"); return CF_ERR_BAD_ARGUMENT; }
    DeviceGuard guard(h->device);
    const int ns = h->nshell, ncart = h->ncart, ngrad = 3 * natom;
    const size_t n2p = (size_t)nbf * nbf, n2c = (size_t)ncart * ncart, ns2 = (size_t)ns * ns;
    if (h->d_shell2atom.n == 0 && h->d_shell2atom.upload(h->shell2atom) != cudaSuccess) { set_error(h, "cudaMalloc failed (shell2atom)"); return CF_ERR_CUDA; }
    DevBuf<double> raw, outp;
    raw.pooled = false; outp.pooled = false;
    if (raw.alloc((size_t)ngrad * n2c) != cudaSuccess || outp.alloc((size_t)ngrad * n2p) != cudaSuccess) {
        set_error(h, "cudaMalloc failed (3*natom gradient matrices)"); raw.release(); outp.release(); return CF_ERR_CUDA;
    }
    CUDA_TRY(cudaMemsetAsync(raw.p, 0, sizeof(double) * ngrad * n2c, 0));
    CUDA_TRY(cudaMemcpyAsync(h->d_Dpure[0].p, D, sizeof(double) * n2p, cudaMemcpyHostToDevice, 0));
    dim3 grid2(ns, ns);
    pure_to_cart_kernel<<<grid2, 64>>>(ns, nbf, ncart, h->d_Dpure[0].p, nullptr, nullptr, 1.0, 0.0, 0.0, h->d_ctrans.p, h->d_ct_off.p,
                                        h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, h->d_Dcart[1].p, h->d_B.p, h->d_Bmax.p);
    (void)ns2;
    const int max_grid = 148 * 8;
    for (ClassPairTask* t : h->tasks) {
        const PairClassHost& B = h->cls[t->bra];
        const PairClassHost& K = h->cls[t->ket];
        if (t->nquartet == 0) continue;
        if (t->d_qoff.n == 0) {
            const int nb = B.npair(), nk = K.npair();
            std::vector<long long> qoff(nb + 1, 0);
            for (int i = 0; i < nb; i++) qoff[i + 1] = qoff[i] + (t->bra == t->ket ? i + 1 : nk);
            if (t->d_qoff.upload(qoff) != cudaSuccess) { set_error(h, "qoff upload failed"); return CF_ERR_CUDA; }
        }
        GradTask gt{};
        gt.bra = B.dev(); gt.ket = K.dev();
        gt.bra_aexp = B.d_aexp.p; gt.ket_aexp = K.d_aexp.p;
        gt.qoff = t->d_qoff.p; gt.nquartet = t->nquartet;
        gt.rank = h->opt.rank; gt.world = h->opt.world_size;
        gt.same_class = (t->bra == t->ket); gt.ncart = ncart;
        gt.D1 = h->d_Dcart[1].p; gt.D2 = h->d_Dcart[1].p; gt.exx = exx;
        gt.shell2atom = h->d_shell2atom.p; gt.gmat = raw.p; gt.ngrad = ngrad;
        gt.prim_cut = 1e-22; gt.thr = h->opt.threshold > 0 ? h->opt.threshold : 0.0;
        fill_rys_tables(gt.rys, h);
        long long chunk = t->nquartet / ((long long)max_grid * 4 * std::max(1, gt.world));
        chunk = std::max(1LL, std::min(64LL, chunk));
        gt.chunk = (int)chunk;
        const long long nchunk_total = (t->nquartet + chunk - 1) / chunk;
        const long long nchunk_local = (nchunk_total - gt.rank + gt.world - 1) / gt.world;
        if (nchunk_local <= 0) continue;
        const int grid = (int)std::min<long long>(nchunk_local, max_grid);
        cudaError_t e = g_gradmat_launch[t->bra](t->ket, gt, grid, 0, nullptr, nullptr);
        if (e != cudaSuccess) { set_error(h, std::string("gradient-matrix kernel launch failed: ") + cudaGetErrorString(e)); raw.release(); outp.release(); return CF_ERR_CUDA; }
    }
    dim3 grid3(ns, ns, ngrad);
    finalize_gradmat_kernel<<<grid3, 64>>>(nbf, ncart, raw.p, h->d_ctrans.p, h->d_ct_off.p, h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p,
                                           h->d_ncartsh.p, outp.p);
    cudaError_t e = cudaMemcpy(G, outp.p, sizeof(double) * ngrad * n2p, cudaMemcpyDeviceToHost);
    raw.release(); outp.release();
    if (e != cudaSuccess) { set_error(h, std::string("gradient-matrix kernels failed: ") + cudaGetErrorString(e)); return CF_ERR_CUDA; }
    return CF_OK;
}

// B[0] = max_k B[1+k] (block 1-norms and max-norms of the batch): the J accumulators of a multi-density build share one scale
__global__ void rowmax_kernel(int n, int nmat, double* __restrict__ B, double* __restrict__ Bmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = 0.0, m = 0.0;
    for (int k = 0; k < nmat; k++) { a = fmax(a, B[(size_t)(1 + k) * n + i]); m = fmax(m, Bmax[(size_t)(1 + k) * n + i]); }
    B[i] = a; Bmax[i] = m;
}
// G = J - K on the pure matrices of one batch member
__global__ void sub_kernel(size_t n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

// One batch (<= 3 densities, already in d_Dpure[0..nmat-1]) of the multi-density build: ONE pass over the integrals digests
// J[D_k] and K[D_k] of every member (QuartetTask::nj / nk), accumulators [J_0..J_2 | K_0..K_2]; G_k -> d_out[k].
// first half: this partition's raw accumulators; a multi-GPU caller all-reduces the 9 n2c payload words in between
static int build_g_accumulate(cf_handle* h, int nmat, double exx, cudaStream_t s) {
    const int ns = h->nshell, nbf = h->nbf, ncart = h->ncart;
    const size_t n2c = (size_t)ncart * ncart, ns2 = (size_t)ns * ns;
    const int nk = exx > 0.0 ? nmat : 0;
    long long* acc = h->d_acc.p;                 // [J_0..J_2 | K_0..K_2 | Jlo_0..Jlo_2 | tail]
    double* tail = reinterpret_cast<double*>(acc + 9 * n2c);
    h->last_tail = (const long long*)tail; h->last_nk = nk;
    h->stats.n_launches_last = 0;
    CUDA_TRY(cudaEventRecord(h->ev[0], s));
    dim3 grid2(ns, ns);
    for (int k = 0; k < nmat; k++)
        pure_to_cart_kernel<<<grid2, 64, 0, s>>>(ns, nbf, ncart, h->d_Dpure[k].p, nullptr, nullptr, 1.0, 0.0, 0.0, h->d_ctrans.p, h->d_ct_off.p,
                                                 h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, h->d_Dcart[1 + k].p,
                                                 h->d_B.p + (size_t)(1 + k) * ns2, h->d_Bmax.p + (size_t)(1 + k) * ns2);
    rowmax_kernel<<<(int)((ns2 + 255) / 256), 256, 0, s>>>((int)ns2, nmat, h->d_B.p, h->d_Bmax.p);
    bounds_kernel<<<1 + nmat, 1024, 0, s>>>(ns, h->d_QS.p, h->d_B.p, h->d_Bmax.p, h->d_rwork.p, h->d_partial.p);
    scales_kernel<<<1, 32, 0, s>>>(h->d_partial.p, nmat, h->qmax_cart, h->opt.threshold, h->density_threshold,
                                   (double)h->stats.shell_pairs_kept, h->opt.j_two_limb, tail);
    h->stats.n_launches_last += nmat + 3;
    CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(long long) * 9 * n2c, s));
    CUDA_TRY(cudaMemsetAsync(h->d_cnt.p, 0, sizeof(unsigned long long) * h->d_cnt.n, s));
    CUDA_TRY(cudaEventRecord(h->ev[1], s));
    CUDA_TRY(cudaEventRecord(h->ev_fork, s));
    for (int i = 0; i < 3; i++) CUDA_TRY(cudaStreamWaitEvent(h->side[i], h->ev_fork, 0));
    int it = 0;
    for (ClassPairTask* t : h->tasks) {
        QuartetTask qt{};
        qt.rank = h->opt.rank; qt.world = h->opt.world_size;
        qt.ncart = ncart; qt.nk = nk; qt.nj = nmat;
        for (int x = 0; x < nmat; x++) {
            qt.Dj[x] = h->d_Dcart[1 + x].p; qt.accJm[x] = acc + (size_t)x * n2c;
            qt.Dk[x] = h->d_Dcart[1 + x].p; qt.accK[x] = acc + (size_t)(3 + x) * n2c;
        }
        qt.Dtot = qt.Dj[0]; qt.accJ = qt.accJm[0];
        qt.jlo_off = (long long)(6 * n2c);
        qt.scales = tail; qt.cnt = h->d_cnt.p + (size_t)t->index * CF_CNT_WORDS;
        qt.store = nullptr; qt.diag = 0; qt.prim_cut = cf_prim_cut();
        fill_rys(qt, h);
        cudaStream_t ts = (it % 4 == 0) ? s : h->side[it % 4 - 1];
        it++;
        int rc = launch_task(h, t, qt, 0, ts);
        if (rc != CF_OK) return rc;
    }
    for (int i = 0; i < 3; i++) { CUDA_TRY(cudaEventRecord(h->ev_join[i], h->side[i])); CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join[i], 0)); }
    CUDA_TRY(cudaEventRecord(h->ev[2], s));
    return CF_OK;
}
// second half: G_k -> d_out[k]
static int build_g_finalize(cf_handle* h, int nmat, double exx, cudaStream_t s) {
    const int ns = h->nshell, nbf = h->nbf, ncart = h->ncart;
    const size_t n2c = (size_t)ncart * ncart, n2p = (size_t)nbf * nbf;
    const int nk = exx > 0.0 ? nmat : 0;
    long long* acc = h->d_acc.p;
    const double* tail = reinterpret_cast<const double*>(acc + 9 * n2c);
    dim3 grid2(ns, ns);
    // G_k = J[2 D_k] - exx K[D_k] = 1/2 (rawJ + rawJ^T) - exx/8 (rawK + rawK^T)   (Int4C2E.cpp:726-728 with the J of D_k, not 2 D_k)
    for (int k = 0; k < nmat; k++) {
        finalize_kernel<<<grid2, 64, 0, s>>>(nbf, ncart, acc + (size_t)k * n2c, acc + (size_t)(6 + k) * n2c, tail, 0, 0.5, h->d_ctrans.p, h->d_ct_off.p, h->d_bf_off.p,
                                             h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, nk ? h->d_out[3].p : h->d_out[k].p);
        h->stats.n_launches_last++;
        if (nk) {
            finalize_kernel<<<grid2, 64, 0, s>>>(nbf, ncart, acc + (size_t)(3 + k) * n2c, nullptr, tail, 1, 0.125 * exx, h->d_ctrans.p, h->d_ct_off.p,
                                                 h->d_bf_off.p, h->d_cao_off.p, h->d_nfun.p, h->d_ncartsh.p, h->d_Dpure[k].p);   // D_k is consumed: reuse its buffer
            sub_kernel<<<(int)((n2p + 255) / 256), 256, 0, s>>>(n2p, h->d_out[3].p, h->d_Dpure[k].p, h->d_out[k].p);
            h->stats.n_launches_last += 2;
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev[3], s));
    return CF_OK;
}
static int build_g_batch(cf_handle* h, int nmat, double exx) {
    int rc = build_g_accumulate(h, nmat, exx, 0);
    return rc != CF_OK ? rc : build_g_finalize(h, nmat, exx, 0);
}

// G_k = J[2 D_k] - exx K[D_k]  (GhfMultiple, Int4C2E.cpp:685-745): batches of three densities share one pass over the integrals
extern "C" int cf_build_g_multi(cf_handle* h, int nbf, int nmat, const double* Ds, double exx, double* Gs) {
    if (!h || !Ds || !Gs || nmat <= 0) return CF_ERR_BAD_ARGUMENT;
    if (nbf != h->nbf) { set_error(h, "nbf does not match the basis of this handle"); return CF_ERR_BAD_ARGUMENT; }
    if (h->multi) return multi_build_g(h, nbf, nmat, Ds, exx, Gs);
    DeviceGuard guard(h->device);
    const size_t n2 = (size_t)nbf * nbf, bytes = sizeof(double) * n2;
    for (int k0 = 0; k0 < nmat; k0 += 3) {
        const int nb = std::min(3, nmat - k0);
        for (int k = 0; k < nb; k++) CUDA_TRY(cudaMemcpyAsync(h->d_Dpure[k].p, Ds + (size_t)(k0 + k) * n2, bytes, cudaMemcpyHostToDevice, 0));
        int rc = build_g_batch(h, nb, exx);
        if (rc != CF_OK) return rc;
        for (int k = 0; k < nb; k++) CUDA_TRY(cudaMemcpyAsync(Gs + (size_t)(k0 + k) * n2, h->d_out[k].p, bytes, cudaMemcpyDeviceToHost, 0));
        rc = fetch_build_info(h, 0);
        if (rc != CF_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(0));
        CUDA_TRY(cudaGetLastError());
        fetch_times(h);
        rc = check_scales(h);
        if (rc != CF_OK) return rc;
    }
    return CF_OK;
}

// ================================================================================================
// One-electron integrals: Int2C1E::CalculateIntegrals(0, ...) for Overlap / Kinetic / Nuclear
// (src/Integral/Int2C1E.cpp:18-67, :313-333) on the device.  S, T, V: DEVICE pointers, nbf x nbf col-major.
// ================================================================================================
extern "C" int cf_one_electron_device(cf_handle* h, int natom, const double* Z, const double* xyz, double* S, double* T, double* V, void* stream) {
    if (!h || natom <= 0 || !Z || !xyz || !S || !T || !V) return CF_ERR_BAD_ARGUMENT;
    if (h->multi) h = multi_part0(h);
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (h->d_atomZ.n < (size_t)natom && (h->d_atomZ.alloc(natom) != cudaSuccess || h->d_atomxyz.alloc(3 * (size_t)natom) != cudaSuccess)) {
        set_error(h, "cudaMalloc failed (atoms)"); return CF_ERR_CUDA;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_atomZ.p, Z, sizeof(double) * natom, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->d_atomxyz.p, xyz, sizeof(double) * 3 * natom, cudaMemcpyHostToDevice, s));
    OneIntTask t{};
    t.ns = h->nshell; t.nbf = h->nbf; t.natom = natom;
    t.l = h->d_l.p; t.nprim = h->d_nprim_sh.p; t.prim_off = h->d_prim_off_sh.p;
    t.exps = h->d_exps.p; t.coefs = h->d_coefs.p; t.xyz = h->d_xyz.p;
    t.Z = h->d_atomZ.p; t.atom_xyz = h->d_atomxyz.p;
    t.ctrans = h->d_ctrans.p; t.ct_off = h->d_ct_off.p; t.bf_off = h->d_bf_off.p; t.nfun = h->d_nfun.p;
    fill_rys_tables(t.rys, h);
    t.S = S; t.T = T; t.V = V;
    const long long npairs = (long long)h->nshell * (h->nshell + 1) / 2;
    oneint_kernel<<<(unsigned)npairs, ONEINT_THREADS, 0, s>>>(t);
    CUDA_TRY(cudaGetLastError());
    return CF_OK;
}

// HOST pointers out (what Int2C1E stores in Overlap / Kinetic / Nuclear)
extern "C" int cf_one_electron(cf_handle* h, int natom, const double* Z, const double* xyz, double* S, double* T, double* V) {
    if (!h || !S || !T || !V) return CF_ERR_BAD_ARGUMENT;
    cf_handle* p = h->multi ? multi_part0(h) : h;
    DeviceGuard guard(p->device);
    const size_t bytes = sizeof(double) * (size_t)p->nbf * p->nbf;
    int rc = cf_one_electron_device(p, natom, Z, xyz, p->d_out[0].p, p->d_out[1].p, p->d_out[2].p, nullptr);
    if (rc != CF_OK) { if (p != h) h->err = p->err; return rc; }
    h = p;
    CUDA_TRY(cudaMemcpyAsync(S, p->d_out[0].p, bytes, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpyAsync(T, p->d_out[1].p, bytes, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpyAsync(V, p->d_out[2].p, bytes, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return CF_OK;
}

// ================================================================================================
// Multi-device handle (SURVEY 8b "Threading", 8e): ONE process, one worker thread + one stream per GPU.
// The reference is a single C++ process (src/main.cpp:14-63) whose Int4C2E spawns OpenMP workers inside
// ContractInts (Int4C2E.cpp:617-621); here the same call fans out over the GPUs of the box:
//   host D (pinned staging) -> H2D on every device -> each device accumulates ITS static partition of the quartet work
//   -> ncclAllReduce(int64, sum) of the fixed-point accumulators over NVLink -> device 0 finalises -> D2H.
// The integer all-reduce makes the result bit-identical to the single-GPU result for any number of devices.
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process this resolves to the copy torch loaded).
// ================================================================================================
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
        if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllReduce || !GetErrorString) { err = "libnccl.so.2 lacks the expected symbols"; return false; }
        return true;
    }
};

struct MultiCtx {
    int ndev = 0;
    std::vector<int> devices;
    std::vector<cf_handle*> parts;
    std::vector<cudaStream_t> streams;
    std::vector<ncclComm_t> comms;
    NcclApi nccl;
    double* pin = nullptr;             // pinned host staging: [3 inputs | 4 outputs] x nbf^2
    size_t n2 = 0;
    // worker pool: thread i owns device i
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    long epoch = 0;
    int done = 0;
    bool quit = false;
    std::function<int(int)> job;
    std::vector<int> rc;
    std::vector<std::string> errs;

    void worker(int i) {
        long seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return quit || epoch != seen; });
                if (quit) return;
                seen = epoch;
                f = job;
            }
            const int r = f(i);
            {
                std::lock_guard<std::mutex> lk(mu);
                rc[i] = r;
                done++;
            }
            cv_done.notify_one();
        }
    }
    void start() {
        rc.assign(ndev, 0); errs.assign(ndev, "");
        for (int i = 0; i < ndev; i++) th.emplace_back([this, i] { worker(i); });
    }
    // run f(i) on every worker, return the first non-zero code
    int run(const std::function<int(int)>& f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = f; done = 0; epoch++;
        }
        cv_go.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return done == ndev; });
        for (int i = 0; i < ndev; i++) if (rc[i] != CF_OK) return rc[i];
        return CF_OK;
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv_go.notify_all();
        for (auto& t : th) if (t.joinable()) t.join();
        th.clear();
    }
};

static cf_handle* multi_part0(const cf_handle* h) { return h->multi->parts[0]; }

static void multi_set_density_threshold(cf_handle* h, double dthr) {
    for (cf_handle* p : h->multi->parts) p->density_threshold = dthr;
}

static void multi_destroy(cf_handle* h) {
    MultiCtx* M = h->multi;
    if (!M) return;
    M->stop();
    for (int i = 0; i < (int)M->comms.size(); i++) if (M->comms[i] && M->nccl.CommDestroy) M->nccl.CommDestroy(M->comms[i]);
    for (int i = 0; i < (int)M->streams.size(); i++) if (M->streams[i]) { DeviceGuard g(M->devices[i]); cudaStreamDestroy(M->streams[i]); }
    for (cf_handle* p : M->parts) if (p) cf_destroy(p);
    if (M->pin) cudaFreeHost(M->pin);
    delete M;
    h->multi = nullptr;
}

// statistics of the whole job from the partitions (setup counts are identical on every partition)
static void multi_refresh_stats(cf_handle* h) {
    MultiCtx* M = h->multi;
    cf_stats st = M->parts[0]->stats;
    st.canonical_quartets_local = st.canonical_quartets;
    for (int k = 0; k < 4; k++) st.flops_alg_jk[k] = 0;
    st.flops_alg_grad = 0; st.quartets_evaluated_last = 0; st.primitive_quartets_executed_last = 0; st.flops_executed_last = 0;
    st.n_launches_last = 0; st.ms_device_last = 0; st.ms_eri_last = 0; st.ms_grad_last = 0;
    for (cf_handle* p : M->parts) {
        for (int k = 0; k < 4; k++) st.flops_alg_jk[k] += p->stats.flops_alg_jk[k];
        st.flops_alg_grad += p->stats.flops_alg_grad;
        st.quartets_evaluated_last += p->stats.quartets_evaluated_last;
        st.primitive_quartets_executed_last += p->stats.primitive_quartets_executed_last;
        st.flops_executed_last += p->stats.flops_executed_last;
        st.n_launches_last += p->stats.n_launches_last;
        st.ms_device_last = std::max(st.ms_device_last, p->stats.ms_device_last);      // the build takes as long as its slowest device
        st.ms_eri_last = std::max(st.ms_eri_last, p->stats.ms_eri_last);
        st.ms_grad_last = std::max(st.ms_grad_last, p->stats.ms_grad_last);
    }
    h->stats = st;
}

extern "C" cf_handle* cf_create_multi(const cf_basis* basis, const cf_options* opts, int ndev, const int* devices) {
    int navail = 0;
    if (cudaGetDeviceCount(&navail) != cudaSuccess || navail == 0) { set_error(nullptr, "no CUDA device (this engine has no CPU fallback)"); return nullptr; }
    if (ndev <= 0) ndev = navail;                       // all GPUs of the box
    if (ndev > navail) { set_error(nullptr, "cf_create_multi: more devices requested than present"); return nullptr; }
    if (ndev == 1 && !devices) return cf_create(basis, opts);
    cf_handle* h = new cf_handle();
    MultiCtx* M = new MultiCtx();
    h->multi = M;
    M->ndev = ndev;
    for (int i = 0; i < ndev; i++) M->devices.push_back(devices ? devices[i] : i);
    M->parts.assign(ndev, nullptr); M->streams.assign(ndev, nullptr); M->comms.assign(ndev, nullptr);
    auto fail = [&](const std::string& s) -> cf_handle* { set_error(nullptr, s); multi_destroy(h); delete h; return nullptr; };
    std::string err;
    if (!M->nccl.load(err)) return fail("cf_create_multi: " + err);
    M->start();
    cf_options o{};
    if (opts) o = *opts;
    // every partition runs the whole setup on its own device, concurrently
    int rc = M->run([&](int i) {
        cf_options oi = o;
        oi.device = M->devices[i]; oi.rank = i; oi.world_size = ndev; oi.verbose = (i == 0) ? o.verbose : 0;
        cf_handle* p = cf_create(basis, &oi);
        if (!p) { M->errs[i] = g_last_error; return CF_ERR_CUDA; }
        M->parts[i] = p;
        DeviceGuard g(p->device);
        if (cudaStreamCreateWithFlags(&M->streams[i], cudaStreamNonBlocking) != cudaSuccess) { M->errs[i] = "cudaStreamCreate failed"; return CF_ERR_CUDA; }
        return CF_OK;
    });
    if (rc != CF_OK) { for (auto& e : M->errs) if (!e.empty()) return fail("cf_create_multi: " + e); return fail("cf_create_multi: partition setup failed"); }
    ncclResult_t nr = M->nccl.CommInitAll(M->comms.data(), ndev, M->devices.data());
    if (nr != ncclSuccess) return fail(std::string("ncclCommInitAll: ") + M->nccl.GetErrorString(nr));
    h->device = M->devices[0];
    h->opt = o; h->opt.world_size = 1; h->opt.rank = 0;
    h->nshell = M->parts[0]->nshell; h->nbf = M->parts[0]->nbf; h->ncart = M->parts[0]->ncart;
    h->diag_ready = true;
    M->n2 = (size_t)h->nbf * h->nbf;
    if (cudaMallocHost(&M->pin, sizeof(double) * 7 * M->n2) != cudaSuccess) return fail("cf_create_multi: cudaMallocHost failed");
    multi_refresh_stats(h);
    return h;
}

extern "C" int cf_num_devices(const cf_handle* h) { return !h ? 0 : h->multi ? h->multi->ndev : 1; }

static int multi_fail(cf_handle* h, int rc) {
    for (int i = 0; i < h->multi->ndev; i++)
        if (h->multi->rc[i] != CF_OK) { set_error(h, "device " + std::to_string(h->multi->devices[i]) + ": " + h->multi->parts[i]->err + h->multi->errs[i]); break; }
    return rc;
}

static int multi_build_jk(cf_handle* h, int nbf, const double* Dd, const double* Da, const double* Db, double exx,
                          double* J, double* Kd, double* Ka, double* Kb) {
    MultiCtx* M = h->multi;
    const size_t n2 = M->n2, bytes = sizeof(double) * n2;
    const double t0 = now_s();
    if (h->opt.verbose > 0) std::printf("Contracting 4c-2e repulsion integrals with 1 matrix ... ");
    const double* src[3] = {Dd, Da, Db};
    double* outs[4] = {J, Kd, Ka, Kb};
    for (int k = 0; k < 3; k++) if (src[k]) std::memcpy(M->pin + k * n2, src[k], bytes);
    const double* dk[3]; int slot[3];
    const int nk = exchange_list(Dd, Da, Db, exx, dk, slot);
    int rc = M->run([&](int i) -> int {
        cf_handle* p = M->parts[i];
        DeviceGuard g(p->device);
        cudaStream_t s = M->streams[i];
        M->errs[i].clear();
        auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess) { M->errs[i] = std::string(what) + ": " + cudaGetErrorString(e); return false; } return true; };
        const double* dev[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < 3; k++)
            if (src[k]) { if (!cu(cudaMemcpyAsync(p->d_Dpure[k].p, M->pin + k * n2, bytes, cudaMemcpyHostToDevice, s), "H2D")) return CF_ERR_CUDA; dev[k] = p->d_Dpure[k].p; }
        int r = cf_accumulate_device(p, nbf, dev[0], dev[1], dev[2], exx, (int64_t*)p->d_acc.p, s);
        if (r != CF_OK) return r;
        ncclResult_t nr = M->nccl.AllReduce(p->d_acc.p, p->d_acc.p, cf_acc_reduce_len(p, nk), ncclInt64, ncclSum, M->comms[i], s);
        if (nr != ncclSuccess) { M->errs[i] = std::string("ncclAllReduce: ") + M->nccl.GetErrorString(nr); return CF_ERR_CUDA; }
        if (i == 0) {   // every device holds the summed accumulators; device 0 turns them into J/K for the host
            r = cf_finalize_device(p, nbf, (const int64_t*)p->d_acc.p, exx, Dd != nullptr, Da != nullptr, Db != nullptr,
                                   p->d_out[0].p, p->d_out[1].p, p->d_out[2].p, p->d_out[3].p, s);
            if (r != CF_OK) return r;
            if (!cu(cudaMemcpyAsync(M->pin + 3 * n2, p->d_out[0].p, bytes, cudaMemcpyDeviceToHost, s), "D2H")) return CF_ERR_CUDA;
            for (int k = 0; k < 3; k++)
                if (src[k] && !cu(cudaMemcpyAsync(M->pin + (4 + k) * n2, p->d_out[1 + k].p, bytes, cudaMemcpyDeviceToHost, s), "D2H")) return CF_ERR_CUDA;
        } else {
            if (!cu(cudaEventRecord(p->ev[3], s), "event")) return CF_ERR_CUDA;
        }
        r = fetch_build_info(p, s);
        if (r != CF_OK) return r;
        if (!cu(cudaStreamSynchronize(s), "synchronize")) return CF_ERR_CUDA;
        fetch_times(p);
        return check_scales(p);
    });
    if (rc != CF_OK) return multi_fail(h, rc);
    std::memcpy(J, M->pin + 3 * n2, bytes);
    for (int k = 0; k < 3; k++) if (src[k]) std::memcpy(outs[1 + k], M->pin + (4 + k) * n2, bytes);
    multi_refresh_stats(h);
    if (h->opt.verbose > 0) std::printf("Done in %f s\n", now_s() - t0);
    return CF_OK;
}

static int multi_build_g(cf_handle* h, int nbf, int nmat, const double* Ds, double exx, double* Gs) {
    MultiCtx* M = h->multi;
    const size_t n2 = M->n2, bytes = sizeof(double) * n2;
    for (int k0 = 0; k0 < nmat; k0 += 3) {
        const int nb = std::min(3, nmat - k0);
        for (int k = 0; k < nb; k++) std::memcpy(M->pin + k * n2, Ds + (size_t)(k0 + k) * n2, bytes);
        int rc = M->run([&](int i) -> int {
            cf_handle* p = M->parts[i];
            DeviceGuard g(p->device);
            cudaStream_t s = M->streams[i];
            M->errs[i].clear();
            auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess) { M->errs[i] = std::string(what) + ": " + cudaGetErrorString(e); return false; } return true; };
            for (int k = 0; k < nb; k++) if (!cu(cudaMemcpyAsync(p->d_Dpure[k].p, M->pin + k * n2, bytes, cudaMemcpyHostToDevice, s), "H2D")) return CF_ERR_CUDA;
            int r = build_g_accumulate(p, nb, exx, s);
            if (r != CF_OK) return r;
            ncclResult_t nr = M->nccl.AllReduce(p->d_acc.p, p->d_acc.p, 9 * (size_t)p->ncart * p->ncart, ncclInt64, ncclSum, M->comms[i], s);
            if (nr != ncclSuccess) { M->errs[i] = std::string("ncclAllReduce: ") + M->nccl.GetErrorString(nr); return CF_ERR_CUDA; }
            if (i == 0) {
                r = build_g_finalize(p, nb, exx, s);
                if (r != CF_OK) return r;
                for (int k = 0; k < nb; k++) if (!cu(cudaMemcpyAsync(M->pin + (3 + k) * n2, p->d_out[k].p, bytes, cudaMemcpyDeviceToHost, s), "D2H")) return CF_ERR_CUDA;
            } else if (!cu(cudaEventRecord(p->ev[3], s), "event")) return CF_ERR_CUDA;
            r = fetch_build_info(p, s);
            if (r != CF_OK) return r;
            if (!cu(cudaStreamSynchronize(s), "synchronize")) return CF_ERR_CUDA;
            fetch_times(p);
            return check_scales(p);
        });
        if (rc != CF_OK) return multi_fail(h, rc);
        for (int k = 0; k < nb; k++) std::memcpy(Gs + (size_t)(k0 + k) * n2, M->pin + (3 + k) * n2, bytes);
    }
    multi_refresh_stats(h);
    return CF_OK;
}

// every partition contracts its share of the quartets; the 3*natom partial vectors are added in device order (fixed)
static int multi_contract_hess(cf_handle* h, int nbf, const double* D, double exx, int natom, double* hess) {
    MultiCtx* M = h->multi;
    const size_t nh2 = 9 * (size_t)natom * natom;
    std::vector<std::vector<double>> part(M->ndev, std::vector<double>(nh2, 0.0));
    int rc = M->run([&](int i) -> int { return cf_contract_hess(M->parts[i], nbf, D, exx, natom, part[i].data()); });
    if (rc != CF_OK) return multi_fail(h, rc);
    for (size_t j = 0; j < nh2; j++) { double s_ = 0.0; for (int i = 0; i < M->ndev; i++) s_ += part[i][j]; hess[j] = s_; }
    multi_refresh_stats(h);
    return CF_OK;
}

static int multi_contract_grads(cf_handle* h, int nbf, const double* D1, const double* D2, double exx, int natom, double* grad) {
    MultiCtx* M = h->multi;
    std::vector<std::vector<double>> part(M->ndev, std::vector<double>(3 * (size_t)natom, 0.0));
    int rc = M->run([&](int i) -> int { return cf_contract_grads(M->parts[i], nbf, D1, D2, exx, natom, part[i].data()); });
    if (rc != CF_OK) return multi_fail(h, rc);
    for (int j = 0; j < 3 * natom; j++) { double s = 0.0; for (int i = 0; i < M->ndev; i++) s += part[i][j]; grad[j] = s; }
    multi_refresh_stats(h);
    return CF_OK;
}
