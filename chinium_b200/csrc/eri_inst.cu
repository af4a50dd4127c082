// eri_inst.cu -- instantiates the generic quartet kernel for ONE bra pair class (-DCF_BRA=0..9) against
// every ket class <= bra.  Compiled once per bra class so the 55 class pairs build in parallel.
#include <cstdlib>
#include <type_traits>
#include "eri_generic.cuh"
#include "eri_tpq.cuh"
#include "eri_tpqa.cuh"
#include "eri_wg.cuh"
#include "eri_grad.cuh"

// developer knob for A/B measurements of the warp-group configurations (see wg_cfg in eri_wg.cuh)
static int cf_wg_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("CF_WG_VARIANT"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
    return v;
}

template <int B, int E, class F>
static inline void host_static_for(F&& f) {
    if constexpr (B < E) { f(std::integral_constant<int, B>{}); host_static_for<B + 1, E>(f); }
}

#ifndef CF_BRA
#error "compile with -DCF_BRA=<bra class index>"
#endif

template <int C> struct ClassL;   // class index -> (la, lb)
template <> struct ClassL<0> { static constexpr int a = 0, b = 0; };
template <> struct ClassL<1> { static constexpr int a = 1, b = 0; };
template <> struct ClassL<2> { static constexpr int a = 1, b = 1; };
template <> struct ClassL<3> { static constexpr int a = 2, b = 0; };
template <> struct ClassL<4> { static constexpr int a = 2, b = 1; };
template <> struct ClassL<5> { static constexpr int a = 2, b = 2; };
template <> struct ClassL<6> { static constexpr int a = 3, b = 0; };
template <> struct ClassL<7> { static constexpr int a = 3, b = 1; };
template <> struct ClassL<8> { static constexpr int a = 3, b = 2; };
template <> struct ClassL<9> { static constexpr int a = 3, b = 3; };

constexpr int group_size(int nout) { return nout <= 640 ? 32 : nout <= 1600 ? 64 : nout <= 3600 ? 128 : 256; }

template <int BRA, int KET>
static cudaError_t launch_pair(const QuartetTask& t, int store, int grid, cudaStream_t s, int* g_out, size_t* smem_out, int* kind_out) {
    constexpr int LA = ClassL<BRA>::a, LB = ClassL<BRA>::b, LC = ClassL<KET>::a, LD = ClassL<KET>::b;
    constexpr int NOUT = cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD);
    constexpr int NROOTS = (LA + LB + LC + LD) / 2 + 1;
    cudaError_t e;
    if constexpr (tpq_ok(LA, LB, LC, LD)) {
        if (!store) {   // thread-per-quartet family (bra-loop kernel): grid counts warp-private work items
            constexpr size_t smem = tpq_smem(NROOTS);
            if (g_out) *g_out = TPQ_THREADS;
            if (smem_out) *smem_out = smem;
            if (kind_out) *kind_out = 1 + 16 * 32;    // items: bra chunk x aligned block of 32 kets, one per warp
            if (grid <= 0) return cudaSuccess;
            // bra-loop kernel; the register accumulators are sized by the number of exchange densities
            auto go = [&](auto k) -> cudaError_t {
                if (smem > 48 * 1024) { cudaError_t e2 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e2 != cudaSuccess) return e2; }
                k<<<grid, TPQ_THREADS, smem, s>>>(t);
                return cudaSuccess;
            };
            // low contraction (digestion/atomics-bound): bra-loop kernel; high contraction (FP64-bound): the leaner
            // one-bra-pair-per-item kernel.  The engine decides per task (QuartetTask::braloop) and builds the matching items.
            if (!t.braloop) e = go(eri_jk_tpq<LA, LB, LC, LD>);
            else if (t.nj > 1) e = go(eri_jk_tpqa<LA, LB, LC, LD, 3, 3>);      // multi-density build
            else if (t.nk <= 1) e = go(eri_jk_tpqa<LA, LB, LC, LD, 1>);
            else if (t.nk == 2) e = go(eri_jk_tpqa<LA, LB, LC, LD, 2>);
            else e = go(eri_jk_tpqa<LA, LB, LC, LD, 3>);
            if (e != cudaSuccess) return e;
            return cudaGetLastError();
        }
    }
    if constexpr (wg_cfg(BRA, KET) != 0) {
        if (!store) {   // warp-group cooperative family; g_wg_variant picks among the compiled alternatives
            bool done = false;
            cudaError_t err = cudaSuccess;
            host_static_for<0, CF_WG_NVAR>([&](auto vtag) {
                constexpr int V = decltype(vtag)::value;
                constexpr int CFG = wg_cfg(BRA, KET, V);
                // identical configurations share one instantiation; the variant index only selects
                if (done || (cf_wg_variant() % CF_WG_NVAR) != V) return;
                done = true;
                constexpr int MK = CFG & 255;
                constexpr bool SWAP = ((CFG >> 8) & 1) != 0;
                constexpr int HS = ((CFG >> 12) & 15) ? ((CFG >> 12) & 15) : 1;
                constexpr int MINB = ((CFG >> 16) & 15) ? ((CFG >> 16) & 15) : 2;
                using C = typename std::conditional<SWAP, WgCfg<LC, LD, LA, LB, MK, HS, MINB>, WgCfg<LA, LB, LC, LD, MK, HS, MINB>>::type;
                if (g_out) *g_out = 32 * WG_WARPS;
                if (smem_out) *smem_out = C::SMEM;
                if (kind_out) *kind_out = 3 + (SWAP ? 8 : 0) + 16 * C::NQ;
                if (grid <= 0) return;
                if constexpr (SWAP) {
                    auto k = eri_jk_wg<LC, LD, LA, LB, MK, HS, MINB>;
                    err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM); if (err != cudaSuccess) return;
                    k<<<grid, 32 * WG_WARPS, C::SMEM, s>>>(t);
                } else {
                    auto k = eri_jk_wg<LA, LB, LC, LD, MK, HS, MINB>;
                    err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM); if (err != cudaSuccess) return;
                    k<<<grid, 32 * WG_WARPS, C::SMEM, s>>>(t);
                }
                err = cudaGetLastError();
            });
            return err;
        }
    }
    if constexpr (tpqs_gs(LA, LB, LC, LD) > 0) {
        if (!store) {   // sliced thread-per-quartet family: GS threads per quartet, work item = bra pair x NQ kets
            constexpr int GS = tpqs_gs(LA, LB, LC, LD);
            constexpr size_t smem = tpqs_smem(NROOTS, GS);
            if (g_out) *g_out = GS * tpqs_nq(GS);
            if (smem_out) *smem_out = smem;
            if (kind_out) *kind_out = 2 + 16 * tpqs_nq(GS);
            if (grid <= 0) return cudaSuccess;
            auto k = eri_jk_tpqs<LA, LB, LC, LD, GS>;
            if (smem > 48 * 1024) { e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
            k<<<grid, GS * tpqs_nq(GS), smem, s>>>(t);
            return cudaGetLastError();
        }
    }
    constexpr int G = group_size(NOUT);
    const size_t smem = eri_generic_smem<LA, LB, LC, LD>(store ? 0 : t.nk);
    if (g_out) *g_out = G;
    if (smem_out) *smem_out = smem;
    if (kind_out) *kind_out = 0;
    if (grid <= 0) return cudaSuccess;   // query only
    if (store) {
        auto k = eri_jk_generic<LA, LB, LC, LD, G, true>;
        if (smem > 48 * 1024) { e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
        k<<<grid, G, smem, s>>>(t);
    } else {
        auto k = eri_jk_generic<LA, LB, LC, LD, G, false>;
        if (smem > 48 * 1024) { e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
        k<<<grid, G, smem, s>>>(t);
    }
    return cudaGetLastError();
}

template <int BRA, int KET>
struct Dispatch {
    static cudaError_t go(int ket, const QuartetTask& t, int store, int grid, cudaStream_t s, int* g, size_t* sm, int* kind) {
        if (ket == KET) return launch_pair<BRA, KET>(t, store, grid, s, g, sm, kind);
        return Dispatch<BRA, KET - 1>::go(ket, t, store, grid, s, g, sm, kind);
    }
};
template <int BRA>
struct Dispatch<BRA, -1> {
    static cudaError_t go(int, const QuartetTask&, int, int, cudaStream_t, int*, size_t*, int*) { return cudaErrorInvalidValue; }
};

#define CF_CAT2(a, b) a##b
#define CF_CAT(a, b) CF_CAT2(a, b)
// cf_launch_bra<N>(ket_class, task, store, grid, stream, &G, &smem, &kind)   kind: 0 generic (grid = CTAs over quartet chunks), 1 thread-per-quartet
cudaError_t CF_CAT(cf_launch_bra, CF_BRA)(int ket, const QuartetTask& t, int store, int grid, cudaStream_t s, int* g, size_t* sm, int* kind) {
    return Dispatch<CF_BRA, CF_BRA>::go(ket, t, store, grid, s, g, sm, kind);
}

// ---- nuclear-gradient kernels (eri_grad.cuh): one generic kernel per class pair ------------------------------------
template <int BRA, int KET>
static cudaError_t launch_grad_pair(const GradTask& t, int grid, cudaStream_t s, int* g_out, size_t* smem_out) {
    constexpr int LA = ClassL<BRA>::a, LB = ClassL<BRA>::b, LC = ClassL<KET>::a, LD = ClassL<KET>::b;
    constexpr int NOUT = cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD);
    if constexpr (grad_tpq_ok(LA, LB, LC, LD)) {
        // thread per quartet, per-warp shared-memory rows of the gradient (needs 4 * ngrad doubles); the larger classes only
        // when the task has enough quartets to fill the machine (CF_GRAD_TPQ_MINQ: developer/test override)
        constexpr int NRG = (LA + LB + LC + LD + 1) / 2 + 1;
        const size_t smem_t = sizeof(double) * (size_t)(tpq_table_len(NRG) + (GRAD_TPQ_THREADS / 32) * t.ngrad);
        static long long minq = -1;
        if (minq < 0) { const char* e = getenv("CF_GRAD_TPQ_MINQ"); minq = e ? atoll(e) : 30000; }
        if (smem_t <= 96 * 1024 && (NOUT <= 9 || t.nquartet >= minq)) {
            if (g_out) *g_out = -GRAD_TPQ_THREADS;          // negative: quartets per CTA block, thread-per-quartet enumeration
            if (smem_out) *smem_out = smem_t;
            if (grid <= 0) return cudaSuccess;
            auto k = eri_grad_tpq<LA, LB, LC, LD>;
            if (smem_t > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t); if (e != cudaSuccess) return e; }
            k<<<grid, GRAD_TPQ_THREADS, smem_t, s>>>(t);
            return cudaGetLastError();
        }
    }
    constexpr int G = group_size(NOUT) < 64 ? 64 : group_size(NOUT);
    const size_t smem = eri_grad_smem<LA, LB, LC, LD>(G);
    if (g_out) *g_out = G;
    if (smem_out) *smem_out = smem;
    if (grid <= 0) return cudaSuccess;
    auto k = eri_grad_generic<LA, LB, LC, LD, G>;
    if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    k<<<grid, G, smem, s>>>(t);
    return cudaGetLastError();
}
template <int BRA, int KET>
struct GradDispatch {
    static cudaError_t go(int ket, const GradTask& t, int grid, cudaStream_t s, int* g, size_t* sm) {
        if (ket == KET) return launch_grad_pair<BRA, KET>(t, grid, s, g, sm);
        return GradDispatch<BRA, KET - 1>::go(ket, t, grid, s, g, sm);
    }
};
template <int BRA>
struct GradDispatch<BRA, -1> {
    static cudaError_t go(int, const GradTask&, int, cudaStream_t, int*, size_t*) { return cudaErrorInvalidValue; }
};
cudaError_t CF_CAT(cf_launch_grad_bra, CF_BRA)(int ket, const GradTask& t, int grid, cudaStream_t s, int* g, size_t* sm) {
    return GradDispatch<CF_BRA, CF_BRA>::go(ket, t, grid, s, g, sm);
}

// ---- matrix-form gradient kernels (Int4C2E::ContractGrads(D), eri_grad.cuh) -----------------------------------------
template <int BRA, int KET>
static cudaError_t launch_gradmat_pair(const GradTask& t, int grid, cudaStream_t s, int* g_out, size_t* smem_out) {
    constexpr int LA = ClassL<BRA>::a, LB = ClassL<BRA>::b, LC = ClassL<KET>::a, LD = ClassL<KET>::b;
    constexpr int NOUT = cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD);
    constexpr int G = group_size(NOUT) < 64 ? 64 : group_size(NOUT);
    const size_t smem = eri_gradmat_smem<LA, LB, LC, LD>(G);
    if (g_out) *g_out = G;
    if (smem_out) *smem_out = smem;
    if (grid <= 0) return cudaSuccess;
    auto k = eri_gradmat_generic<LA, LB, LC, LD, G>;
    if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    k<<<grid, G, smem, s>>>(t);
    return cudaGetLastError();
}
template <int BRA, int KET>
struct GradMatDispatch {
    static cudaError_t go(int ket, const GradTask& t, int grid, cudaStream_t s, int* g, size_t* sm) {
        if (ket == KET) return launch_gradmat_pair<BRA, KET>(t, grid, s, g, sm);
        return GradMatDispatch<BRA, KET - 1>::go(ket, t, grid, s, g, sm);
    }
};
template <int BRA>
struct GradMatDispatch<BRA, -1> {
    static cudaError_t go(int, const GradTask&, int, cudaStream_t, int*, size_t*) { return cudaErrorInvalidValue; }
};
cudaError_t CF_CAT(cf_launch_gradmat_bra, CF_BRA)(int ket, const GradTask& t, int grid, cudaStream_t s, int* g, size_t* sm) {
    return GradMatDispatch<CF_BRA, CF_BRA>::go(ket, t, grid, s, g, sm);
}
