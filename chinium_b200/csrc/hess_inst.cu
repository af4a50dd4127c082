// hess_inst.cu -- instantiates the second-derivative (nuclear Hessian) kernel of eri_hess.cuh for ONE bra pair class
// (-DCF_BRA=0..9) against every ket class <= bra; its own translation units so the J/K units keep their build time.
#include <cstdlib>
#include "eri_hess.cuh"

#ifndef CF_BRA
#error "compile with -DCF_BRA=<bra class index>"
#endif

template <int C> struct HClassL { static constexpr int a = C >= 6 ? 3 : C >= 3 ? 2 : C >= 1 ? 1 : 0, b = C - a * (a + 1) / 2; };

constexpr int hess_group_size(int nout) { return nout <= 1600 ? 64 : nout <= 3600 ? 128 : 256; }

template <int BRA, int KET>
static cudaError_t launch_hess_pair(const GradTask& t, int grid, cudaStream_t s, int* g_out, size_t* smem_out) {
    constexpr int LA = HClassL<BRA>::a, LB = HClassL<BRA>::b, LC = HClassL<KET>::a, LD = HClassL<KET>::b;
    constexpr int NOUT = cf_ncart(LA) * cf_ncart(LB) * cf_ncart(LC) * cf_ncart(LD);
    constexpr int G = hess_group_size(NOUT);
    const size_t smem = eri_hess_smem<LA, LB, LC, LD>(G);
    if (g_out) *g_out = G;
    if (smem_out) *smem_out = smem;
    if (grid <= 0) return cudaSuccess;
    auto k = eri_hess_generic<LA, LB, LC, LD, G>;
    if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    k<<<grid, G, smem, s>>>(t);
    return cudaGetLastError();
}
template <int BRA, int KET>
struct HessDispatch {
    static cudaError_t go(int ket, const GradTask& t, int grid, cudaStream_t s, int* g, size_t* sm) {
        if (ket == KET) return launch_hess_pair<BRA, KET>(t, grid, s, g, sm);
        return HessDispatch<BRA, KET - 1>::go(ket, t, grid, s, g, sm);
    }
};
template <int BRA>
struct HessDispatch<BRA, -1> {
    static cudaError_t go(int, const GradTask&, int, cudaStream_t, int*, size_t*) { return cudaErrorInvalidValue; }
};

#define CF_HCAT2(a, b) a##b
#define CF_HCAT(a, b) CF_HCAT2(a, b)
cudaError_t CF_HCAT(cf_launch_hess_bra, CF_BRA)(int ket, const GradTask& t, int grid, cudaStream_t s, int* g, size_t* sm) {
    return HessDispatch<CF_BRA, CF_BRA>::go(ket, t, grid, s, g, sm);
}
