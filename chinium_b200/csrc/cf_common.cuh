// cf_common.cuh -- device-side data layout shared by all kernels of the J/K engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define CF_LMAX_DEV 3            // highest shell l the instantiated kernels cover (f); tables go to 9 roots
#define CF_NCLS ((CF_LMAX_DEV + 1) * (CF_LMAX_DEV + 2) / 2)   // pair classes (la>=lb)
#define CF_PSTRIDE 32            // primitive pairs of 32 consecutive shell pairs are interleaved (coalesced per warp)

// Boys-function table F_m(T_i), T_i = i * BOYS_DT, i = 0..BOYS_NROW-1, m = 0..BOYS_NCOL-1 (row-major)
#define BOYS_DT 0.125
#define BOYS_NROW 289            // T in [0, 36]
#define BOYS_NCOL 12
#define BOYS_TMAX 36.0

__host__ __device__ constexpr int cf_ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int cf_pair_class(int la, int lb) { return la * (la + 1) / 2 + lb; }  // la>=lb

// Shell pairs of one class (la,lb), sorted by (primitive count desc, Schwarz bound desc).  SoA, device pointers.
// Primitive pair i of shell pair n lives in slot pbase[n] + i*CF_PSTRIDE of the primitive arrays: pairs are
// grouped in blocks of 32 and their primitives interleaved, so a warp whose lanes hold 32 consecutive pairs reads
// primitive i of all of them with one coalesced request.  Padding slots have c = 0.
struct PairClassDev {
    int npair;
    const int* sa;          // [npair] shell with the larger l (ties: larger index)
    const int* sb;
    const int* cao_a;       // [npair] first Cartesian AO of shell a / b
    const int* cao_b;
    const int* pbase;       // [npair] slot of the first primitive pair
    const int* nprim;       // [npair] surviving primitive pairs
    const double* A;        // [npair*3] centre of a
    const double* AB;       // [npair*3] A - B
    const double* Q;        // [npair] Schwarz bound max_ij sqrt((ij|ij)) over the shell pair's basis functions
    // primitive pairs (slots)
    const double* p;        // exponent sum
    const double* hp;       // 0.5 / p
    const double* Px;       // product centre
    const double* Py;
    const double* Pz;
    const double* c;        // ca*cb*exp(-ab/p |AB|^2) * sqrt(2) pi^(5/4) / p
};

// Work counters of one build, per class-pair task: CF_CNT_SLOTS words of evaluated shell quartets followed by
// CF_CNT_SLOTS words of EXECUTED primitive quartets (those that passed the primitive cutoff and were really run
// through roots + recurrences + assembly).  Every warp keeps its counts in registers over all its work items and
// flushes them once at the end of the kernel; the slots spread the flushes (same-address atomics serialise in L2).
#define CF_CNT_SLOTS 32
#define CF_CNT_WORDS (2 * CF_CNT_SLOTS)
__device__ __forceinline__ int cf_cnt_slot() { return (int)((blockIdx.x * 4u + (threadIdx.x >> 5)) & (CF_CNT_SLOTS - 1)); }
// warp-level flush (all 32 lanes call it): nq / np are per-lane counts
__device__ __forceinline__ void cf_cnt_flush(unsigned long long* cnt, unsigned nq, unsigned np) {
    if (!cnt) return;
    nq = __reduce_add_sync(0xffffffffu, nq);
    np = __reduce_add_sync(0xffffffffu, np);
    if ((threadIdx.x & 31) == 0) {
        if (nq) atomicAdd(cnt + cf_cnt_slot(), (unsigned long long)nq);
        if (np) atomicAdd(cnt + CF_CNT_SLOTS + cf_cnt_slot(), (unsigned long long)np);
    }
}

struct RysTablesDev {
    const double* table;
    const double* asym;
    const double* boys;     // [BOYS_NROW][BOYS_NCOL]
};

// One launch = one (bra class, ket class) rectangle/triangle of the pair x pair grid.
struct QuartetTask {
    PairClassDev bra, ket;
    const long long* qoff;  // [bra.npair+1] prefix sum of ket counts per bra pair (generic kernels)
    long long nquartet;     // qoff[bra.npair]
    int chunk;              // quartets per work chunk (generic kernels)
    int rank, world;
    int same_class;         // bra class == ket class (triangular, diagonal weight 1/2)
    int ncart;              // leading dimension of the Cartesian matrices
    int nk;                 // number of exchange densities (0..3)
    const double* Dtot;     // [ncart*ncart] Cartesian total density (2Dd+Da+Db), symmetric
    int nj;                 // Coulomb densities digested by this launch: 1 for the J/K build (Dj[0] == Dtot, accJm[0] == accJ), up to 3
    const double* Dj[3];    //   for the multi-density build (Int4C2E.cpp:685-745), where J_k and K_k share the integrals
    long long* accJm[3];
    const double* Dk[3];    // Cartesian exchange densities
    long long* accJ;        // [ncart*ncart] fixed-point raw J
    long long* accK[3];
    const double* scales;   // device: [0] J scale, [1] K scale (powers of two, written by scales_kernel of this build),
                            //   [4] effective Schwarz threshold, [6] != 0: the J adds also feed the low limb
    long long jlo_off;      // words from a J accumulator (accJm[x]) to its LOW LIMB: contributions are rounded to the grid
                            //   1/scaleJ of the high limb and the rounding residual, scaled by 2^31, goes to the low limb.
                            //   Used when scales[6] != 0 (large systems, where ~1e6 roundings per element would add up to
                            //   the 1e-10 bar); integer adds in both limbs, so the result stays bit-identical for any schedule
    double* store;          // STORE mode: Cartesian blocks, NOUT doubles per quartet in flat order
    int diag;               // Schwarz mode: quartet q is (pair q | pair q)
    RysTablesDev rys;
    double prim_cut;        // skip primitive quartets with |c_ab c_cd| below this
    double thr;             // Cauchy-Schwarz threshold on Q_ab * Q_cd (<= 0: none), reference Int4C2E.cpp:108-113
    // work items of the thread-per-quartet / sliced / warp-group kernels: item = (x: bra pair, y: first ket pair,
    // z: number of consecutive ket pairs <= NQ of the kernel).  The list is built on the device at setup from the
    // Schwarz bounds (only ket runs that can pass Q_ab * Q_cd > thr are listed; triangular limit folded in).
    const int4* items;
    long long nitem;
    // bra-loop kernels (eri_tpqa.cuh): bra pairs (positions in the bra class arrays) grouped by shell a, Schwarz bound
    // descending inside a group; item = (x: first entry of `border`, y: first ket pair of an aligned block of 32,
    // z: ket pairs in the block, w: bra pairs in the chunk)
    int braloop;            // thread-per-quartet classes: 1 = bra-loop kernel + chunk items (low contraction), 0 = one bra pair per item
    const int* border;
    // bra-role copy in `border` order: record e = (x: position in the class arrays, y: shell b | primitive count << 16,
    // z: first Cartesian AO of b, w: offset of its CONTIGUOUS primitives) and (Q, ABx, ABy, ABz); primitive value v of
    // primitive k lives at bprim[v * bprim_stride + w + k], v = p, 1/(2p), Px, Py, Pz, c
    const int4* brec_i;
    const double* brec_d;
    const double* bprim;
    long long bprim_stride;
    unsigned long long* cnt;       // [CF_CNT_WORDS] this task's work counters of the build (null: Schwarz/STORE mode)
};
