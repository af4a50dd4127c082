/* chinium_fock.h -- C ABI of the B200-native direct-SCF J/K (Fock) build engine.
 *
 * Drop-in boundary for ONE path of FreemanTheMaverick/Chinium: the four-centre
 * two-electron-integral J/K contraction behind `class Int4C2E`
 * (reference src/Integral/Int4C2E.h:10-50), called once per SCF iteration by
 * src/HartreeFockKohnSham (e.g. Restricted/SP.cpp:47, Unrestricted/SP.cpp:58,
 * Universal.cpp:38).  The reference has no FFI; a maintainer binds these entry
 * points from the C++ adaptor `chinium_b200/cpp/Int4C2E_b200.hpp`, which keeps the
 * reference's method names.
 *
 * Conventions (all from the reference):
 *   - matrices are column-major FP64 nbf x nbf, ld = nbf (Eigen::MatrixXd,
 *     src/Macro/Abbreviation.h:4); density inputs are symmetric.
 *   - shell `type`: 0 = S, +1 = Cartesian P (x,y,z), -l = pure shell of angular
 *     momentum l in the order m = -l..+l of Racah-normalised real solid harmonics
 *     (src/Integral/Macro.h:5-8, src/Grid/AO/Pure*.hpp).
 *   - `coefs_normalized` are MwfnShell.NormalizedCoefficients
 *     (src/Integral/Normalization.cpp:11-18); coordinates in bohr.
 *   - J_ij = sum_kl (ij|kl) (2 Dd + Da + Db)_kl ;  KX_ik = exx * sum_jl (ij|kl) DX_jl
 *     i.e. exactly what Gunified returns (src/Integral/Int4C2E.cpp:601-671).
 *
 * There is NO CPU fallback: every entry point that computes returns
 * CF_ERR_NO_DEVICE when no sm_100 GPU is usable.
 */
#ifndef CHINIUM_FOCK_H
#define CHINIUM_FOCK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CF_OK 0
#define CF_ERR_NO_DEVICE 1      /* no CUDA device / not sm_100              */
#define CF_ERR_BAD_ARGUMENT 2   /* null pointer, nbf mismatch, bad type      */
#define CF_ERR_UNSUPPORTED 3    /* angular momentum above CF_MAX_L, ...      */
#define CF_ERR_CUDA 4           /* a CUDA runtime call failed                */
#define CF_ERR_RANGE 5          /* fixed-point accumulator range exceeded    */
#define CF_ERR_STATE 6          /* calls out of order (reference asserts: Int4C2E.cpp:514,538,555) */

#define CF_MAX_L 3              /* highest shell angular momentum the device kernels accept (f) */

/* Flat basis description; replaces __Make_Basis_Set__ (src/Integral/Macro.h:1-25).
 * All pointers are HOST pointers, borrowed only for the duration of cf_create. */
typedef struct cf_basis {
    int nshell;
    const int* type;              /* [nshell] 0, +1, -1, -2, -3, ...                       */
    const int* nprim;             /* [nshell]                                              */
    const int* prim_offset;       /* [nshell] offset of the shell's first primitive        */
    const double* exps;           /* [sum nprim]                                           */
    const double* coefs_normalized; /* [sum nprim]                                         */
    const double* center_xyz;     /* [3*nshell] bohr                                       */
    const int* shell2atom;        /* [nshell] (may be NULL; kept for the derivative rows)  */
} cf_basis;

/* Options; zero-initialise then override.  Replaces the hard-coded
 * Int4C2E(mwfn, exx, threshold) arguments (src/HartreeFockKohnSham/SelfConsistentField.cpp:47). */
typedef struct cf_options {
    double threshold;   /* Cauchy-Schwarz threshold on sqrt((ab|ab)(cd|cd)) as in Int4C2E.cpp:108-113.
                           <= 0 (the reference passes -1): no Schwarz screening; only shell pairs whose
                           largest primitive overlap prefactor is < pair_cutoff are dropped.          */
    double pair_cutoff; /* primitive pairs whose overlap prefactor is below this are dropped at setup; 0 -> default
                           1e-20 (cannot change J/K at the 1e-12 level).  Shell pairs left without primitives are
                           not evaluated; the reference-style counts (ref_*) are NOT affected by it                */
    int device;         /* CUDA device ordinal; -1 -> current device                                  */
    int rank;           /* this handle computes partition `rank` of `world_size` (static, cost-       */
    int world_size;     /*  balanced split of the quartet work; 0/0 or 0/1 = everything)              */
    int verbose;        /* >0: print the reference's "... Done in %f s" lines (Int4C2E.cpp:500-587)   */
    int j_two_limb;     /* J accumulator low limb: 0 = automatic (switched on per build when the estimated rounding
                           noise of ~1e6 fixed-point adds per element exceeds 1e-11), 1 = always, -1 = never */
    int reserved[7];
} cf_options;

typedef struct cf_handle cf_handle;

/* Counters mirrored from the reference's screening printout (Int4C2E.cpp:516-531) plus work metrics. */
typedef struct cf_stats {
    int nshell, nbf, ncart;            /* ncart: Cartesian working dimension on the device            */
    int64_t shell_pairs_total;         /* nshell(nshell+1)/2                                         */
    int64_t shell_pairs_kept;
    int64_t canonical_quartets;        /* surviving canonical shell quartets, whole job             */
    int64_t canonical_quartets_local;  /* ... in this handle's partition                            */
    int64_t unique_integrals;          /* the reference's RepulsionLength for the surviving quartets */
    int64_t primitive_quartets;
    double  flops_alg_jk[4];           /* F_alg (SURVEY 8d model) for nK = 0,1,2,3, this partition    */
    int     n_launches_last;           /* kernels launched by the last cf_build_* call               */
    double  ms_device_last;            /* CUDA-event time of the last build (all kernels)            */
    double  ms_eri_last;               /* ... of the ERI/digestion kernels only                      */
    double  fixedpoint_scale_log2[2];  /* log2 of the J and K accumulator scales of the last build   */
    double  threshold_effective_last;  /* Schwarz threshold the last build applied: max(threshold, density_threshold / max|D|) */
    int64_t quartets_evaluated_last;   /* shell quartets this partition actually evaluated in the last build            */
    double  flops_alg_grad;            /* F_alg of one cf_contract_grads call (DESIGN.md model), this partition          */
    double  ms_grad_last;              /* CUDA-event time of the gradient kernels of the last cf_contract_grads call     */
    int64_t primitive_quartets_executed_last; /* primitive quartets the last build really ran (device counters): the ones
                                                 that survive pair_cutoff at setup AND the in-kernel primitive cutoff   */
    double  flops_executed_last;       /* SURVEY 8d model on EXECUTED work: sum over class pairs of executed primitive
                                          quartets x per-primitive flops + digestion term of the evaluated quartets     */
    int     j_two_limb_last;           /* 1: the last build accumulated J in two limbs                                   */
    double  j_rounding_estimate_last;  /* the estimate the automatic choice was based on (absolute, J units)            */
    int64_t ref_repulsion_length;      /* the reference's RepulsionLength / ShellQuartetLength (Int4C2E.cpp:79-128):     */
    int64_t ref_shell_quartet_length;  /*   its own loop nest and uniqueness predicate, independent of pair_cutoff        */
} cf_stats;

/* cf_create: pair build + Schwarz bounds + class sort + task lists, all on the device.
 * Replaces getRepulsionDiag / getRepulsionLength / getRepulsionIndices / getThreadPointers /
 * CalculateIntegrals(0) (SelfConsistentField.cpp:49-53). Returns NULL on failure; the reason is
 * available from cf_last_error(NULL). */
cf_handle* cf_create(const cf_basis* basis, const cf_options* opts);
void cf_destroy(cf_handle* h);

/* Multi-GPU handle: ONE process drives `ndev` GPUs of the box (devices[i], or 0..ndev-1 when devices == NULL; ndev <= 0 =
 * all GPUs).  The reference is a single C++ process whose ContractInts spawns its workers internally (Int4C2E.cpp:617-621,
 * SURVEY 8b "Threading"); here one worker thread + one stream per device do the same: every device evaluates its static
 * partition of the quartet work, the fixed-point accumulators are summed with ncclAllReduce(ncclInt64, ncclSum) over
 * NVLink, device 0 finalises.  Results are bit-identical to the single-GPU handle.  The HOST calls (cf_build_jk,
 * cf_build_g_multi, cf_contract_grads, cf_get_repulsion_diag, cf_get_stats, cf_set_density_threshold) accept such a
 * handle; the *_device calls do not (they return CF_ERR_BAD_ARGUMENT).  libnccl.so.2 is bound at run time (dlopen);
 * if it cannot be loaded the call fails (NULL) -- there is no fallback path.  opts->rank / world_size are ignored. */
cf_handle* cf_create_multi(const cf_basis* basis, const cf_options* opts, int ndev, const int* devices);
int cf_num_devices(const cf_handle* h);
const char* cf_last_error(const cf_handle* h);
int cf_get_stats(const cf_handle* h, cf_stats* out);
int cf_nbf(const cf_handle* h);

/* Density-weighted screening for direct SCF (the stored-ERI reference has no counterpart):
 * with dthr > 0 a shell quartet is evaluated iff Q_ab Q_cd > threshold (the reference's test, Int4C2E.cpp:108-113)
 * AND Q_ab Q_cd * max|D| > dthr, max|D| being the largest element of the densities of THAT call (derived on the
 * device per build).  Meant for incremental builds G[D_n - D_(n-1)]: the neglected contributions are bounded by
 * dthr per quartet.  dthr <= 0 switches it off (default).  Results stay bit-identical for any number of GPUs. */
int cf_set_density_threshold(cf_handle* h, double dthr);

/* Schwarz diagonal (ab|ab) for all basis-function pairs, nbf x nbf col-major
 * (the reference's Diag1212, Int4C2E.cpp:19-77). */
int cf_get_repulsion_diag(cf_handle* h, double* diag1212);

/* NOTE on partitions: a handle created by cf_create with world_size > 1 evaluates ITS share of the quartets only (one
 * process per GPU, the caller owns the communication): the host calls below then return partial results, and the route
 * is cf_accumulate_device -> integer all-reduce of the accumulator -> cf_finalize_device (chinium_b200/distributed.py),
 * respectively a sum of the partial gradient vectors.  A C++ host that simply wants all GPUs of the box uses
 * cf_create_multi instead and calls cf_build_jk as usual. */

/* The hot call; replaces Int4C2E::ContractInts(Dd,Da,Db,nthreads,output) (Int4C2E.cpp:673-683).
 * HOST pointers. Dd/Da/Db: nullable (absent = the reference's 0x0 matrix). J is always written;
 * KX is written iff DX is given (zeros when exx <= 0, Int4C2E.cpp:638), and may be NULL otherwise. */
int cf_build_jk(cf_handle* h, int nbf,
                const double* Dd, const double* Da, const double* Db, double exx,
                double* J, double* Kd, double* Ka, double* Kb);

/* Same with DEVICE pointers on the handle's device (steady-state SCF keeps D/J/K resident);
 * work is enqueued on `stream` (a cudaStream_t; NULL = default stream) and not synchronised. */
int cf_build_jk_device(cf_handle* h, int nbf,
                       const double* Dd, const double* Da, const double* Db, double exx,
                       double* J, double* Kd, double* Ka, double* Kb, void* stream);

/* Multi-GPU building blocks (one process per GPU).  The raw accumulators are 64-bit
 * fixed point, so the cross-rank sum is an INTEGER all-reduce (ncclInt64/ncclSum) and the result is
 * bit-identical for any world_size.
 *   cf_accumulate_device : this rank's partition -> acc (device, int64[cf_acc_len(h,nk)]), zeroed first
 *   (caller all-reduces the first cf_acc_reduce_len(h,nk) words of acc)
 *   cf_finalize_device   : acc -> J, K (device, col-major nbf x nbf)
 * Accumulator layout: [J | K_0..K_{nk-1} | J low limb | tail]; the tail (8 words) holds the fixed-point scales of the
 * build that filled THIS accumulator (every rank derives bit-identical scales from the same densities, so the tail is
 * not part of the all-reduce).  Scales therefore travel with their accumulator: accumulate(A), accumulate(B),
 * finalize(A) is well defined.  The densities' Cartesian work space is per handle: calls on one handle must be issued
 * on ONE stream (or be ordered by the caller); a handle is not re-entrant, like the reference's class.
 * The integer sums are taken modulo 2^64: partial sums (per rank, per limb) may wrap around, the final value fits. */
size_t cf_acc_reduce_len(const cf_handle* h, int nk);
size_t cf_acc_len(const cf_handle* h, int nk);
int cf_accumulate_device(cf_handle* h, int nbf,
                         const double* Dd, const double* Da, const double* Db, double exx,
                         int64_t* acc, void* stream);
int cf_finalize_device(cf_handle* h, int nbf, const int64_t* acc, double exx,
                       int has_d, int has_a, int has_b,
                       double* J, double* Kd, double* Ka, double* Kb, void* stream);

/* Multi-density build; replaces Int4C2E::ContractInts(std::vector<EigenMatrix>&, ...) /
 * GhfMultiple (Int4C2E.cpp:685-745): G_k = J[2 D_k] - exx * K[D_k].  HOST pointers,
 * Ds/Gs are nmat consecutive nbf x nbf col-major matrices. */
int cf_build_g_multi(cf_handle* h, int nbf, int nmat, const double* Ds, double exx, double* Gs);

/* Nuclear-gradient contraction; replaces Int4C2E::ContractGrads(D1, D2, output) (Int4C2E.cpp:747-763)
 * on top of getRepulsion1 (:312-408):  grad[3*atom + xyz] = sum_ij D1_ij G^(atom,xyz)[D2]_ij,  G^(A,x) = d/dA_x of
 * (J[2 D2] - exx K[D2]) at fixed D2, the derivative acting on the four centres of every ERI (libint2's 12 buffers,
 * :377-389).  The reference's 3*natom intermediate matrices are not formed.  HOST pointers; D1, D2 symmetric nbf x nbf;
 * needs cf_basis.shell2atom at cf_create.  Quartets are screened with the handle's Schwarz threshold like the reference's
 * list.  A handle with world_size > 1 returns ITS PARTITION's share: sum the vectors over the ranks. */
int cf_contract_grads(cf_handle* h, int nbf, const double* D1, const double* D2, double exx, int natom, double* grad);

/* One-electron integrals: overlap, kinetic energy and nuclear attraction of `natom` point charges Z at
 * xyz (bohr, [3*natom]); replaces Int2C1E::CalculateIntegrals(0, ...) for Overlap / Kinetic / Nuclear
 * (src/Integral/Int2C1E.cpp:18-67, :313-333; multipoles and ECP terms are outside this engine).  nbf x nbf col-major,
 * exactly symmetric.  Z / xyz are HOST pointers in both forms; S, T, V are host pointers (cf_one_electron) or device
 * pointers on the handle's device (cf_one_electron_device, enqueued on `stream`, not synchronised). */
int cf_one_electron(cf_handle* h, int natom, const double* Z, const double* xyz, double* S, double* T, double* V);
int cf_one_electron_device(cf_handle* h, int natom, const double* Z, const double* xyz, double* S, double* T, double* V, void* stream);

/* Matrix form of the gradient contraction; replaces Int4C2E::ContractGrads(D, output) (Int4C2E.cpp:766-790, consumer
 * Restricted/Hess.cpp:72): G[(3*atom + xyz) * nbf*nbf ...] = d/dR_(atom,xyz) ( J[2 D] - exx K[D] ) at fixed D, 3*natom
 * symmetric nbf x nbf col-major matrices, HOST pointers.  (The reference keeps the result in its GradCache; the C++
 * adaptor does the same.)  Single-device handles; with world_size > 1 the partition's share (sum over ranks). */
int cf_contract_grads_matrices(cf_handle* h, int nbf, const double* D, double exx, int natom, double* G);

/* Nuclear-Hessian contraction (tail); replaces Int4C2E::ContractHesss(D1, D2, output)
 * (Int4C2E.cpp:792-811, which uses D = D2 only, :793) on top of getRepulsion2 (:410-492): hess[(3*atomY + y) * 3*natom +
 * 3*atomX + x] = sum over unique ERIs of d^2 (ab|cd) / dR_X,x dR_Y,y * [ 2 D_ab D_cd - exx/2 (D_ac D_bd + D_ad D_bc) ] at
 * fixed D, the derivatives acting on all ordered pairs of the four centres (libint2's 78 buffers, :468-477, and the
 * raw + raw^T - diag assembly of :487-491).  3*natom x 3*natom col-major, exactly symmetric, HOST pointers; consumer
 * Restricted/Hess.cpp:67.  Needs cf_basis.shell2atom.  With world_size > 1 the partition's share (sum over ranks);
 * multi-device handles return the total. */
int cf_contract_hess(cf_handle* h, int nbf, const double* D, double exx, int natom, double* hess);

/* After the caller has synchronised the stream of a *_device call: refresh ms_device_last / ms_eri_last and run the
 * fixed-point range check of that build (the *_device calls never synchronise, so CF_ERR_RANGE for a non-finite or
 * astronomically large density is reported here; cf_build_jk reports it itself). */
int cf_sync_stats(cf_handle* h);
/* Developer/benchmark aid: run every (bra class, ket class) kernel alone and time it with CUDA events.
 * DEVICE density pointers. rows: 8 doubles each = bra class, ket class, quartets, ms, F_alg at nominal contraction depth,
 * threads per quartet, primitive quartets executed by the launch, model flops of the executed work. */
int cf_profile_tasks(cf_handle* h, int nbf, const double* Dd, const double* Da, const double* Db, double exx,
                     double* rows, int max_rows, int* nrows);

/* Library/device facts for logs and benchmarks. */
int cf_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor);
/* Measured FP64 FMA peak of the device (a register-resident DFMA loop, CUDA events), TFLOP/s. */
int cf_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CHINIUM_FOCK_H */
