// cf_common.cuh -- device-side data layout shared by all kernels of the J/K engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define CF_LMAX_DEV 3            // highest shell l the instantiated kernels cover (f); tables go to 9 roots
#define CF_NCLS ((CF_LMAX_DEV + 1) * (CF_LMAX_DEV + 2) / 2)   // pair classes (la>=lb)

__host__ __device__ constexpr int cf_ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int cf_pair_class(int la, int lb) { return la * (la + 1) / 2 + lb; }  // la>=lb

// Shell pairs of one class (la,lb), Schwarz-sorted (descending).  SoA, device pointers.
struct PairClassDev {
    int npair;
    const int* sa;          // [npair] shell with the larger l (ties: larger index)
    const int* sb;
    const int* cao_a;       // [npair] first Cartesian AO of shell a / b
    const int* cao_b;
    const int* prim_off;    // [npair] first primitive pair
    const int* nprim;       // [npair] surviving primitive pairs
    const double* A;        // [npair*3] centre of a
    const double* AB;       // [npair*3] A - B
    const double* Q;        // [npair] Schwarz bound max_ij sqrt((ij|ij)) over Cartesian functions
    // primitive pairs (shared pool of the class)
    const double* p;        // exponent sum
    const double* P;        // [*3] product centre
    const double* c;        // ca*cb*exp(-ab/p |AB|^2) * sqrt(2) pi^(5/4) / p
};

struct RysTablesDev {
    const double* table;
    const double* asym;
};

// One launch = one (bra class, ket class) rectangle/triangle of the pair x pair grid.
struct QuartetTask {
    PairClassDev bra, ket;
    const long long* qoff;  // [bra.npair+1] prefix sum of ket counts per bra pair (kets form a prefix of the sorted list)
    long long nquartet;     // qoff[bra.npair]
    long long q_begin, q_end;   // this rank's slice (chunk-interleaved, see engine.cu)
    int chunk;              // quartets per work chunk
    int rank, world;
    int same_class;         // bra class == ket class (triangular, diagonal weight 1/2)
    int ncart;              // leading dimension of the Cartesian matrices
    int nk;                 // number of exchange densities (0..3)
    const double* Dtot;     // [ncart*ncart] Cartesian total density (2Dd+Da+Db), symmetric
    const double* Dk[3];    // Cartesian exchange densities
    long long* accJ;        // [ncart*ncart] fixed-point raw J
    long long* accK[3];
    double scaleJ, scaleK;  // powers of two
    double* store;          // STORE mode: Cartesian blocks, NOUT doubles per quartet in flat order
    int diag;               // Schwarz mode: quartet q is (pair q | pair q)
    RysTablesDev rys;
    double prim_cut;        // skip primitive quartets with |c_ab c_cd| below this
};

