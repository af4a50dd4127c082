// red_micro.cu -- throughput of 64-bit integer RED (atomicAdd without return) to an L2-resident array under the
// access patterns of the J/K digestion.  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_micro red_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
// mode 0: every lane a random element; 1: all lanes of a warp the same random element; 2: lanes contiguous (coalesced)
// 3: random 6-element column segment per lane-group of 6 (like a d-shell row block); 4: smem atomics (64-bit)
__global__ void red(unsigned long long* acc, size_t n, int iters, int mode, unsigned seed) {
    unsigned s = seed + blockIdx.x * 9781u + threadIdx.x * 6271u;
    const int lane = threadIdx.x & 31;
    __shared__ unsigned long long sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        s = s * 1664525u + 1013904223u;
        unsigned r = s >> 4;
        if (mode == 1) r = __shfl_sync(0xffffffffu, r, 0);
        if (mode == 2) r = __shfl_sync(0xffffffffu, r, 0) + lane;
        if (mode == 3) r = __shfl_sync(0xffffffffu, r, (lane / 6) * 6) + lane % 6;
        if (mode == 4) { atomicAdd(&sm[r & 4095], 1ull); continue; }
        atomicAdd(&acc[r % n], 1ull);
    }
    if (mode == 4) { __syncthreads(); if (threadIdx.x == 0) acc[blockIdx.x] = sm[0]; }
}
int main() {
    const size_t n = 684 * 684;   // c18 Cartesian matrix
    unsigned long long* acc; cudaMalloc(&acc, 8 * n * 2); cudaMemset(acc, 0, 16 * n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"random per lane", "same address per warp", "contiguous per warp", "6-element segments", "shared-memory 64-bit atomics"};
    for (int mode = 0; mode < 5; mode++)
        for (int tpb : {128, 512}) {
            const int blocks = 148 * (2048 / tpb) , iters = 2000;
            red<<<blocks, tpb>>>(acc, n, 10, mode, 1);
            cudaEventRecord(e0);
            red<<<blocks, tpb>>>(acc, n, iters, mode, 7);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)blocks * tpb * iters;
            printf("%-30s tpb %3d: %.1f G atomics/s  (%.1f per clk @1.965GHz)\n", names[mode], tpb, ops / (ms * 1e-3) / 1e9, ops / (ms * 1e-3) / 1.965e9);
        }
    return 0;
}
