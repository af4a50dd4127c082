// fp64_micro.cu -- DFMA dependent-issue latency and throughput vs (warps per SM, ILP) on the device.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_micro fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, int iters, long long* cycles) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP>
void run(int warps_per_sm, double* out, long long* dcy) {
    const int iters = 1 << 14;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<ILP><<<148, 32 * warps_per_sm>>>(out, iters, dcy);
    cudaEventRecord(e0);
    chain<ILP><<<148, 32 * warps_per_sm>>>(out, iters, dcy);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cy; cudaMemcpy(&cy, dcy, 8, cudaMemcpyDeviceToHost);
    double fl = 2.0 * ILP * iters * 32.0 * warps_per_sm * 148;
    printf("warps/SM %2d ILP %d : %.2f cycles per DFMA-round (%.2f per DFMA), %.2f TFLOP/s\n", warps_per_sm, ILP, (double)cy / iters,
           (double)cy / iters / ILP, fl / (ms * 1e-3) / 1e12);
}
int main() {
    double* out; long long* dcy;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&dcy, 8);
    for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w, out, dcy); run<2>(w, out, dcy); run<4>(w, out, dcy); run<8>(w, out, dcy); }
    return 0;
}
