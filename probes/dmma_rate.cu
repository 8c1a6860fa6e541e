// dmma_rate: sustained fp64 throughput of one SM — DMMA.8x8x4 (mma.sync m8n8k4 f64) against plain DFMA, as a function
// of the number of resident warps.  Decides whether a tensor-core path can lift the fp64 MTTKRP/TTM (SIMT kernel:
// 29 % of the DFMA peak, bench.py `fp64`).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/dmma_rate probes/dmma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE>   // 0: DMMA, 8 independent accumulator tiles per warp; 1: DFMA, 16 independent chains per thread
__global__ void rate_kernel(int iters, double* out, long long* clk) {
    double acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0000001, b = 0.9999999;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int t = 0; t < 8; ++t)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[2 * t]), "+d"(acc[2 * t + 1]) : "d"(a), "d"(b));
        } else {
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t] = fma(acc[t], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
    double* d; long long* c;
    CK(cudaMalloc(&d, 8 * 1024 * 256)); CK(cudaMalloc(&c, 8));
    const int iters = 4096;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        for (int mode = 0; mode < 2; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) rate_kernel<0><<<1, warps * 32>>>(iters, d, c); else rate_kernel<1><<<1, warps * 32>>>(iters, d, c);
                CK(cudaDeviceSynchronize());
            }
            long long clk; CK(cudaMemcpy(&clk, c, 8, cudaMemcpyDeviceToHost));
            const double fma = mode == 0 ? (double)iters * 8 * 256 * warps : (double)iters * 16 * 32 * warps;
            printf("[fp64] %-5s %2d warps on one SM: %7.1f FMA/clk/SM\n", mode == 0 ? "DMMA" : "DFMA", warps, fma / clk);
        }
    }
    return 0;
}
