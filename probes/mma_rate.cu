// mma_rate: sustained cost of tcgen05.mma kind::tf32 (M=128, K=8) as a function of N, A-operand source
// (TMEM vs shared memory) and the number of independent accumulators, with the issue loop fully unrolled
// and every operand precomputed, so that the tensor pipe — not the issuing thread — is what is measured.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/mma_rate probes/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
template <int N, int CHAINS, int TS>
__global__ void __launch_bounds__(128) rate_kernel(int reps, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        constexpr uint32_t idesc = idesc_tf32(128, N);
        uint64_t ad[4], bd[4]; uint32_t at[4], dt[CHAINS];
        for (int k = 0; k < 4; ++k) { ad[k] = desc_sw128(smem_u32(smem) + k * 32); bd[k] = desc_sw128(smem_u32(smem + 16384) + k * 32); at[k] = tmem + 448 + k * 8; }
        for (int c = 0; c < CHAINS; ++c) dt[c] = tmem + c * N;
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (TS) mma_ts(dt[i % CHAINS], at[i & 3], bd[i & 3], idesc);
                else mma_ss(dt[i % CHAINS], ad[i & 3], bd[i & 3], idesc);
            }
            if ((r & 7) == 7) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t done = 0; long long spins = 0;
                while (!done) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
                    if (++spins > (1LL << 24)) { out[0] = -1; break; }
                }
                phase ^= 1;
            }
        }
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N, int CHAINS, int TS>
void run(long long* d) {
    const int reps = 64;   // 64 x 32 = 2048 MMAs, a commit+wait every 256
    CK(cudaFuncSetAttribute(rate_kernel<N, CHAINS, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int w = 0; w < 2; ++w) { rate_kernel<N, CHAINS, TS><<<1, 128, 64 * 1024>>>(reps, d); CK(cudaDeviceSynchronize()); }
    long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    printf("[mma] %s tf32 M=128 N=%3d K=8 chains=%d: %7.1f clk/MMA  (ideal %d)\n", TS ? "A=TMEM" : "A=smem", N, CHAINS, (double)c / (reps * 32.0), N / 2);
}
int main() {
    long long* d; CK(cudaMalloc(&d, 8));
    run<32, 1, 1>(d); run<32, 2, 1>(d); run<32, 4, 1>(d);
    run<64, 1, 1>(d); run<64, 2, 1>(d); run<64, 4, 1>(d);
    run<96, 1, 1>(d); run<96, 2, 1>(d);
    run<128, 1, 1>(d); run<128, 2, 1>(d);
    run<256, 1, 1>(d);
    run<32, 1, 0>(d); run<32, 4, 0>(d); run<64, 1, 0>(d); run<64, 4, 0>(d); run<128, 1, 0>(d); run<128, 2, 0>(d); run<256, 1, 0>(d);
    return 0;
}
