// tmem_contend: does tcgen05.st traffic (the convert warps) slow tcgen05.mma (and vice versa)?
// Warp 0 lane 0 issues back-to-back TS-mode tf32 MMAs (M=128, N, K=8); warps 4-7 (one per TMEM lane quarter)
// store 32 columns each with tcgen05.st + wait::st in a loop; both sides are timed with clock64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/tmem_contend probes/tmem_contend.cu
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
#define TMEM_ST32(taddr, r) \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " \
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
        :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]), \
           "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]) : "memory")
#define TMEM_LD32(taddr, r) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]), \
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) \
        : "r"(taddr))

// mma_n: number of MMAs the MMA thread issues (0 = none); st_n: number of (4 x st + wait) rounds per store warp (0 = none)
// ld_n: rounds of tcgen05.ld by warps 8-11
template <int N>
__global__ void __launch_bounds__(384) contend_kernel(int mma_n, int st_n, int ld_n, int wait_each, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
    if (tid == 0 && mma_n > 0) {
        constexpr uint32_t idesc = idesc_tf32(128, N);
        uint64_t bd[4]; uint32_t at[4];
        for (int k = 0; k < 4; ++k) { bd[k] = desc_sw128(smem_u32(smem + 16384) + k * 32); at[k] = tmem + 448 + k * 8; }
        long long t0 = clock64();
        for (int r = 0; r < mma_n / 32; ++r) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mma_ts(tmem, at[i & 3], bd[i & 3], idesc);
            if ((r & 7) == 7 || r == mma_n / 32 - 1) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t done = 0; long long spins = 0; const uint32_t phase = (r >> 3) & 1;
                while (!done) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
                    if (++spins > (1LL << 24)) { out[0] = -1; break; }
                }
            }
        }
        out[0] = clock64() - t0;
    }
    if (warp >= 4 && warp < 8 && st_n > 0) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;     // columns 256..383
        uint32_t r[32];
        for (int i = 0; i < 32; ++i) r[i] = tid * 31 + i;
        long long t0 = clock64();
        for (int it = 0; it < st_n; ++it) {
            TMEM_ST32(base, r); TMEM_ST32(base + 32, r); TMEM_ST32(base + 64, r); TMEM_ST32(base + 96, r);
            if (wait_each) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if ((tid & 31) == 0) out[1 + (warp & 3)] = clock64() - t0;
    }
    if (warp >= 8 && warp < 12 && ld_n > 0) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128;     // columns 128..255
        uint32_t r[32]; uint32_t acc = 0;
        long long t0 = clock64();
        for (int it = 0; it < ld_n; ++it) {
            TMEM_LD32(base, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; ++i) acc += r[i];
            TMEM_LD32(base + 32, r); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; ++i) acc += r[i];
        }
        if ((tid & 31) == 0) out[5 + (warp & 3)] = clock64() - t0 + (acc == 12345 ? 1 : 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int N>
void run(long long* d, int mma_n, int st_n, int ld_n, int wait_each) {
    CK(cudaFuncSetAttribute(contend_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    long long h[9];
    for (int w = 0; w < 2; ++w) { CK(cudaMemset(d, 0, 72)); contend_kernel<N><<<1, 384, 64 * 1024>>>(mma_n, st_n, ld_n, wait_each, d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 72, cudaMemcpyDeviceToHost));
    printf("[contend] N=%3d mma=%5d st_rounds=%5d(wait_each=%d) ld_rounds=%5d : ", N, mma_n, st_n, wait_each, ld_n);
    if (mma_n) printf("%6.1f clk/MMA  ", (double)h[0] / mma_n);
    if (st_n) printf("%6.1f clk/st-round(4x32col)  ", (double)h[1] / st_n);
    if (ld_n) printf("%6.1f clk/ld-round(2x32col)", (double)h[5] / ld_n);
    printf("\n");
}
int main() {
    long long* d; CK(cudaMalloc(&d, 72));
    run<64>(d, 2048, 0, 0, 0);
    run<64>(d, 0, 512, 0, 1);
    run<64>(d, 0, 512, 0, 0);
    run<64>(d, 0, 0, 512, 0);
    run<64>(d, 2048, 512, 0, 1);
    run<64>(d, 2048, 512, 0, 0);
    run<64>(d, 2048, 0, 1024, 0);
    run<64>(d, 2048, 512, 1024, 1);
    run<32>(d, 2048, 512, 0, 1);
    run<128>(d, 2048, 512, 0, 1);
    return 0;
}
