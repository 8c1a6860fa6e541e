// bf16_cross: can the two small cross terms of the 3xTF32 split run as ONE bf16 MMA?
//
//   x*y ~= hi(x)*hi(y) [tf32, K=8]  +  [bf16(hi x) | bf16(lo x)] . [bf16(lo y) ; bf16(hi y)]  [kind::f16, K=16]
//
// Checks (1) the TMEM layout of a 16-bit A operand (two K elements per 32-bit column, low half first), (2) the
// accuracy of the mixed product against fp64, (3) the sustained cost of alternating tf32 N=64 / bf16 N=64 MMAs
// (today's rank-64 engine issues tf32 N=128 + tf32 N=64 per K step: 64 + 48 clk).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/bf16_cross probes/bf16_cross.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {          // a -> low half, b -> high half, RN
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
#define TMEM_ST8(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory")
#define TMEM_LD16(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr))

constexpr int N = 64, KSTEPS = 4;        // one 32-element K unit

// A: [128][32] fp32 row-major; B: [64][32] fp32 (row n, col k); out: [128][64] = A B^T via the mixed scheme
__global__ void __launch_bounds__(128) check_kernel(const float* A, const float* B, float* out, int mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* bhi = reinterpret_cast<float*>(smem);                 // [64 rows][128 B] tf32 hi, SW128
    uint32_t* bx = reinterpret_cast<uint32_t*>(smem + 8192);     // [64 rows][128 B] bf16 [lo8 | hi8] per K step, SW128
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 64 * 32; e += 128) {
        const int n = e >> 5, k = e & 31;
        const float v = B[n * 32 + k];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        const int chunk = k >> 2, within = k & 3;
        bhi[n * 32 + ((chunk ^ (n & 7)) << 2) + within] = h;
    }
    for (int e = tid; e < 64 * 32; e += 128) {                   // word w of row n: K step s = w / 8, j = w % 8
        const int n = e >> 5, w = e & 31, s = w >> 3, j = w & 7;
        const int k0 = 8 * s + 2 * (j & 3);
        float v0 = B[n * 32 + k0], v1 = B[n * 32 + k0 + 1];
        float h0 = __uint_as_float(__float_as_uint(v0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(v1) & 0xFFFFE000u);
        const uint32_t word = j < 4 ? pack_bf16(v0 - h0, v1 - h1) : pack_bf16(h0, h1);     // [lo | hi]
        const int chunk = w >> 2, within = w & 3;
        bx[n * 32 + ((chunk ^ (n & 7)) << 2) + within] = word;
    }
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
    // A operand: lane = row.  columns [256, 288): tf32 hi;  [288, 320): bf16 [hi8 | lo8] per K step
    {
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        const float* a = A + tid * 32;
        for (int s = 0; s < KSTEPS; ++s) {
            uint32_t hi[8], xw[8];
            float h[8], l[8];
            for (int j = 0; j < 8; ++j) {
                const float v = a[8 * s + j];
                h[j] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                l[j] = v - h[j];
                hi[j] = __float_as_uint(h[j]);
            }
            for (int j = 0; j < 4; ++j) { xw[j] = pack_bf16(h[2 * j], h[2 * j + 1]); xw[4 + j] = pack_bf16(l[2 * j], l[2 * j + 1]); }
            TMEM_ST8(lane_addr + 256 + 8 * s, hi);
            TMEM_ST8(lane_addr + 288 + 8 * s, xw);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint64_t dh = desc_sw128(smem_u32(bhi)), dx = desc_sw128(smem_u32(bx));
        for (int s = 0; s < KSTEPS; ++s) {
            mma_ts_tf32(tmem + 0, tmem + 256 + 8 * s, dh + 2 * s, idesc_tf32(128, N), s > 0);           // hi*hi -> cols [0, 64)
            if (mode == 1) mma_ts_f16(tmem + 64, tmem + 288 + 8 * s, dx + 2 * s, idesc_bf16(128, N), s > 0);   // cross -> [64, 128)
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0; long long spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            if (++spins > (1LL << 24)) break;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t r0[16], r1[16];
            TMEM_LD16(lane_addr + c0, r0);
            TMEM_LD16(lane_addr + 64 + c0, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 16; ++c) out[tid * N + c0 + c] = __uint_as_float(r0[c]) + (mode == 1 ? __uint_as_float(r1[c]) : 0.f);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// sustained cost: MODE 0 = today's pair (tf32 N=128 + tf32 N=64), 1 = tf32 N=64 + bf16 N=64, 2 = bf16 N=64 only, 3 = tf32 N=64 only
template <int MODE>
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
        uint64_t bd[4]; uint32_t at[4];
        for (int k = 0; k < 4; ++k) { bd[k] = desc_sw128(smem_u32(smem + 16384) + k * 32); at[k] = tmem + 448 + k * 8; }
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) { mma_ts_tf32(tmem, at[i & 3], bd[i & 3], idesc_tf32(128, 128), 1); mma_ts_tf32(tmem + 64, at[i & 3], bd[i & 3], idesc_tf32(128, 64), 1); }
                if (MODE == 1) { mma_ts_tf32(tmem, at[i & 3], bd[i & 3], idesc_tf32(128, 64), 1); mma_ts_f16(tmem + 64, at[i & 3], bd[i & 3], idesc_bf16(128, 64), 1); }
                if (MODE == 2) { mma_ts_f16(tmem, at[i & 3], bd[i & 3], idesc_bf16(128, 64), 1); mma_ts_f16(tmem + 64, at[i & 3], bd[i & 3], idesc_bf16(128, 64), 1); }
                if (MODE == 3) { mma_ts_tf32(tmem, at[i & 3], bd[i & 3], idesc_tf32(128, 64), 1); mma_ts_tf32(tmem + 64, at[i & 3], bd[i & 3], idesc_tf32(128, 64), 1); }
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
        out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int MODE>
void rate(long long* d, const char* what) {
    const int reps = 64;
    CK(cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int w = 0; w < 2; ++w) { rate_kernel<MODE><<<1, 128, 64 * 1024>>>(reps, d); CK(cudaDeviceSynchronize()); }
    long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
    printf("[rate] %-44s %7.1f clk per K step (pair of MMAs)\n", what, (double)c / (reps * 16.0));
}

int main() {
    const int M = 128, K = 32;
    float *hA = new float[M * K], *hB = new float[N * K], *hO = new float[M * N];
    for (int variant = 0; variant < 2; ++variant) {
        srand(7 + variant);
        for (int i = 0; i < M * K; ++i) hA[i] = variant ? (float)rand() / RAND_MAX - 0.5f : (float)rand() / RAND_MAX;
        for (int i = 0; i < N * K; ++i) hB[i] = variant ? (float)rand() / RAND_MAX - 0.5f : (float)rand() / RAND_MAX;
        float *dA, *dB, *dO;
        CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dB, N * K * 4)); CK(cudaMalloc(&dO, M * N * 4));
        CK(cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
        for (int mode = 0; mode < 2; ++mode) {
            check_kernel<<<1, 128, 32 * 1024>>>(dA, dB, dO, mode);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(hO, dO, M * N * 4, cudaMemcpyDeviceToHost));
            double num = 0, den = 0, worst = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    double t = 0;
                    for (int k = 0; k < K; ++k) t += (double)hA[m * K + k] * (double)hB[n * K + k];
                    const double e = hO[m * N + n] - t;
                    num += e * e; den += t * t;
                    if (fabs(e) > worst) worst = fabs(e);
                }
            printf("[check] %s data, %s: rel Frobenius error %.3e, max abs %.3e\n", variant ? "zero-mean" : "uniform",
                   mode ? "tf32 hi*hi + bf16 cross (K=16)" : "tf32 hi*hi only", sqrt(num / den), worst);
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dO);
    }
    long long* d; CK(cudaMalloc(&d, 8));
    rate<0>(d, "today: tf32 N=128 + tf32 N=64");
    rate<1>(d, "new:   tf32 N=64 + bf16 N=64 (K=16)");
    rate<2>(d, "bf16 N=64 + bf16 N=64");
    rate<3>(d, "tf32 N=64 + tf32 N=64");
    return 0;
}
