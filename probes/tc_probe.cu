// tc_probe: hardware facts the tcgen05 MTTKRP/TTM kernels depend on, measured on the B200.
//   1. SS-mode kind::tf32 MMA with K-major SWIZZLE_128B operands written by plain st.shared
//      (validates the smem/instruction descriptors used in mttkrp_tc.cu)
//   2. TS-mode MMA with the A operand written to TMEM by tcgen05.st.32x32b (lane = row, column = k)
//   3. operand handling (truncation vs rounding of fp32 -> tf32) and accumulator rounding (RZ vs RN)
//   4. read-only streaming bandwidth for the access patterns MTTKRP needs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/tc_probe probes/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                           // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // SBO: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

#define TMEM_LD32(taddr, r) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]), \
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) \
        : "r"(taddr))
#define TMEM_ST32(taddr, r) \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " \
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
        :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]), \
           "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]) : "memory")

// One CTA, 128 threads. A: [128][K] row-major, B: [32][K] row-major (i.e. B^T, K-major), D: [128][32].
// K multiple of 32, K <= 192.  mode 0: A from smem (SS); mode 1: A from TMEM (TS).
constexpr int PN = 32;
__global__ void __launch_bounds__(128) probe_mma_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ D, int K, int mode, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nslab = K / 32;
    unsigned char* smem_al = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1 KB alignment
    unsigned char* sA = smem_al;                            // nslab * 16 KB
    unsigned char* sB = smem_al + (size_t)nslab * 16384;       // nslab * 4 KB (32 rows * 128 B)

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // canonical K-major SWIZZLE_128B: row r of a slab at r*128 B, 16-byte chunk c stored at c ^ (r & 7)
    for (int e = tid; e < 128 * K; e += 128) {
        int r = e / K, k = e % K;
        int slab = k / 32, kk = k % 32;
        uint32_t off = slab * 16384 + r * 128 + (((kk >> 2) ^ (r & 7)) << 4) + (kk & 3) * 4;
        *reinterpret_cast<float*>(sA + off) = A[e];
    }
    for (int e = tid; e < PN * K; e += 128) {
        int r = e / K, k = e % K;
        int slab = k / 32, kk = k % 32;
        uint32_t off = slab * 4096 + r * 128 + (((kk >> 2) ^ (r & 7)) << 4) + (kk & 3) * 4;
        *reinterpret_cast<float*>(sB + off) = B[e];
    }
    const uint32_t a_col0 = 64;   // TMEM columns [64, 64+K) hold A in TS mode
    if (mode == 1) {
        for (int slab = 0; slab < nslab; ++slab) {
            uint32_t r[32];
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(A[(size_t)tid * K + slab * 32 + i]);
            uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + a_col0 + slab * 32;
            TMEM_ST32(taddr, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, PN);
        for (int ks = 0; ks < K / 8; ++ks) {
            int slab = ks / 4, sub = ks % 4;
            uint64_t bdesc = make_desc_kmajor_sw128(smem_u32(sB + slab * 4096) + sub * 32);
            if (mode == 0) {
                uint64_t adesc = make_desc_kmajor_sw128(smem_u32(sA + slab * 16384) + sub * 32);
                mma_ss(tmem_base, adesc, bdesc, idesc, ks > 0);
            } else {
                mma_ts(tmem_base, tmem_base + a_col0 + ks * 8, bdesc, idesc, ks > 0);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    // bounded wait on phase 0
    uint32_t done = 0;
    for (int it = 0; it < 20000000 && !done; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    }
    if (!done) { if (tid == 0) *status = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        uint32_t r[32];
        uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        TMEM_LD32(taddr, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) D[(size_t)tid * PN + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512));
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rna_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static int run_mma(const std::vector<float>& A, const std::vector<float>& B, int K, int mode, std::vector<float>& D) {
    float *dA, *dB, *dD; int* dS;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * PN * 4)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * PN * 4)); CK(cudaMemset(dS, 0, 4));
    size_t smem = (size_t)(K / 32) * (16384 + 4096) + 1024;
    CK(cudaFuncSetAttribute(probe_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_mma_kernel<<<1, 128, smem>>>(dA, dB, dD, K, mode, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
    int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    D.resize(128 * PN);
    CK(cudaMemcpy(D.data(), dD, 128 * PN * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
    return st;
}

static void probe_mma() {
    for (int mode = 0; mode < 2; ++mode) {
        const int K = 96;
        std::vector<float> A(128 * K), B(PN * K), D;
        srand(1);
        for (auto& v : A) v = (float)(rand() % 2001 - 1000) / 1000.0f;
        for (auto& v : B) v = (float)(rand() % 2001 - 1000) / 1000.0f;
        int st = run_mma(A, B, K, mode, D);
        double err_t = 0, err_r = 0, err_f = 0, nrm = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < PN; ++n) {
            double st_ = 0, sr = 0, sf = 0;
            for (int k = 0; k < K; ++k) {
                st_ += (double)trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
                sr += (double)rna_tf32(A[m * K + k]) * rna_tf32(B[n * K + k]);
                sf += (double)A[m * K + k] * B[n * K + k];
            }
            double d = D[m * PN + n];
            err_t += (d - st_) * (d - st_); err_r += (d - sr) * (d - sr); err_f += (d - sf) * (d - sf); nrm += sf * sf;
        }
        printf("[mma %s] timeout=%d  rel err vs trunc-tf32 inputs %.3e | vs rna-tf32 inputs %.3e | vs fp32 inputs %.3e\n",
               mode == 0 ? "SS" : "TS", st, sqrt(err_t / nrm), sqrt(err_r / nrm), sqrt(err_f / nrm));
        printf("    D[0][0..3] = %g %g %g %g ; D[127][28..31] = %g %g %g %g\n", D[0], D[1], D[2], D[3],
               D[127 * PN + 28], D[127 * PN + 29], D[127 * PN + 30], D[127 * PN + 31]);
    }
    // accumulator rounding: one product of 1.0 followed by 23 separate MMAs each adding 0.75 ulp
    for (int mode = 0; mode < 2; ++mode) {
        const int K = 192;
        std::vector<float> A(128 * K, 0.f), B(PN * K, 0.f), D;
        for (int m = 0; m < 128; ++m) for (int s = 0; s < K / 8; ++s) A[m * K + 8 * s] = s == 0 ? 1.0f : ldexpf(1.5f, -24);
        for (int n = 0; n < PN; ++n) for (int s = 0; s < K / 8; ++s) B[n * K + 8 * s] = 1.0f;
        run_mma(A, B, K, mode, D);
        float rn = 1.0f + 23 * ldexpf(1.0f, -23);
        printf("[acc rounding %s] D = 1 + %.3f ulp  (RN expects +23 ulp = %.9g, RZ expects +0; exact sum = +17.25 ulp)\n",
               mode == 0 ? "SS" : "TS", (D[5 * PN + 7] - 1.0f) / ldexpf(1.0f, -23), rn);
        // all 8 k of ONE mma tiny: intra-instruction summation
        std::fill(A.begin(), A.end(), 0.f); std::fill(B.begin(), B.end(), 0.f);
        for (int m = 0; m < 128; ++m) { A[m * K + 0] = 1.0f; for (int k = 8; k < 16; ++k) A[m * K + k] = ldexpf(1.0f, -26); }
        for (int n = 0; n < PN; ++n) { B[n * K + 0] = 1.0f; for (int k = 8; k < 16; ++k) B[n * K + k] = 1.0f; }
        run_mma(A, B, K, mode, D);
        printf("[intra-mma sum %s] 1 + 8 x 2^-26 (=0.25 ulp total... exact 1+2^-23): D = 1 + %.3f ulp\n", mode == 0 ? "SS" : "TS",
               (D[5 * PN + 7] - 1.0f) / ldexpf(1.0f, -23));
        // operand conversion: 1 + 2^-11 + 2^-12 : trunc -> 1, RN -> 1 + 2^-10
        std::fill(A.begin(), A.end(), 0.f); std::fill(B.begin(), B.end(), 0.f);
        for (int m = 0; m < 128; ++m) A[m * K] = 1.0f + ldexpf(1.0f, -11) + ldexpf(1.0f, -12);
        for (int n = 0; n < PN; ++n) B[n * K] = 1.0f;
        run_mma(A, B, K, mode, D);
        printf("[operand cvt %s] A = 1+2^-11+2^-12: D = 1 + %.4f * 2^-10  (0 => truncation, 1 => round-to-nearest)\n",
               mode == 0 ? "SS" : "TS", (D[5 * PN + 7] - 1.0f) / ldexpf(1.0f, -10));
    }
}

// ---------------------------------------------------------------------------------------------
// streaming-read bandwidth probes
__global__ void __launch_bounds__(256) read_linear_kernel(const float4* __restrict__ x, size_t n4, float* sink) {
    float acc = 0.f;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(x + i));
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *sink = acc;
}

// tile pattern of mode-0 MTTKRP: CTA (jt, split) walks k over its range, reading 128 rows x W bytes per step,
// rows `row_stride` floats apart.
__global__ void __launch_bounds__(256) read_tiles_kernel(const float* __restrict__ x, size_t row_stride, int rows_per_tile,
                                                         int wfloats, size_t k_per_cta, int jtiles, float* sink) {
    const int jt = blockIdx.x % jtiles;
    const size_t split = blockIdx.x / jtiles;
    const size_t k0 = split * k_per_cta;
    float acc = 0.f;
    const int vec_per_row = wfloats / 4;
    const int vecs = rows_per_tile * vec_per_row;
    for (size_t k = k0; k < k0 + k_per_cta; k += wfloats) {
        for (int e = threadIdx.x; e < vecs; e += 256) {
            int r = e / vec_per_row, c = e % vec_per_row;
            const float4* p = reinterpret_cast<const float4*>(x + ((size_t)jt * rows_per_tile + r) * row_stride + k) + c;
            float4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

static void probe_bandwidth() {
    const size_t N = (size_t)1 << 30;  // 4 GiB of floats
    float* x; float* sink;
    CK(cudaMalloc(&x, N * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(x, 0, N * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time_it = [&](auto launch, const char* name) {
        for (int i = 0; i < 2; ++i) launch();
        CK(cudaEventRecord(e0));
        const int reps = 5;
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("[bw] %-58s %8.1f GB/s\n", name, (double)N * 4 * reps / (ms * 1e-3) / 1e9);
    };
    for (int mult : {4, 8, 16, 32}) {
        char name[128]; snprintf(name, sizeof name, "linear float4, grid = 148 x %d, 256 thr", mult);
        time_it([&] { read_linear_kernel<<<148 * mult, 256>>>(reinterpret_cast<const float4*>(x), N / 4, sink); }, name);
    }
    // mode-0 pattern: X[1024 rows][1M floats]; 8 j-tiles of 128 rows; splits chosen so that grid ~ 148*4
    for (int w : {32, 64, 128, 256}) {
        for (int splits : {74, 148, 296}) {
            size_t kper = ((size_t)1 << 20) / splits / w * w;
            char name[128]; snprintf(name, sizeof name, "tiles 128 rows x %4d B, row stride 4 MiB, %3d splits x 8 jt", w * 4, splits);
            time_it([&] { read_tiles_kernel<<<8 * splits, 256>>>(x, (size_t)1 << 20, 128, w, kper, 8, sink); }, name);
        }
    }
    // middle-mode pattern: per a, X[a][1024 rows][1024 floats] (row stride 4 KiB)
    for (int w : {32, 128}) {
        char name[128]; snprintf(name, sizeof name, "tiles 128 rows x %4d B, row stride 4 KiB (middle mode), 592 CTAs", w * 4);
        // emulate: treat the tensor as 8192 j-tiles... each CTA walks k over one a-slab range
        time_it([&] { read_tiles_kernel<<<8 * 74, 256>>>(x, (size_t)1 << 10, 128, w, 1024, 8, sink); }, name);
    }
    cudaFree(x); cudaFree(sink);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device: %s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    if (argc < 2 || strcmp(argv[1], "bw") != 0) probe_mma();
    if (argc < 2 || strcmp(argv[1], "mma") != 0) probe_bandwidth();
    return 0;
}
