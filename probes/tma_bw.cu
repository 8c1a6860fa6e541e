// tma_bw: streaming-read bandwidth of TMA tile loads for the MTTKRP mode-0 access pattern
// (128-row tiles, rows 4 MiB apart) as a function of how the tile is cut into boxes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/tma_bw probes/tma_bw.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Each CTA: one producer thread streams its K range through a ring of NS stages; a stage is
// 128 rows x (KO x 128 bytes), fetched as (128/RB) x KO boxes of RB rows x 128 bytes, issued row-group major.
__global__ void __launch_bounds__(32) tma_stream_kernel(const __grid_constant__ CUtensorMap tmap, int RB, int KO, int NS,
                                                        long long k_per_cta, int jtiles, int* status) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[16];
    const int jt = blockIdx.x % jtiles;
    const long long split = blockIdx.x / jtiles;
    const long long k0 = split * k_per_cta;
    const int stage_bytes = 128 * KO * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const long long nst = k_per_cta / (KO * 32);
        auto issue = [&](long long st) {
            const int s = (int)(st % NS);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(stage_bytes) : "memory");
            const long long k = k0 + st * KO * 32;
            for (int rg = 0; rg < 128 / RB; ++rg)
                for (int ko = 0; ko < KO; ++ko) {
                    unsigned char* dst = smem + (size_t)s * stage_bytes + ko * 16384 + rg * RB * 128;
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&full[s])),
                                   "r"((int)(k + ko * 32)), "r"(jt * 128 + rg * RB) : "memory");
                }
        };
        for (long long st = 0; st < nst && st < NS; ++st) issue(st);
        for (long long st = 0; st < nst; ++st) {
            const int s = (int)(st % NS);
            const uint32_t parity = (uint32_t)((st / NS) & 1);
            uint32_t done = 0; long long spins = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&full[s])), "r"(parity) : "memory");
                if (++spins > (1LL << 26)) { *status = 1; return; }
            }
            if (st + NS < nst) issue(st + NS);
        }
    }
}

// Variant C: unswizzled boxes of RB rows x W floats (W*4 bytes contiguous per row).
__global__ void __launch_bounds__(32) tma_streamC_kernel(const __grid_constant__ CUtensorMap tmap, int RB, int W, int NS,
                                                         long long k_per_cta, int jtiles, int* status) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[16];
    const int jt = blockIdx.x % jtiles;
    const long long split = blockIdx.x / jtiles;
    const long long k0 = split * k_per_cta;
    const int stage_bytes = 128 * W * 4;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const long long nst = k_per_cta / W;
        auto issue = [&](long long st) {
            const int s = (int)(st % NS);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(stage_bytes) : "memory");
            for (int rg = 0; rg < 128 / RB; ++rg) {
                unsigned char* dst = smem + (size_t)s * stage_bytes + (size_t)rg * RB * W * 4;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&full[s])),
                               "r"((int)(k0 + st * W)), "r"(jt * 128 + rg * RB) : "memory");
            }
        };
        for (long long st = 0; st < nst && st < NS; ++st) issue(st);
        for (long long st = 0; st < nst; ++st) {
            const int s = (int)(st % NS);
            const uint32_t parity = (uint32_t)((st / NS) & 1);
            uint32_t done = 0; long long spins = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&full[s])), "r"(parity) : "memory");
                if (++spins > (1LL << 26)) { *status = 1; return; }
            }
            if (st + NS < nst) issue(st + NS);
        }
    }
}

// Variant B: one box = RB rows x KO adjacent 128-byte lines (3-D map: 32 floats, K/32 lines, rows), so that the
// KO lines of a row are fetched back to back.  smem layout per box: [row][ko][128 B].
__global__ void __launch_bounds__(32) tma_stream3_kernel(const __grid_constant__ CUtensorMap tmap, int RB, int KO, int NS,
                                                         long long k_per_cta, int jtiles, int* status) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[16];
    const int jt = blockIdx.x % jtiles;
    const long long split = blockIdx.x / jtiles;
    const long long k0 = split * k_per_cta;
    const int stage_bytes = 128 * KO * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const long long nst = k_per_cta / (KO * 32);
        auto issue = [&](long long st) {
            const int s = (int)(st % NS);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(stage_bytes) : "memory");
            const long long kline = (k0 + st * KO * 32) / 32;
            for (int rg = 0; rg < 128 / RB; ++rg) {
                unsigned char* dst = smem + (size_t)s * stage_bytes + (size_t)rg * RB * KO * 128;
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&full[s])),
                               "r"(0), "r"((int)kline), "r"(jt * 128 + rg * RB) : "memory");
            }
        };
        for (long long st = 0; st < nst && st < NS; ++st) issue(st);
        for (long long st = 0; st < nst; ++st) {
            const int s = (int)(st % NS);
            const uint32_t parity = (uint32_t)((st / NS) & 1);
            uint32_t done = 0; long long spins = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(&full[s])), "r"(parity) : "memory");
                if (++spins > (1LL << 26)) { *status = 1; return; }
            }
            if (st + NS < nst) issue(st + NS);
        }
    }
}

int main() {
    const size_t ROWS = 1024, K = (size_t)1 << 20;
    float* x; int* status;
    CK(cudaMalloc(&x, ROWS * K * 4)); CK(cudaMalloc(&status, 4));
    CK(cudaMemset(x, 0, ROWS * K * 4)); CK(cudaMemset(status, 0, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const int jtiles = 8;
    for (int splits : {18, 37}) {
        for (int RB : {128, 32, 8}) {
            CUtensorMap tmap;
            cuuint64_t dims[2] = {K, ROWS}, strides[1] = {K * 4};
            cuuint32_t box[2] = {32, (cuuint32_t)RB}, estr[2] = {1, 1};
            CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            for (int KO : {1, 2, 4}) {
                if (RB == 128 && KO > 1 && false) continue;
                const int stage_bytes = 128 * KO * 128;
                const int NS = (192 * 1024) / stage_bytes > 12 ? 12 : (192 * 1024) / stage_bytes;
                long long kper = (long long)(K / splits) / (KO * 32) * (KO * 32);
                size_t smem = (size_t)NS * stage_bytes + 1024;
                auto launch = [&] { tma_stream_kernel<<<jtiles * splits, 32, smem>>>(tmap, RB, KO, NS, kper, jtiles, status); };
                launch(); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                for (int i = 0; i < 3; ++i) launch();
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
                double bytes = (double)jtiles * splits * kper * 128 * 4 * 3;
                printf("[tma] %3d CTAs, box %3d rows x 128 B, %d k-lines/row-group, %2d stages of %3d KB: %8.1f GB/s%s\n",
                       jtiles * splits, RB, KO, NS, stage_bytes / 1024, bytes / (ms * 1e-3) / 1e9, st ? "  TIMEOUT" : "");
            }
        }
    }
    CK(cudaFuncSetAttribute(tma_stream3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (int splits : {18, 37}) {
        for (int KO : {2, 4, 8}) {
            for (int RB : {128, 64, 32}) {
                if ((size_t)RB * KO * 128 > 65536) continue;
                CUtensorMap tmap;
                cuuint64_t dims[3] = {32, K / 32, ROWS}, strides[2] = {128, K * 4};
                cuuint32_t box[3] = {32, (cuuint32_t)KO, (cuuint32_t)RB}, estr[3] = {1, 1, 1};
                CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode3 failed %d (KO %d RB %d)\n", (int)r, KO, RB); continue; }
                const int stage_bytes = 128 * KO * 128;
                const int NS = (192 * 1024) / stage_bytes > 12 ? 12 : (192 * 1024) / stage_bytes;
                long long kper = (long long)(K / splits) / (KO * 32) * (KO * 32);
                size_t smem = (size_t)NS * stage_bytes + 1024;
                auto launch = [&] { tma_stream3_kernel<<<jtiles * splits, 32, smem>>>(tmap, RB, KO, NS, kper, jtiles, status); };
                launch(); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                for (int i = 0; i < 3; ++i) launch();
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
                double bytes = (double)jtiles * splits * kper * 128 * 4 * 3;
                printf("[tma3] %3d CTAs, box %3d rows x %d adjacent 128B lines, %2d stages of %3d KB: %8.1f GB/s%s\n",
                       jtiles * splits, RB, KO, NS, stage_bytes / 1024, bytes / (ms * 1e-3) / 1e9, st ? "  TIMEOUT" : "");
            }
        }
    }
    // Variant C: no swizzle, box inner = 64 / 128 / 256 floats (256 B .. 1 KB contiguous per row)
    for (int splits : {18, 37}) {
        for (int W : {64, 128, 256}) {
            for (int RB : {128, 32}) {
                if ((size_t)RB * W * 4 > 65536) continue;
                CUtensorMap tmap;
                cuuint64_t dims[2] = {K, ROWS}, strides[1] = {K * 4};
                cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)RB}, estr[2] = {1, 1};
                CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encodeC failed %d (W %d RB %d)\n", (int)r, W, RB); continue; }
                // reuse kernel A with KO = W/32 "lines" but a single box per row group: emulate via RB rows, KO=1 and box bytes
                const int KO = W / 32;
                const int stage_bytes = 128 * KO * 128;
                const int NS = (192 * 1024) / stage_bytes > 12 ? 12 : (192 * 1024) / stage_bytes;
                long long kper = (long long)(K / splits) / W * W;
                size_t smem = (size_t)NS * stage_bytes + 1024;
                auto launch = [&] { tma_streamC_kernel<<<jtiles * splits, 32, smem>>>(tmap, RB, W, NS, kper, jtiles, status); };
                launch(); CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                for (int i = 0; i < 3; ++i) launch();
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                int st; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
                double bytes = (double)jtiles * splits * kper * 128 * 4 * 3;
                printf("[tmaC] %3d CTAs, unswizzled box %3d rows x %4d B, %2d stages of %3d KB: %8.1f GB/s%s\n",
                       jtiles * splits, RB, W * 4, NS, stage_bytes / 1024, bytes / (ms * 1e-3) / 1e9, st ? "  TIMEOUT" : "");
            }
        }
    }
    return 0;
}
