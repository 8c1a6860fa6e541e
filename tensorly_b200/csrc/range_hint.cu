// Range hint for the fp16-split tensor-core engine (tc_stream.cu, HF variant).
//
// The tf32 engine needs nothing but the data: tf32 has fp32's exponent.  Its fp16 sibling — half the MMA
// instructions and operand traffic per tensor byte, which is what the rank-64 configs (C4 / C5) are bound by under
// the 1 kW power cap — computes on x * 2^k with k chosen so that max |x| lands in [2^14, 2^15): it needs max |x|.
// A stateless MTTKRP call cannot afford a pass over the tensor to find it, so the caller that owns the tensor (the
// ALS drivers: the tensor is constant over the whole decomposition) computes it ONCE with tlb200_tensor_absmax and
// registers it with tlb200_hint_tensor_absmax; MTTKRP / TTM calls on that base pointer then take the fp16 engine.
// The value stays on the device (no host sync); no hint, no fp16.
//
// Also here: the per-column fp16 split of the small operand (one CTA per column finds the column's own power-of-two
// scale, so factor columns of very different magnitude all keep 22 significant bits).
#include "hf_split.cuh"

#include <mutex>

namespace tlb200 {
namespace {

constexpr int kAbsmaxBlocks = kNumSMs * 8;

__global__ void __launch_bounds__(256)
absmax_f4_kernel(const float4* __restrict__ x, int64_t n4, const float* __restrict__ tail, int ntail, unsigned* __restrict__ out_bits) {
    unsigned m = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        const float4 v = __ldg(x + i);
        m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) m = max(m, __float_as_uint(fabsf(tail[threadIdx.x])));
    // |x| as an unsigned integer orders like |x| (NaN sorts above everything: a NaN in the tensor poisons the hint
    // the same way it poisons the result)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ unsigned red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = max(m, red[i]);
        atomicMax(out_bits, m);
    }
}

__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n, unsigned* __restrict__ out_bits) {
    unsigned m = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        m = max(m, __float_as_uint(fabsf(__ldg(x + i))));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ unsigned red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = max(m, red[i]);
        atomicMax(out_bits, m);
    }
}

struct Hint { const void* x; const float* absmax; };
constexpr int kMaxHints = 32;
Hint g_hints[kMaxHints];
int g_nhints = 0;
std::mutex g_hint_mutex;

// one CTA per row i of the matrix (= column of the B operand): scale from the row's own max
__global__ void __launch_bounds__(256)
split_matrix_f16_kernel(const float* __restrict__ m, int64_t I, int64_t J, int64_t mrs, int64_t mcs, int64_t Kpad,
                        __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ col_inv) {
    const int i = blockIdx.x;
    unsigned mx = 0;
    if (i < I)
        for (int64_t j = threadIdx.x; j < J; j += 256) mx = max(mx, __float_as_uint(fabsf(m[i * mrs + j * mcs])));
    float inv;
    const float sc = hf_block_scale(mx, &inv);
    if (threadIdx.x == 0) col_inv[i] = inv;
    for (int64_t j = threadIdx.x; j < Kpad; j += 256) {
        const float v = (i < I && j < J) ? m[i * mrs + j * mcs] * sc : 0.f;
        __half h, l;
        hf_split1(v, h, l);
        hi[(int64_t)i * Kpad + j] = h;
        lo[(int64_t)i * Kpad + j] = l;
    }
}

}  // namespace

const float* tc_range_hint(const void* x) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("TLB200_DISABLE_HF"); off = (e && atoi(e) != 0) ? 1 : 0; }
    if (off) return nullptr;
    std::lock_guard<std::mutex> lk(g_hint_mutex);
    for (int i = 0; i < g_nhints; ++i)
        if (g_hints[i].x == x) return g_hints[i].absmax;
    return nullptr;
}

int launch_split_matrix_f16(const float* m, int64_t I, int64_t J, int64_t mrs, int64_t mcs, int rp, int64_t kpad,
                            __half* hi, __half* lo, float* col_inv, cudaStream_t stream) {
    split_matrix_f16_kernel<<<(unsigned)rp, 256, 0, stream>>>(m, I, J, mrs, mcs, kpad, hi, lo, col_inv);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

}  // namespace tlb200

using namespace tlb200;

extern "C" int tlb200_tensor_absmax(const void* x, int64_t n, int dtype, void* absmax_out, void* stream) {
    if (!x || !absmax_out || n < 0) return TLB200_EINVAL;
    if (dtype != TLB200_F32) return TLB200_EUNSUPPORTED;      // the hint feeds the fp32 tensor-core engine only
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(absmax_out, 0, sizeof(float), s) != cudaSuccess) return TLB200_ECUDA;
    if (n == 0) return TLB200_OK;
    const float* xf = static_cast<const float*>(x);
    if (reinterpret_cast<uintptr_t>(x) % 16 == 0) {
        const int64_t n4 = n / 4;
        absmax_f4_kernel<<<kAbsmaxBlocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x), n4, xf + n4 * 4, (int)(n - n4 * 4),
                                                       static_cast<unsigned*>(absmax_out));
    } else {
        absmax_kernel<<<kAbsmaxBlocks, 256, 0, s>>>(xf, n, static_cast<unsigned*>(absmax_out));
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

extern "C" int tlb200_hint_tensor_absmax(const void* x, const void* absmax_device) {
    if (!x) return TLB200_EINVAL;
    std::lock_guard<std::mutex> lk(g_hint_mutex);
    int at = -1;
    for (int i = 0; i < g_nhints; ++i)
        if (g_hints[i].x == x) { at = i; break; }
    if (!absmax_device) {                       // withdraw
        if (at >= 0) g_hints[at] = g_hints[--g_nhints];
        return TLB200_OK;
    }
    if (at < 0) {
        if (g_nhints == kMaxHints) {            // full: the oldest entry goes
            for (int i = 1; i < kMaxHints; ++i) g_hints[i - 1] = g_hints[i];
            --g_nhints;
        }
        at = g_nhints++;
    }
    g_hints[at].x = x;
    g_hints[at].absmax = static_cast<const float*>(absmax_device);
    return TLB200_OK;
}
