// fp16 split helpers shared by the prep kernels of the fp16-split tensor-core engine (see range_hint.cu).
#pragma once
#include "common.cuh"

#include <cuda_fp16.h>
#include <cstdlib>

namespace tlb200 {

// device pointer to max |x| registered for the tensor at `x` (tlb200_hint_tensor_absmax), or null
const float* tc_range_hint(const void* x);

// [rp rows][kpad] fp16 hi / lo tables of m[i * mrs + j * mcs] (zero padded) + the inverse row scales
int launch_split_matrix_f16(const float* m, int64_t I, int64_t J, int64_t mrs, int64_t mcs, int rp, int64_t kpad,
                            __half* hi, __half* lo, float* col_inv, cudaStream_t stream);

// transposed Khatri-Rao table as fp16 hi / lo [pad_cols][rows_padded] + the inverse column scales [pad_cols]
int launch_khatri_rao_t_f16(const float* const* mats, const int64_t* rows, const int64_t* row_stride,
                            const int64_t* col_stride, int nmats, int64_t rank, const float* weights, __half* hi,
                            __half* lo, int64_t rows_padded, int64_t pad_cols, float* col_inv, cudaStream_t stream);

#ifdef __CUDACC__
// power-of-two scale that maps a column whose max |v| has the float bits `mx` into [2^14, 2^15), and its inverse
__device__ __forceinline__ float hf_scale_from_bits(unsigned mx, float* inv) {
    const int e = (int)((mx >> 23) & 0xFFu);
    int se = 268 - e;
    se = se < 1 ? 1 : (se > 253 ? 253 : se);
    *inv = __uint_as_float((unsigned)(254 - se) << 23);
    return __uint_as_float((unsigned)se << 23);
}
// block-wide (256 threads) max of the per-thread |v| bit patterns, then the scale; every thread gets the result
__device__ __forceinline__ float hf_block_scale(unsigned mx, float* inv) {
    __shared__ unsigned hf_red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) hf_red[threadIdx.x >> 5] = mx;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) mx = max(mx, hf_red[i]);
    __syncthreads();
    return hf_scale_from_bits(mx, inv);
}
// hi = RN_fp16(v), lo = RN_fp16((v - hi) * 2^11): v = hi + lo / 2^11 to 22 significant bits
__device__ __forceinline__ void hf_split1(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn((v - __half2float(h)) * 2048.f);
}
#endif

}  // namespace tlb200
