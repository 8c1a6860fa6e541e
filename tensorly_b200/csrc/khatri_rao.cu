// Khatri-Rao product, bit-exact against the reference's left fold of broadcast
// multiplies (tensorly/tenalg/core_tenalg/_khatri_rao.py:94-109):
//   out[(i_0..i_{m-1}), r] = ((((M_0[i_0,r] * w[r]) * M_1[i_1,r]) * M_2[i_2,r]) ...) * mask[row]
// One IEEE-rounded multiply per step (__fmul_rn/__dmul_rn: never contracted, never
// reassociated).  Output-bandwidth bound: each output element is written once, inputs
// are tiny and L2-resident.
#include "common.cuh"
#include "hf_split.cuh"

namespace tlb200 {
namespace {

template <typename T>
struct KrArgs {
    const T* mat[TLB200_MAX_NDIM];
    int64_t rows[TLB200_MAX_NDIM];
    int64_t rs[TLB200_MAX_NDIM];
    int64_t cs[TLB200_MAX_NDIM];
    int nmats;
};

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// blockDim = (32, 8): x runs over columns, y over rows.
template <typename T>
__global__ void __launch_bounds__(256)
khatri_rao_kernel(KrArgs<T> a, int64_t total_rows, int64_t rank, int64_t pad_cols,
                  const T* __restrict__ weights, const T* __restrict__ mask, T* __restrict__ out,
                  int64_t out_ld) {
    for (int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; row < total_rows;
         row += (int64_t)gridDim.x * blockDim.y) {
        int64_t idx[TLB200_MAX_NDIM];
        int64_t rem = row;
#pragma unroll
        for (int i = TLB200_MAX_NDIM - 1; i >= 0; --i) {
            if (i < a.nmats) {
                int64_t q = rem / a.rows[i];
                idx[i] = rem - q * a.rows[i];
                rem = q;
            }
        }
        T mk = mask ? mask[row] : T(1);
        for (int64_t c = threadIdx.x; c < pad_cols; c += 32) {
            T v = T(0);
            if (c < rank) {
                v = a.mat[0][idx[0] * a.rs[0] + c * a.cs[0]];
                if (weights) v = mul_rn(v, weights[c]);
#pragma unroll
                for (int i = 1; i < TLB200_MAX_NDIM; ++i)
                    if (i < a.nmats) v = mul_rn(v, a.mat[i][idx[i] * a.rs[i] + c * a.cs[i]]);
                if (mask) v = mul_rn(v, mk);
            }
            out[row * out_ld + c] = v;
        }
    }
}

// Large outputs: the row of the result with prefix index p (all matrices but the last, last fastest) and index il
// of the last matrix is  pre[p, :] * F_last[il, :]  where pre = the left fold over the first nmats-1 matrices —
// exactly the reference's association.  A CTA computes PB prefix rows once into shared memory and then streams
// PB x LB output rows: per 16 bytes written one shared-memory load, VW rounded multiplies and one vector store,
// instead of an index decomposition with 64-bit divisions per element.  Threads run along the columns first, so
// every warp writes whole contiguous 512-byte pieces.
template <typename T, int VW>
struct alignas(sizeof(T) * VW) KrVec {
    T v[VW];
};

template <typename T, int VW>
__global__ void __launch_bounds__(256)
khatri_rao_fast_kernel(KrArgs<T> a, int64_t prefix_rows, int64_t last_rows, int rank, int PB, int64_t LB,
                       const T* __restrict__ weights, const T* __restrict__ mask, T* __restrict__ out, int64_t out_ld) {
    extern __shared__ __align__(16) unsigned char kr_smem[];
    T* pre = reinterpret_cast<T*>(kr_smem);          // [PB][rank]
    using V = KrVec<T, VW>;
    const int tid = threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * PB;
    const int npb = (int)min((int64_t)PB, prefix_rows - p0);
    const int last = a.nmats - 1;
    for (int e = tid; e < npb * rank; e += 256) {
        const int pb = e / rank, c = e - pb * rank;
        int64_t rem = p0 + pb;
        int64_t idx[TLB200_MAX_NDIM];
#pragma unroll
        for (int i = TLB200_MAX_NDIM - 2; i >= 0; --i) {
            if (i < last) {
                const int64_t q = rem / a.rows[i];
                idx[i] = rem - q * a.rows[i];
                rem = q;
            }
        }
        T v = a.mat[0][idx[0] * a.rs[0] + c * a.cs[0]];
        if (weights) v = mul_rn(v, weights[c]);
#pragma unroll
        for (int i = 1; i < TLB200_MAX_NDIM - 1; ++i)
            if (i < last) v = mul_rn(v, a.mat[i][idx[i] * a.rs[i] + c * a.cs[i]]);
        pre[pb * rank + c] = v;
    }
    __syncthreads();
    const int RV = rank / VW, NR = 256 / RV;
    const int rv = tid % RV, rl = tid / RV;
    if (rl >= NR) return;
    const T* fl_base = a.mat[last];
    const int64_t frs = a.rs[last], fcs = a.cs[last];
    const int64_t l_begin = (int64_t)blockIdx.y * LB;
    const int64_t l_end = min(last_rows, l_begin + LB);
    for (int64_t il = l_begin + rl; il < l_end; il += NR) {
        T fl[VW];
#pragma unroll
        for (int c = 0; c < VW; ++c) fl[c] = fl_base[il * frs + (int64_t)(rv * VW + c) * fcs];
#pragma unroll 4
        for (int pb = 0; pb < npb; ++pb) {
            const int64_t row = (p0 + pb) * last_rows + il;
            const V pv = *reinterpret_cast<const V*>(pre + pb * rank + rv * VW);
            V o;
#pragma unroll
            for (int c = 0; c < VW; ++c) o.v[c] = mul_rn(pv.v[c], fl[c]);
            if (mask) {
                const T mk = mask[row];
#pragma unroll
                for (int c = 0; c < VW; ++c) o.v[c] = mul_rn(o.v[c], mk);
            }
            *reinterpret_cast<V*>(out + row * out_ld + rv * VW) = o;
        }
    }
}

// Transposed, zero-padded variant for the tcgen05 engine: out[c * out_ld + row] for row < rows_padded
// (rows >= total_rows and columns >= rank are written as zero); with out_lo the value is split into its
// tf32 truncation (out) and the exact remainder (out_lo).  blockDim = (32, 8): x runs over rows
// so that the stores are coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
khatri_rao_t_kernel(KrArgs<T> a, int64_t total_rows, int64_t rows_padded, int64_t rank, int64_t pad_cols,
                    const T* __restrict__ weights, T* __restrict__ out, int64_t out_ld, T* __restrict__ out_lo) {
    for (int64_t row = (int64_t)blockIdx.x * 32 + threadIdx.x; row < rows_padded; row += (int64_t)gridDim.x * 32) {
        int64_t idx[TLB200_MAX_NDIM];
        int64_t rem = row < total_rows ? row : 0;
#pragma unroll
        for (int i = TLB200_MAX_NDIM - 1; i >= 0; --i) {
            if (i < a.nmats) {
                int64_t q = rem / a.rows[i];
                idx[i] = rem - q * a.rows[i];
                rem = q;
            }
        }
        for (int64_t c = threadIdx.y; c < pad_cols; c += 8) {
            T v = T(0);
            if (c < rank && row < total_rows) {
                v = a.mat[0][idx[0] * a.rs[0] + c * a.cs[0]];
                if (weights) v = mul_rn(v, weights[c]);
#pragma unroll
                for (int i = 1; i < TLB200_MAX_NDIM; ++i)
                    if (i < a.nmats) v = mul_rn(v, a.mat[i][idx[i] * a.rs[i] + c * a.cs[i]]);
            }
            if (out_lo) {     // tf32 split for the tensor-core engine: hi = truncation, lo = exact remainder
                const float f = (float)v;
                const float h = __uint_as_float(__float_as_uint(f) & 0xFFFFE000u);
                out[c * out_ld + row] = (T)h;
                out_lo[c * out_ld + row] = (T)(f - h);
            } else {
                out[c * out_ld + row] = v;
            }
        }
    }
}

// fp16-split variant for the tensor-core engine's fp16 sibling: one CTA per column c finds the column's own
// power-of-two scale (pass 1: max |v|), then writes hi / lo halves (pass 2).  The products are formed twice — the
// table is tiny next to the tensor pass it feeds.
__global__ void __launch_bounds__(256)
khatri_rao_t_f16_kernel(KrArgs<float> a, int64_t total_rows, int64_t rows_padded, int64_t rank,
                        const float* __restrict__ weights, __half* __restrict__ hi, __half* __restrict__ lo,
                        float* __restrict__ col_inv) {
    const int64_t c = blockIdx.x;
    auto value = [&](int64_t row) {
        int64_t rem = row;
        float v = 1.f;
        bool first = true;
        // factor 0 is the slowest index: peel from the last matrix; multiply in the reference's order (0, 1, ...)
        int64_t idx[TLB200_MAX_NDIM];
#pragma unroll
        for (int i = TLB200_MAX_NDIM - 1; i >= 0; --i) {
            if (i < a.nmats) {
                const int64_t q = rem / a.rows[i];
                idx[i] = rem - q * a.rows[i];
                rem = q;
            }
        }
#pragma unroll
        for (int i = 0; i < TLB200_MAX_NDIM; ++i) {
            if (i < a.nmats) {
                const float f = a.mat[i][idx[i] * a.rs[i] + c * a.cs[i]];
                if (first) { v = weights ? mul_rn(f, weights[c]) : f; first = false; }
                else v = mul_rn(v, f);
            }
        }
        return v;
    };
    unsigned mx = 0;
    if (c < rank)
        for (int64_t row = threadIdx.x; row < total_rows; row += 256) mx = max(mx, __float_as_uint(fabsf(value(row))));
    float inv;
    const float sc = hf_block_scale(mx, &inv);
    if (threadIdx.x == 0) col_inv[c] = inv;
    for (int64_t row = threadIdx.x; row < rows_padded; row += 256) {
        const float v = (c < rank && row < total_rows) ? value(row) * sc : 0.f;
        __half h, l;
        hf_split1(v, h, l);
        hi[c * rows_padded + row] = h;
        lo[c * rows_padded + row] = l;
    }
}

}  // namespace

int launch_khatri_rao_t_f16(const float* const* mats, const int64_t* rows, const int64_t* row_stride,
                            const int64_t* col_stride, int nmats, int64_t rank, const float* weights, __half* hi,
                            __half* lo, int64_t rows_padded, int64_t pad_cols, float* col_inv, cudaStream_t stream) {
    if (nmats < 1 || nmats > TLB200_MAX_NDIM || rank < 0 || pad_cols < rank) return TLB200_EINVAL;
    KrArgs<float> a;
    a.nmats = nmats;
    int64_t total = 1;
    for (int i = 0; i < nmats; ++i) {
        a.mat[i] = mats[i]; a.rows[i] = rows[i]; a.rs[i] = row_stride[i]; a.cs[i] = col_stride[i];
        total *= rows[i];
    }
    for (int i = nmats; i < TLB200_MAX_NDIM; ++i) { a.mat[i] = nullptr; a.rows[i] = 1; a.rs[i] = 0; a.cs[i] = 0; }
    if (rows_padded < total) return TLB200_EINVAL;
    if (rows_padded == 0 || pad_cols == 0) return TLB200_OK;
    khatri_rao_t_f16_kernel<<<(unsigned)pad_cols, 256, 0, stream>>>(a, total, rows_padded, rank, weights, hi, lo, col_inv);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <typename T>
int launch_khatri_rao_t(const T* const* mats, const int64_t* rows, const int64_t* row_stride, const int64_t* col_stride,
                        int nmats, int64_t rank, const T* weights, T* out, int64_t rows_padded, int64_t pad_cols,
                        T* out_lo, cudaStream_t stream) {
    if (nmats < 1 || nmats > TLB200_MAX_NDIM || rank < 0 || pad_cols < rank) return TLB200_EINVAL;
    KrArgs<T> a;
    a.nmats = nmats;
    int64_t total = 1;
    for (int i = 0; i < nmats; ++i) {
        a.mat[i] = mats[i]; a.rows[i] = rows[i]; a.rs[i] = row_stride[i]; a.cs[i] = col_stride[i];
        total *= rows[i];
    }
    for (int i = nmats; i < TLB200_MAX_NDIM; ++i) { a.mat[i] = nullptr; a.rows[i] = 1; a.rs[i] = 0; a.cs[i] = 0; }
    if (rows_padded < total) return TLB200_EINVAL;
    if (rows_padded == 0 || pad_cols == 0) return TLB200_OK;
    int64_t blocks = ceil_div(rows_padded, 32);
    if (blocks > (int64_t)kNumSMs * 64) blocks = (int64_t)kNumSMs * 64;
    khatri_rao_t_kernel<T><<<(unsigned)blocks, dim3(32, 8), 0, stream>>>(a, total, rows_padded, rank, pad_cols, weights, out, rows_padded, out_lo);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}
template int launch_khatri_rao_t<float>(const float* const*, const int64_t*, const int64_t*, const int64_t*, int, int64_t,
                                        const float*, float*, int64_t, int64_t, float*, cudaStream_t);

template <typename T>
int launch_khatri_rao(const T* const* mats, const int64_t* rows, const int64_t* row_stride,
                      const int64_t* col_stride, int nmats, int64_t rank, const T* weights,
                      const T* mask, T* out, int64_t out_ld, int64_t pad_cols, cudaStream_t stream) {
    if (nmats < 1 || nmats > TLB200_MAX_NDIM || rank < 0 || pad_cols < rank || out_ld < pad_cols)
        return TLB200_EINVAL;
    KrArgs<T> a;
    a.nmats = nmats;
    int64_t total = 1;
    for (int i = 0; i < nmats; ++i) {
        if (rows[i] < 0) return TLB200_EINVAL;
        a.mat[i] = mats[i];
        a.rows[i] = rows[i];
        a.rs[i] = row_stride[i];
        a.cs[i] = col_stride[i];
        total *= rows[i];
    }
    for (int i = nmats; i < TLB200_MAX_NDIM; ++i) { a.mat[i] = nullptr; a.rows[i] = 1; a.rs[i] = 0; a.cs[i] = 0; }
    if (total == 0 || pad_cols == 0) return TLB200_OK;
    {   // streaming kernel for large, vectorisable outputs (everything the public khatri_rao is used for at scale)
        constexpr int VW = sizeof(T) == 4 ? 4 : 2;
        const int64_t last_rows = rows[nmats - 1];
        if (nmats >= 2 && pad_cols == rank && rank % VW == 0 && rank / VW <= 256 && out_ld % VW == 0 &&
            reinterpret_cast<uintptr_t>(out) % 16 == 0 && last_rows >= 16 && total >= 4096 && last_rows > 0) {
            const int64_t prefix_rows = total / last_rows;
            const int PB = 8;
            const int64_t LB = 512;
            const int64_t gx = ceil_div(prefix_rows, PB), gy = ceil_div(last_rows, LB);
            if (gx <= 0x7fffffffLL && gy <= 65535) {
                const size_t smem = sizeof(T) * PB * (size_t)rank;
                khatri_rao_fast_kernel<T, VW><<<dim3((unsigned)gx, (unsigned)gy), 256, smem, stream>>>(
                    a, prefix_rows, last_rows, (int)rank, PB, LB, weights, mask, out, out_ld);
                TLB_CHECK_LAUNCH();
                return TLB200_OK;
            }
        }
    }
    int64_t blocks = ceil_div(total, 8);
    if (blocks > (int64_t)kNumSMs * 64) blocks = (int64_t)kNumSMs * 64;
    khatri_rao_kernel<T><<<(unsigned)blocks, dim3(32, 8), 0, stream>>>(a, total, rank, pad_cols, weights, mask, out, out_ld);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template int launch_khatri_rao<float>(const float* const*, const int64_t*, const int64_t*, const int64_t*, int, int64_t,
                                      const float*, const float*, float*, int64_t, int64_t, cudaStream_t);
template int launch_khatri_rao<double>(const double* const*, const int64_t*, const int64_t*, const int64_t*, int, int64_t,
                                       const double*, const double*, double*, int64_t, int64_t, cudaStream_t);

}  // namespace tlb200

using namespace tlb200;

extern "C" int tlb200_khatri_rao(const void* const* mats, const int64_t* rows, const int64_t* row_stride,
                                 const int64_t* col_stride, int nmats, int64_t rank, const void* weights,
                                 const void* mask, int dtype, void* out, int64_t out_ld, void* stream) {
    if (!mats || !rows || !row_stride || !col_stride || !dtype_valid(dtype) || nmats < 1 || nmats > TLB200_MAX_NDIM ||
        rank < 0 || out_ld < rank)
        return TLB200_EINVAL;
    for (int i = 0; i < nmats; ++i)
        if (!mats[i] && rows[i] * rank) return TLB200_EINVAL;
    set_last_path("khatri_rao");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return launch_khatri_rao<float>(reinterpret_cast<const float* const*>(mats), rows, row_stride, col_stride, nmats,
                                        rank, (const float*)weights, (const float*)mask, (float*)out, out_ld, rank, s);
    return launch_khatri_rao<double>(reinterpret_cast<const double* const*>(mats), rows, row_stride, col_stride, nmats,
                                     rank, (const double*)weights, (const double*)mask, (double*)out, out_ld, rank, s);
}
