// CP reconstruction and masked-ALS imputation.
//
//   tlb200_cp_to_tensor : out[i_0..i_{N-1}] = sum_r w_r * prod_n F_n[i_n, r]
//       reference: tensorly/cp_tensor.py:433-485 — materialises the Khatri-Rao matrix of modes 1..N-1
//       (prod I_n x R), one GEMM, then fold.  Here the Khatri-Rao rows are formed per column tile in shared
//       memory and the tensor is written exactly once.
//   tlb200_cp_impute    : out = x * mask + rec * (1 - mask),  sums: ||out||^2 and ||out - rec||^2
//       reference: the masked branch of error_calc, tensorly/decomposition/_cp.py:195-207 — cp_to_tensor
//       (a tensor-sized temporary), the blend (three more), tl.norm twice.  Here: one pass that reads x and
//       mask, forms rec in registers, writes the imputed tensor and accumulates both norms; rec never
//       exists in memory.
//
// Kernel (this file: fp64, rank > 64, small problems): out viewed as a matrix [I_0][C], C = prod_{n>=1} I_n
// (contiguous), computed as a rank-R outer-product GEMM with 128 x 128 tiles, 8 x 8 register micro-tiles, operands in
// shared memory ([r][row] / [r][col], rank chunks of 32).  Arithmetic intensity is 2R flop per 4 (plain) or 12
// (impute) bytes: FMA-bound on the CUDA cores at R >= 16 — large fp32 problems of rank <= 64 therefore take the
// tensor-core kernel of recon_tc.cu (3.2x / 2.2x this one at 512 x 1024 x 1024, rank 32).
#include "common.cuh"
#include "recon_tc.cuh"

namespace tlb200 {
namespace {

constexpr int RTI = 128;       // tile rows (mode 0)
constexpr int RTC = 128;       // tile columns (modes 1.., contiguous)
template <typename T> struct RChunk { static constexpr int value = 128 / sizeof(T); };   // rank chunk: 32 fp32 / 16 fp64
constexpr int RLD = RTI + 4;   // padded leading dimension of the operand tiles (keeps 16-byte alignment)
constexpr int RTHREADS = 256;

template <typename T> struct RVec;
template <> struct RVec<float> { using type = float4; static constexpr int W = 4; };
template <> struct RVec<double> { using type = double2; static constexpr int W = 2; };

struct ReconGeom {
    int ndim;
    int64_t shape[TLB200_MAX_NDIM];
    int64_t rs[TLB200_MAX_NDIM], cs[TLB200_MAX_NDIM];
    const void* f[TLB200_MAX_NDIM];
    int64_t I, C;
};

// MODE 0: out = rec.   MODE 1: out = x*mask + rec*(1-mask), partial[blk] = {sum out^2, sum (out-rec)^2}.
// MODE 2: out = rec * mask.
template <typename T, int MODE>
__global__ void __launch_bounds__(RTHREADS)
recon_kernel(const ReconGeom g, int R, const T* __restrict__ w, const T* __restrict__ x, const T* __restrict__ mask,
             T* __restrict__ out, double* __restrict__ partial) {
    constexpr int VW = RVec<T>::W;
    using Vec = typename RVec<T>::type;
    constexpr int NG = 8 / VW;                   // column groups per thread
    constexpr int RKC = RChunk<T>::value;
    __shared__ __align__(16) T As[RKC * RLD];    // [r][row]
    __shared__ __align__(16) T Ks[RKC * RLD];    // [r][col]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t c0 = (int64_t)blockIdx.x * RTC;
    const int64_t i0 = (int64_t)blockIdx.y * RTI;

    // this thread's column for the Khatri-Rao tile: decompose once (last mode fastest)
    const int kc = tid & (RTC - 1), kh = tid >> 7;           // column, rank half (16 ranks each)
    const T* kptr[TLB200_MAX_NDIM];
    int64_t kcs[TLB200_MAX_NDIM];
    const bool kvalid = c0 + kc < g.C;
    {
        int64_t rem = kvalid ? c0 + kc : 0;
        for (int n = g.ndim - 1; n >= 1; --n) {
            const int64_t idx = rem % g.shape[n];
            rem /= g.shape[n];
            kptr[n] = static_cast<const T*>(g.f[n]) + idx * g.rs[n];
            kcs[n] = g.cs[n];
        }
    }
    const T* f0 = static_cast<const T*>(g.f[0]);

    T acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = T(0);

    for (int r0 = 0; r0 < R; r0 += RKC) {
        const int rk = min(RKC, R - r0);
        __syncthreads();
        // A tile: As[r][row] = F0[i0 + row, r0 + r] * w[r0 + r]   (global reads contiguous in r)
        for (int e = tid; e < RTI * RKC; e += RTHREADS) {
            const int r = e % RKC, row = e / RKC;
            T v = T(0);
            if (r < rk && i0 + row < g.I) {
                v = f0[(i0 + row) * g.rs[0] + (int64_t)(r0 + r) * g.cs[0]];
                if (w) v *= w[r0 + r];
            }
            As[r * RLD + row] = v;
        }
        // Khatri-Rao tile: Ks[r][col] = prod_{n>=1} F_n[idx_n(col), r0 + r]
#pragma unroll 4
        for (int rr = 0; rr < RKC / 2; ++rr) {
            const int r = kh * (RKC / 2) + rr;
            T v = T(0);
            if (kvalid && r < rk) {
                v = kptr[1][(int64_t)(r0 + r) * kcs[1]];
                for (int n = 2; n < g.ndim; ++n) v *= kptr[n][(int64_t)(r0 + r) * kcs[n]];
            }
            Ks[r * RLD + kc] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < RKC; ++r) {
            T av[8], kv[8];
#pragma unroll
            for (int h = 0; h < 8 / VW; ++h) {
                const Vec v = *reinterpret_cast<const Vec*>(&As[r * RLD + ty * 8 + h * VW]);
                if constexpr (VW == 4) { av[4 * h] = v.x; av[4 * h + 1] = v.y; av[4 * h + 2] = v.z; av[4 * h + 3] = v.w; }
                else { av[2 * h] = v.x; av[2 * h + 1] = v.y; }
            }
#pragma unroll
            for (int gq = 0; gq < NG; ++gq) {
                const Vec v = *reinterpret_cast<const Vec*>(&Ks[r * RLD + gq * (16 * VW) + tx * VW]);
                if constexpr (VW == 4) { kv[4 * gq] = v.x; kv[4 * gq + 1] = v.y; kv[4 * gq + 2] = v.z; kv[4 * gq + 3] = v.w; }
                else { kv[2 * gq] = v.x; kv[2 * gq + 1] = v.y; }
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fma(av[a], kv[b], acc[a][b]);
        }
    }

    // epilogue
    const bool vec_ok = (g.C % VW) == 0 && (reinterpret_cast<uintptr_t>(out) % sizeof(Vec)) == 0 &&
                        (MODE != 1 || (reinterpret_cast<uintptr_t>(x) % sizeof(Vec)) == 0) &&
                        (MODE == 0 || (reinterpret_cast<uintptr_t>(mask) % sizeof(Vec)) == 0);
    double s_out = 0.0, s_res = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t gi = i0 + ty * 8 + a;
        if (gi >= g.I) continue;
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
            const int64_t gc = c0 + gq * (16 * VW) + tx * VW;
            if (gc >= g.C) continue;
            const int64_t off = gi * g.C + gc;
            T v[VW];
#pragma unroll
            for (int b = 0; b < VW; ++b) v[b] = acc[a][gq * VW + b];
            if (vec_ok && gc + VW <= g.C) {
                if constexpr (MODE == 1) {
                    const Vec xv = *reinterpret_cast<const Vec*>(x + off);
                    const Vec mv = *reinterpret_cast<const Vec*>(mask + off);
                    T xs[VW], ms[VW];
                    if constexpr (VW == 4) { xs[0] = xv.x; xs[1] = xv.y; xs[2] = xv.z; xs[3] = xv.w; ms[0] = mv.x; ms[1] = mv.y; ms[2] = mv.z; ms[3] = mv.w; }
                    else { xs[0] = xv.x; xs[1] = xv.y; ms[0] = mv.x; ms[1] = mv.y; }
#pragma unroll
                    for (int b = 0; b < VW; ++b) {
                        const T o = xs[b] * ms[b] + v[b] * (T(1) - ms[b]);       // the reference's expression
                        const double d = (double)o - (double)v[b];
                        s_out += (double)o * (double)o;
                        s_res += d * d;
                        v[b] = o;
                    }
                } else if constexpr (MODE == 2) {
                    const Vec mv = *reinterpret_cast<const Vec*>(mask + off);
                    if constexpr (VW == 4) { v[0] *= mv.x; v[1] *= mv.y; v[2] *= mv.z; v[3] *= mv.w; }
                    else { v[0] *= mv.x; v[1] *= mv.y; }
                }
                Vec ov;
                if constexpr (VW == 4) { ov.x = v[0]; ov.y = v[1]; ov.z = v[2]; ov.w = v[3]; }
                else { ov.x = v[0]; ov.y = v[1]; }
                *reinterpret_cast<Vec*>(out + off) = ov;
            } else {
#pragma unroll
                for (int b = 0; b < VW; ++b) {
                    if (gc + b >= g.C) continue;
                    T o = v[b];
                    if constexpr (MODE == 1) {
                        const T xs = x[off + b], ms = mask[off + b];
                        o = xs * ms + v[b] * (T(1) - ms);
                        const double d = (double)o - (double)v[b];
                        s_out += (double)o * (double)o;
                        s_res += d * d;
                    } else if constexpr (MODE == 2) {
                        o = v[b] * mask[off + b];
                    }
                    out[off + b] = o;
                }
            }
        }
    }
    if constexpr (MODE == 1) {
        __shared__ double red[2][RTHREADS / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s_out += __shfl_xor_sync(0xffffffffu, s_out, o);
            s_res += __shfl_xor_sync(0xffffffffu, s_res, o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = s_out; red[1][tid >> 5] = s_res; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < RTHREADS / 32; ++i) { a += red[0][i]; b += red[1][i]; }
            const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
            partial[2 * blk] = a;
            partial[2 * blk + 1] = b;
        }
    }
}

// stats[0] = sqrt(sum (out-rec)^2) / sqrt(sum out^2)   (the masked rec_error of _cp.py:205 + :478)
// stats[1] = sum out^2 (the new ||tensor||^2),  stats[2] = sum (out-rec)^2; fixed summation order.
template <typename T>
__global__ void __launch_bounds__(256)
impute_finish_kernel(const double* __restrict__ partial, int64_t n, T* __restrict__ stats) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[0] = (T)(sqrt(sb[0]) / sqrt(sa[0]));
        stats[1] = (T)sa[0];
        stats[2] = (T)sb[0];
    }
}

int make_geom(const void* const* factors, const int64_t* shape, const int64_t* frs, const int64_t* fcs, int ndim,
              int64_t rank, ReconGeom* g) {
    if (!factors || !shape || !frs || !fcs || ndim < 2 || ndim > TLB200_MAX_NDIM || rank < 1) return TLB200_EINVAL;
    g->ndim = ndim;
    g->I = shape[0];
    g->C = 1;
    for (int n = 0; n < ndim; ++n) {
        if (shape[n] < 1 || !factors[n]) return TLB200_EINVAL;
        g->shape[n] = shape[n]; g->rs[n] = frs[n]; g->cs[n] = fcs[n]; g->f[n] = factors[n];
        if (n >= 1) g->C *= shape[n];
    }
    for (int n = ndim; n < TLB200_MAX_NDIM; ++n) { g->shape[n] = 1; g->rs[n] = 0; g->cs[n] = 0; g->f[n] = nullptr; }
    if (ceil_div(g->I, RTI) > 65535 || ceil_div(g->C, RTC) >= (1LL << 31)) return TLB200_EUNSUPPORTED;
    return TLB200_OK;
}

template <typename T>
int launch(const ReconGeom& g, int64_t rank, const T* w, const T* x, const T* mask, T* out, double* partial, T* stats,
           cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(g.C, RTC), (unsigned)ceil_div(g.I, RTI));
    if (x == nullptr) {
        if (mask == nullptr) recon_kernel<T, 0><<<grid, RTHREADS, 0, stream>>>(g, (int)rank, w, nullptr, nullptr, out, nullptr);
        else recon_kernel<T, 2><<<grid, RTHREADS, 0, stream>>>(g, (int)rank, w, nullptr, mask, out, nullptr);
        TLB_CHECK_LAUNCH();
        return TLB200_OK;
    }
    recon_kernel<T, 1><<<grid, RTHREADS, 0, stream>>>(g, (int)rank, w, x, mask, out, partial);
    TLB_CHECK_LAUNCH();
    impute_finish_kernel<T><<<1, 256, 0, stream>>>(partial, (int64_t)grid.x * grid.y, stats);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

}  // namespace
}  // namespace tlb200

using namespace tlb200;

extern "C" int tlb200_cp_to_tensor(const void* const* factors, const int64_t* shape, const int64_t* f_row_stride,
                                   const int64_t* f_col_stride, int ndim, int64_t rank, const void* weights,
                                   const void* mask, int dtype, void* out, void* stream) {
    ReconGeom g;
    if (!dtype_valid(dtype) || !out) return TLB200_EINVAL;
    int st = make_geom(factors, shape, f_row_stride, f_col_stride, ndim, rank, &g);
    if (st) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (recon_tc_supported(shape, ndim, rank, dtype)) {
        set_last_path("tcgen05");
        return recon_tc_launch(factors, shape, f_row_stride, f_col_stride, ndim, rank, (const float*)weights, nullptr,
                               (const float*)mask, (float*)out, nullptr, s);
    }
    set_last_path("simt");
    if (dtype == TLB200_F32)
        return launch<float>(g, rank, (const float*)weights, nullptr, (const float*)mask, (float*)out, nullptr, nullptr, s);
    return launch<double>(g, rank, (const double*)weights, nullptr, (const double*)mask, (double*)out, nullptr, nullptr, s);
}

extern "C" size_t tlb200_cp_impute_workspace_bytes(const int64_t* shape, int ndim) {
    if (!shape || ndim < 2 || ndim > TLB200_MAX_NDIM) return 0;
    int64_t C = 1;
    for (int n = 1; n < ndim; ++n) C *= shape[n];
    return align_up((size_t)(ceil_div(shape[0], RTI) * ceil_div(C, RTC)) * 2 * sizeof(double), 256) + 256;
}

extern "C" int tlb200_cp_impute(const void* x, const void* mask, const void* const* factors, const int64_t* shape,
                                const int64_t* f_row_stride, const int64_t* f_col_stride, int ndim, int64_t rank,
                                const void* weights, int dtype, void* out, void* stats, void* workspace,
                                size_t workspace_bytes, void* stream) {
    ReconGeom g;
    if (!dtype_valid(dtype) || !out || !x || !mask || !stats || !workspace) return TLB200_EINVAL;
    int st = make_geom(factors, shape, f_row_stride, f_col_stride, ndim, rank, &g);
    if (st) return st;
    if (workspace_bytes < tlb200_cp_impute_workspace_bytes(shape, ndim)) return TLB200_EWORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(workspace);
    if (recon_tc_supported(shape, ndim, rank, dtype)) {
        set_last_path("tcgen05");
        st = recon_tc_launch(factors, shape, f_row_stride, f_col_stride, ndim, rank, (const float*)weights,
                             (const float*)x, (const float*)mask, (float*)out, partial, s);
        if (st) return st;
        impute_finish_kernel<float><<<1, 256, 0, s>>>(partial, (int64_t)recon_tc_grid(shape, ndim), (float*)stats);
        TLB_CHECK_LAUNCH();
        return TLB200_OK;
    }
    set_last_path("simt");
    if (dtype == TLB200_F32)
        return launch<float>(g, rank, (const float*)weights, (const float*)x, (const float*)mask, (float*)out, partial,
                             (float*)stats, s);
    return launch<double>(g, rank, (const double*)weights, (const double*)x, (const double*)mask, (double*)out, partial,
                          (double*)stats, s);
}
