// mode_dot (TTM) and multi_mode_dot (TTM chain).
//
// Reference: tensorly/tenalg/core_tenalg/n_mode_product.py:5-76 computes
// fold(dot(M, unfold(X, mode))) — a permuting copy of X for middle modes, a GEMM and a
// non-contiguous fold view.  Here X is viewed in place as X[L, J, T] (J = shape[mode]) and
//   out[l, i, t] = sum_j M[i, j] * X[l, j, t]
// is computed directly into a C-contiguous out[L, I, T] with the tensor streamed once:
//   T >= 32 : per l, stream rows of T contiguous elements (streamed dim = t)
//   T == 1  : one GEMM with streamed dim = l, contraction contiguous
//   else    : streamed dim = l, one batch per t (strided loads; rare small-T case)
#include "common.cuh"
#include "stream_gemm.cuh"
#include "ttm_tc.cuh"

namespace tlb200 {

struct TtmDims { int64_t L, J, T; };

static int ttm_dims(const int64_t* shape, int ndim, int mode, TtmDims* d) {
    if (!shape || ndim < 1 || ndim > TLB200_MAX_NDIM || mode < 0 || mode >= ndim) return TLB200_EINVAL;
    d->L = 1; d->T = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] < 1) return TLB200_EINVAL;
        if (i < mode) d->L *= shape[i];
        if (i > mode) d->T *= shape[i];
    }
    d->J = shape[mode];
    return TLB200_OK;
}

template <typename T>
static int ttm_simt(const T* x, const TtmDims& d, const T* m, int64_t I, int64_t mrs, int64_t mcs, T* out,
                    cudaStream_t stream) {
    const int dtype = sizeof(T) == 8 ? TLB200_F64 : TLB200_F32;
    const int KT = dtype == TLB200_F64 ? 16 : 32;
    StreamGemmParams<T> p;
    p.X = x;
    p.KA = 1; p.KB = d.J; p.sXa = 0;
    p.P = nullptr; p.ldP = 0;
    p.Q = m; p.sQb = mcs; p.sQn = mrs;   // B(b=j, n=i) = M[i, j]
    p.N = I;
    p.C = out; p.sCsplit = 0;
    p.chunks_per_a = ceil_div(d.J, KT);
    p.total_chunks = p.chunks_per_a;
    p.chunks_per_split = p.total_chunks;
    p.nsplit = 1;
    bool kmajor;
    if (d.T >= 32) {
        p.M = d.T; p.sXm = 1; p.sXb = d.T; p.sXbatch = d.J * d.T; p.nbatch = d.L;
        p.sCm = 1; p.sCn = d.T; p.sCbatch = I * d.T;
        kmajor = false;
    } else if (d.T == 1) {
        p.M = d.L; p.sXm = d.J; p.sXb = 1; p.sXbatch = 0; p.nbatch = 1;
        p.sCm = I; p.sCn = 1; p.sCbatch = 0;
        kmajor = true;
    } else {
        p.M = d.L; p.sXm = d.J * d.T; p.sXb = d.T; p.sXbatch = 1; p.nbatch = d.T;
        p.sCm = I * d.T; p.sCn = d.T; p.sCbatch = 1;
        kmajor = true;
    }
    return launch_stream_gemm<T>(p, stream_gemm_tr_for(I, dtype), kmajor, stream);
}

// workspace one mode_dot needs (tcgen05 path: the pre-split hi/lo copies of the matrix)
static size_t mode_dot_ws(const TtmDims& d, int64_t I, int dtype, int path) {
    if (path != TLB200_PATH_SIMT && dtype == TLB200_F32 && ttm_tc_supported(d.L, d.J, d.T, I))
        return ttm_tc_workspace(d.L, d.J, d.T, I);
    return 0;
}

template <typename T>
static int mode_dot_impl(const T* x, const int64_t* shape, int ndim, int mode, const T* m, int64_t I, int64_t mrs,
                         int64_t mcs, T* out, void* ws, size_t ws_bytes, int path, cudaStream_t stream) {
    TtmDims d;
    int st = ttm_dims(shape, ndim, mode, &d);
    if (st) return st;
    if (path != TLB200_PATH_SIMT && sizeof(T) == 4 && ttm_tc_supported(d.L, d.J, d.T, I) &&
        reinterpret_cast<uintptr_t>(x) % 16 == 0) {
        if (!ws || ws_bytes < ttm_tc_workspace(d.L, d.J, d.T, I)) return TLB200_EWORKSPACE;
        set_last_path("tcgen05");
        return ttm_tc_launch(reinterpret_cast<const float*>(x), d.L, d.J, d.T, reinterpret_cast<const float*>(m), I,
                             mrs, mcs, reinterpret_cast<float*>(out), ws, stream);
    }
    if (path == TLB200_PATH_TCGEN05) return TLB200_EUNSUPPORTED;
    set_last_path("simt");
    return ttm_simt<T>(x, d, m, I, mrs, mcs, out, stream);
}

}  // namespace tlb200

using namespace tlb200;

extern "C" size_t tlb200_mode_dot_workspace_bytes(const int64_t* shape, int ndim, int mode, int64_t rows_out, int dtype,
                                                  int path) {
    TtmDims d;
    if (ttm_dims(shape, ndim, mode, &d) || !dtype_valid(dtype) || rows_out < 1) return 0;
    return mode_dot_ws(d, rows_out, dtype, path);
}

extern "C" int tlb200_mode_dot(const void* x, const int64_t* shape, int ndim, int mode, const void* m, int64_t rows_out,
                               int64_t m_row_stride, int64_t m_col_stride, int dtype, void* out, void* workspace,
                               size_t workspace_bytes, int path, void* stream) {
    if (!x || !m || !out || rows_out < 1 || !dtype_valid(dtype) || path < TLB200_PATH_AUTO || path > TLB200_PATH_TCGEN05)
        return TLB200_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return mode_dot_impl<float>((const float*)x, shape, ndim, mode, (const float*)m, rows_out, m_row_stride,
                                    m_col_stride, (float*)out, workspace, workspace_bytes, path, s);
    return mode_dot_impl<double>((const double*)x, shape, ndim, mode, (const double*)m, rows_out, m_row_stride,
                                 m_col_stride, (double*)out, workspace, workspace_bytes, path, s);
}

// ---- chain -------------------------------------------------------------------------
static int chain_check(const int64_t* shape, int ndim, const int* modes, const int64_t* rows_out, int nmats) {
    if (!shape || ndim < 1 || ndim > TLB200_MAX_NDIM || nmats < 0 || nmats > ndim || (nmats && (!modes || !rows_out)))
        return TLB200_EINVAL;
    for (int k = 0; k < nmats; ++k) {
        if (modes[k] < 0 || modes[k] >= ndim || rows_out[k] < 1) return TLB200_EINVAL;
        if (k && modes[k] <= modes[k - 1]) return TLB200_EINVAL;  // distinct, ascending
    }
    return TLB200_OK;
}

// Largest intermediate (elements) of the chain evaluated in the given order.
static int64_t chain_max_intermediate(const int64_t* shape, int ndim, const int* modes, const int64_t* rows_out,
                                      int nmats) {
    int64_t cur[TLB200_MAX_NDIM];
    for (int i = 0; i < ndim; ++i) cur[i] = shape[i];
    int64_t mx = 0;
    for (int k = 0; k + 1 < nmats; ++k) {
        cur[modes[k]] = rows_out[k];
        int64_t n = 1;
        for (int i = 0; i < ndim; ++i) n *= cur[i];
        if (n > mx) mx = n;
    }
    return mx;
}

// largest per-step scratch (tcgen05 steps need the pre-split matrix)
static size_t chain_step_ws(const int64_t* shape, int ndim, const int* modes, const int64_t* rows_out, int nmats, int dtype,
                            int path) {
    int64_t cur[TLB200_MAX_NDIM];
    for (int i = 0; i < ndim; ++i) cur[i] = shape[i];
    size_t mx = 0;
    for (int k = 0; k < nmats; ++k) {
        TtmDims d;
        if (ttm_dims(cur, ndim, modes[k], &d)) return 0;
        const size_t w = mode_dot_ws(d, rows_out[k], dtype, path);
        if (w > mx) mx = w;
        cur[modes[k]] = rows_out[k];
    }
    return align_up(mx, 256);
}

extern "C" size_t tlb200_multi_mode_dot_workspace_bytes(const int64_t* shape, int ndim, const int* modes,
                                                        const int64_t* rows_out, int nmats, int dtype, int path) {
    if (chain_check(shape, ndim, modes, rows_out, nmats) || !dtype_valid(dtype)) return 0;
    const int64_t mx = chain_max_intermediate(shape, ndim, modes, rows_out, nmats);
    return 2 * align_up((size_t)mx * dtype_size(dtype), 256) + chain_step_ws(shape, ndim, modes, rows_out, nmats, dtype, path) + 512;
}

template <typename T>
static int chain_impl(const T* x, const int64_t* shape, int ndim, const int* modes, const T* const* mats,
                      const int64_t* rows_out, const int64_t* mrs, const int64_t* mcs, int nmats, T* out,
                      void* workspace, int path, cudaStream_t stream) {
    const int64_t mx = chain_max_intermediate(shape, ndim, modes, rows_out, nmats);
    const int dtype = sizeof(T) == 8 ? TLB200_F64 : TLB200_F32;
    const size_t step_ws = chain_step_ws(shape, ndim, modes, rows_out, nmats, dtype, path);
    Carver ws(workspace);
    T* buf[2] = {ws.take<T>((size_t)mx), ws.take<T>((size_t)mx)};
    void* scratch = step_ws ? ws.take<char>(step_ws) : nullptr;
    int64_t cur[TLB200_MAX_NDIM];
    int64_t total = 1;
    for (int i = 0; i < ndim; ++i) { cur[i] = shape[i]; total *= shape[i]; }
    if (nmats == 0) {
        set_last_path("copy");
        if (cudaMemcpyAsync(out, x, sizeof(T) * total, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) return TLB200_ECUDA;
        return TLB200_OK;
    }
    const T* src = x;
    for (int k = 0; k < nmats; ++k) {
        T* dst = (k == nmats - 1) ? out : buf[k & 1];
        int st = mode_dot_impl<T>(src, cur, ndim, modes[k], mats[k], rows_out[k], mrs[k], mcs[k], dst, scratch, step_ws, path,
                                  stream);
        if (st) return st;
        cur[modes[k]] = rows_out[k];
        src = dst;
    }
    return TLB200_OK;
}

extern "C" int tlb200_multi_mode_dot(const void* x, const int64_t* shape, int ndim, const int* modes,
                                     const void* const* mats, const int64_t* rows_out, const int64_t* m_row_stride,
                                     const int64_t* m_col_stride, int nmats, int dtype, void* out, void* workspace,
                                     size_t workspace_bytes, int path, void* stream) {
    int st = chain_check(shape, ndim, modes, rows_out, nmats);
    if (st) return st;
    if (!x || !out || !dtype_valid(dtype) || (nmats && (!mats || !m_row_stride || !m_col_stride)) ||
        path < TLB200_PATH_AUTO || path > TLB200_PATH_TCGEN05)
        return TLB200_EINVAL;
    for (int k = 0; k < nmats; ++k)
        if (!mats[k]) return TLB200_EINVAL;
    if (workspace_bytes < tlb200_multi_mode_dot_workspace_bytes(shape, ndim, modes, rows_out, nmats, dtype, path))
        return TLB200_EWORKSPACE;
    if (nmats > 0 && (!workspace || reinterpret_cast<uintptr_t>(workspace) % 256)) return TLB200_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return chain_impl<float>((const float*)x, shape, ndim, modes, reinterpret_cast<const float* const*>(mats), rows_out,
                                 m_row_stride, m_col_stride, nmats, (float*)out, workspace, path, s);
    return chain_impl<double>((const double*)x, shape, ndim, modes, reinterpret_cast<const double* const*>(mats), rows_out,
                              m_row_stride, m_col_stride, nmats, (double*)out, workspace, path, s);
}
