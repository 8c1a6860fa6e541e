// Dimension-tree reuse for CP-ALS: one pass over the tensor,
//     T = X x_{N-1} F_{N-1}^T        (tlb200_mode_dot, shape I_0 x .. x I_{N-2} x R),
// serves the MTTKRPs of ALL modes n < N-1 of a sweep, because
//     M_n[j, r] = w_r * sum_{a, b} T[a, j, b, r] * P[a, r] * Q[b, r],
// P = Khatri-Rao of the factors before mode n, Q = of the factors between n and N-1.  T is R / I_{N-1} the size
// of X (3 % at C2), so a 3-way sweep streams X twice instead of three times with exactly the same ALS algebra
// (the reference recomputes the full MTTKRP per mode: tensorly/decomposition/_cp.py:407-428).
//
// The kernels here are plain bandwidth-bound SIMT reductions over T (no GEMM shape: the contraction is
// element-wise in r).  Deterministic: fixed lane order inside a CTA, split partials summed in split order.
#include "common.cuh"
#include "stream_gemm.cuh"

namespace tlb200 {
namespace {

template <typename T, int VW>
struct alignas(sizeof(T) * VW) Vec {
    T v[VW];
};

constexpr int kThreads = 256;

// T viewed as [A][J][B][R] (contiguous).  Grid (J, a-splits).  Threads: rv = vector column, bl = lane over b.
//   partial[split][j][:] = sum_{a in split} P[a, :] * sum_b T[a, j, b, :] * Q[b, :]
template <typename T, int VW>
__global__ void __launch_bounds__(kThreads)
from_ttm_inner_kernel(const T* __restrict__ t, int64_t A, int64_t J, int64_t B, int R, const T* __restrict__ P,
                      const T* __restrict__ Q, int64_t a_per_split, int b_splits, int64_t b_per_split, T* __restrict__ out,
                      int64_t out_ld, int64_t out_split) {
    using V = Vec<T, VW>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V* red = reinterpret_cast<V*>(smem_raw);
    const int RV = R / VW;
    const int nbl = kThreads / RV;
    const int rv = threadIdx.x % RV, bl = threadIdx.x / RV;
    const int64_t j = blockIdx.x;
    const int64_t a0 = (int64_t)(blockIdx.y / b_splits) * a_per_split;      // split = (a range, b range)
    const int64_t a1 = min(A, a0 + a_per_split);
    const int64_t b0 = (int64_t)(blockIdx.y % b_splits) * b_per_split;
    const int64_t b1 = min(B, b0 + b_per_split);
    V acc;
#pragma unroll
    for (int c = 0; c < VW; ++c) acc.v[c] = T(0);
    if (bl < nbl) {
        for (int64_t a = a0; a < a1; ++a) {
            const V* row = reinterpret_cast<const V*>(t + ((a * J + j) * B) * R) + rv;
            V tmp;
#pragma unroll
            for (int c = 0; c < VW; ++c) tmp.v[c] = T(0);
#pragma unroll 4
            for (int64_t b = b0 + bl; b < b1; b += nbl) {
                const V x = row[b * RV];
                const V q = reinterpret_cast<const V*>(Q + b * R)[rv];
#pragma unroll
                for (int c = 0; c < VW; ++c) tmp.v[c] += x.v[c] * q.v[c];
            }
            if (P != nullptr) {
                const V p = reinterpret_cast<const V*>(P + a * R)[rv];
#pragma unroll
                for (int c = 0; c < VW; ++c) acc.v[c] += p.v[c] * tmp.v[c];
            } else {
#pragma unroll
                for (int c = 0; c < VW; ++c) acc.v[c] += tmp.v[c];
            }
        }
        red[bl * RV + rv] = acc;
    }
    __syncthreads();
    if (threadIdx.x < RV) {
        V s = red[threadIdx.x];
        for (int l = 1; l < nbl; ++l) {
            const V o = red[l * RV + threadIdx.x];
#pragma unroll
            for (int c = 0; c < VW; ++c) s.v[c] += o.v[c];
        }
        T* dst = out + (int64_t)blockIdx.y * out_split + j * out_ld + threadIdx.x * VW;
#pragma unroll
        for (int c = 0; c < VW; ++c) dst[c] = s.v[c];
    }
}

// B == 1: T viewed as [A][J][R].  Grid (j blocks, a-splits).  Threads: rv = vector column, jl = lane over j.
//   partial[split][j][:] = Q[0, :] * sum_{a in split} P[a, :] * T[a, j, :]
// B == 1 also happens when every lead mode after `mode` is a singleton: Q is then a single row (it may carry the
// weights), and P is absent when `mode` is the first mode — both are optional here.
template <typename T, int VW>
__global__ void __launch_bounds__(kThreads)
from_ttm_outer_kernel(const T* __restrict__ t, int64_t A, int64_t J, int R, const T* __restrict__ P,
                      const T* __restrict__ Qrow, int64_t a_per_split, T* __restrict__ out, int64_t out_ld,
                      int64_t out_split) {
    using V = Vec<T, VW>;
    const int RV = R / VW;
    const int njl = kThreads / RV;
    const int rv = threadIdx.x % RV, jl = threadIdx.x / RV;
    const int64_t j = (int64_t)blockIdx.x * njl + jl;
    if (jl >= njl || j >= J) return;
    const int64_t a0 = (int64_t)blockIdx.y * a_per_split;
    const int64_t a1 = min(A, a0 + a_per_split);
    V acc;
#pragma unroll
    for (int c = 0; c < VW; ++c) acc.v[c] = T(0);
    if (P != nullptr) {
#pragma unroll 8
        for (int64_t a = a0; a < a1; ++a) {
            const V x = reinterpret_cast<const V*>(t + (a * J + j) * R)[rv];
            const V p = reinterpret_cast<const V*>(P + a * R)[rv];
#pragma unroll
            for (int c = 0; c < VW; ++c) acc.v[c] += x.v[c] * p.v[c];
        }
    } else {
        for (int64_t a = a0; a < a1; ++a) {
            const V x = reinterpret_cast<const V*>(t + (a * J + j) * R)[rv];
#pragma unroll
            for (int c = 0; c < VW; ++c) acc.v[c] += x.v[c];
        }
    }
    if (Qrow != nullptr) {
        const V q = reinterpret_cast<const V*>(Qrow)[rv];
#pragma unroll
        for (int c = 0; c < VW; ++c) acc.v[c] *= q.v[c];
    }
    T* dst = out + (int64_t)blockIdx.y * out_split + j * out_ld + rv * VW;
#pragma unroll
    for (int c = 0; c < VW; ++c) dst[c] = acc.v[c];
}

struct Geometry {
    int64_t A, J, B;
    int pf, pc, qf, qc;
    int64_t splits, a_per_split;       // splits = a_splits * b_splits partial results
    int64_t b_splits, b_per_split;
};

int make_geometry(const int64_t* lead_shape, int nlead, int mode, int64_t rank, int dtype, Geometry* g) {
    if (!lead_shape || nlead < 2 || nlead >= TLB200_MAX_NDIM || mode < 0 || mode >= nlead || rank < 1 || rank > 256)
        return TLB200_EINVAL;
    g->A = 1; g->B = 1;
    for (int i = 0; i < nlead; ++i) {
        if (lead_shape[i] < 1) return TLB200_EINVAL;
        if (i < mode) g->A *= lead_shape[i];
        if (i > mode) g->B *= lead_shape[i];
    }
    g->J = lead_shape[mode];
    g->pf = 0; g->pc = mode; g->qf = mode + 1; g->qc = nlead - 1 - mode;
    // enough CTAs to fill the machine a few times over; every split re-reads nothing, it only adds a partial
    const int vmax = dtype == TLB200_F32 ? 4 : 2;
    const int64_t rv = rank % vmax == 0 ? rank / vmax : rank;                 // vector columns per row of T
    const int64_t j_per_cta = kThreads / rv > 0 ? kThreads / rv : 1;
    const int64_t ctas_per_split = g->B > 1 ? g->J : ceil_div(g->J, j_per_cta);
    int64_t s = ceil_div((int64_t)kNumSMs * 8, ctas_per_split);
    if (s < 1) s = 1;
    int64_t sa = s > g->A ? g->A : s;
    g->a_per_split = ceil_div(g->A, sa);
    sa = ceil_div(g->A, g->a_per_split);
    // still too few CTAs (a short outer range): split the inner range too, in pieces of at least 8 rows per lane
    g->b_splits = 1;
    g->b_per_split = g->B;
    if (g->B > 1 && sa < s) {
        const int64_t lanes = kThreads / rv > 0 ? kThreads / rv : 1;
        int64_t sb = ceil_div(s, sa);
        const int64_t sb_max = g->B / (lanes * 8) > 0 ? g->B / (lanes * 8) : 1;
        if (sb > sb_max) sb = sb_max;
        g->b_per_split = ceil_div(g->B, sb);
        g->b_splits = ceil_div(g->B, g->b_per_split);
    }
    g->splits = sa * g->b_splits;
    return TLB200_OK;
}

size_t workspace_for(const Geometry& g, int64_t rank, int dtype) {
    const size_t es = dtype_size(dtype);
    size_t total = 256;
    if (g.pc > 0) total += align_up((size_t)g.A * rank * es, 256);
    if (g.qc > 0) total += align_up((size_t)g.B * rank * es, 256);
    total += align_up((size_t)g.splits * g.J * rank * es, 256);       // also for one split: the *_partials entry point
    return total;
}

template <typename T, int VW>
int launch(const T* t, const Geometry& g, int R, const T* P, const T* Q, T* dst, int64_t dst_ld, int64_t dst_split,
           cudaStream_t stream) {
    const int RV = R / VW;
    if (g.B > 1) {
        const size_t smem = sizeof(T) * VW * (size_t)(kThreads / RV) * RV;
        dim3 grid((unsigned)g.J, (unsigned)g.splits);
        from_ttm_inner_kernel<T, VW><<<grid, kThreads, smem, stream>>>(t, g.A, g.J, g.B, R, P, Q, g.a_per_split, (int)g.b_splits,
                                                                        g.b_per_split, dst, dst_ld, dst_split);
    } else {
        const int njl = kThreads / RV;
        dim3 grid((unsigned)ceil_div(g.J, njl), (unsigned)g.splits);
        from_ttm_outer_kernel<T, VW><<<grid, kThreads, 0, stream>>>(t, g.A, g.J, R, P, Q, g.a_per_split, dst, dst_ld, dst_split);
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <typename T>
int run(const T* t, const int64_t* lead_shape, int nlead, int mode, const T* const* factors, const int64_t* frs,
        const int64_t* fcs, int64_t rank, const T* weights, T* out, int64_t out_ld, void* workspace, const Geometry& g,
        cudaStream_t stream, tlb200_partials_t* info = nullptr) {
    Carver ws(workspace);
    const T* P = g.pc > 0 ? ws.take<T>((size_t)g.A * rank) : nullptr;
    const T* Q = g.qc > 0 ? ws.take<T>((size_t)g.B * rank) : nullptr;
    T* partial = (g.splits > 1 || info != nullptr) ? ws.take<T>((size_t)g.splits * g.J * rank) : nullptr;
    const T* w = weights;      // folded into the first table that exists
    int st;
    // a table made of one unweighted, row-major factor is used in place (3-way sweeps: no prep launch at all)
    if (P) {
        if (const T* direct = table_is_factor<T>(factors, frs, fcs, g.pf, g.pc, w, rank, rank)) {
            P = direct;
        } else {
            st = launch_khatri_rao<T>(factors + g.pf, lead_shape + g.pf, frs + g.pf, fcs + g.pf, g.pc, rank, w, nullptr,
                                      const_cast<T*>(P), rank, rank, stream);
            if (st) return st;
            w = nullptr;
        }
    }
    if (Q) {
        if (const T* direct = table_is_factor<T>(factors, frs, fcs, g.qf, g.qc, w, rank, rank)) {
            Q = direct;
        } else {
            st = launch_khatri_rao<T>(factors + g.qf, lead_shape + g.qf, frs + g.qf, fcs + g.qf, g.qc, rank, w, nullptr,
                                      const_cast<T*>(Q), rank, rank, stream);
            if (st) return st;
        }
    }
    T* dst = partial ? partial : out;
    const int64_t dst_ld = partial ? rank : out_ld;
    const int64_t dst_split = partial ? g.J * rank : 0;
    const int R = (int)rank;
    constexpr int VMAX = sizeof(T) == 4 ? 4 : 2;
    const bool vec_ok = R % VMAX == 0 && R / VMAX <= kThreads && reinterpret_cast<uintptr_t>(t) % 16 == 0 &&
                        (partial != nullptr || (out_ld % VMAX == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0));
    if (R > kThreads) return TLB200_EUNSUPPORTED;
    st = vec_ok ? launch<T, VMAX>(t, g, R, P, Q, dst, dst_ld, dst_split, stream)
                : launch<T, 1>(t, g, R, P, Q, dst, dst_ld, dst_split, stream);
    if (st) return st;
    if (info != nullptr) {          // leave the partials unsummed for tlb200_cp_update_fused
        info->data = partial;
        info->splits = g.splits;
        info->split_stride = g.J * rank;
        info->ld = rank;
        info->rows = g.J;
        info->rank = rank;
        return TLB200_OK;
    }
    if (partial) {
        int64_t blocks = ceil_div(g.J * rank, 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        splitk_reduce_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(partial, g.splits, g.J, rank, rank, out, out_ld);
        TLB_CHECK_LAUNCH();
    }
    return TLB200_OK;
}

}  // namespace
}  // namespace tlb200

using namespace tlb200;

extern "C" size_t tlb200_mttkrp_from_ttm_workspace_bytes(const int64_t* lead_shape, int nlead, int mode, int64_t rank,
                                                         int dtype) {
    Geometry g;
    if (!dtype_valid(dtype) || make_geometry(lead_shape, nlead, mode, rank, dtype, &g)) return 0;
    return workspace_for(g, rank, dtype);
}

extern "C" int tlb200_mttkrp_from_ttm(const void* t, const int64_t* lead_shape, int nlead, int mode,
                                      const void* const* factors, const int64_t* f_row_stride, const int64_t* f_col_stride,
                                      int64_t rank, const void* weights, int dtype, void* out, int64_t out_ld,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    Geometry g;
    if (!dtype_valid(dtype)) return TLB200_EINVAL;
    int st = make_geometry(lead_shape, nlead, mode, rank, dtype, &g);
    if (st) return st;
    if (!t || !factors || !f_row_stride || !f_col_stride || !out || out_ld < rank || !workspace) return TLB200_EINVAL;
    for (int i = 0; i < nlead; ++i)
        if (i != mode && !factors[i]) return TLB200_EINVAL;
    if (workspace_bytes < workspace_for(g, rank, dtype)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return run<float>((const float*)t, lead_shape, nlead, mode, reinterpret_cast<const float* const*>(factors), f_row_stride,
                          f_col_stride, rank, (const float*)weights, (float*)out, out_ld, workspace, g, s);
    return run<double>((const double*)t, lead_shape, nlead, mode, reinterpret_cast<const double* const*>(factors), f_row_stride,
                       f_col_stride, rank, (const double*)weights, (double*)out, out_ld, workspace, g, s);
}

extern "C" int tlb200_mttkrp_from_ttm_partials(const void* t, const int64_t* lead_shape, int nlead, int mode,
                                               const void* const* factors, const int64_t* f_row_stride,
                                               const int64_t* f_col_stride, int64_t rank, const void* weights, int dtype,
                                               void* workspace, size_t workspace_bytes, tlb200_partials_t* partials,
                                               void* stream) {
    Geometry g;
    if (!dtype_valid(dtype) || !partials) return TLB200_EINVAL;
    int st = make_geometry(lead_shape, nlead, mode, rank, dtype, &g);
    if (st) return st;
    if (!t || !factors || !f_row_stride || !f_col_stride || !workspace) return TLB200_EINVAL;
    for (int i = 0; i < nlead; ++i)
        if (i != mode && !factors[i]) return TLB200_EINVAL;
    if (workspace_bytes < workspace_for(g, rank, dtype)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return run<float>((const float*)t, lead_shape, nlead, mode, reinterpret_cast<const float* const*>(factors), f_row_stride,
                          f_col_stride, rank, (const float*)weights, nullptr, rank, workspace, g, s, partials);
    return run<double>((const double*)t, lead_shape, nlead, mode, reinterpret_cast<const double* const*>(factors), f_row_stride,
                       f_col_stride, rank, (const double*)weights, nullptr, rank, workspace, g, s, partials);
}
