// HALS non-negative least squares as ONE kernel.
//
// Reference: tensorly/solvers/nnls.py:139-173 — for every inner iteration a Python loop over the rank with ~6
// array-library calls per row of V (R x n), i.e. up to 100 x R x 6 launches per mode update of
// non_negative_parafac_hals (tensorly/decomposition/_nn_cp.py:326-336).  The columns of V are independent, so
// here one thread owns one column (one row of the factor being updated) for the WHOLE solve:
//
//   per column, per inner iteration, for k = 0..R-1 (Gauss-Seidel order, as in the reference):
//     num = UtM[k] - UtU[k,:] . v + UtU[k,k] v_k  (- sparsity);   den = UtU[k,k] (+ 2 ridge)
//     new = max(num / den, epsilon);   v_k <- new
//
// The thread keeps the gradient g = UtM - UtU v in registers and applies the rank-1 correction
// g -= (new - v_k) UtU[:, k] after each coordinate: R FMAs per coordinate instead of an R-term dot product plus
// R more for the update.  g is stored ROTATED (the current coordinate is always register 0), so the coordinate
// loop is a real loop with static register indices — the same trick as the LU in cp_als.cu.  v lives in shared
// memory (dynamic index), UtU's rotated columns are broadcast from shared memory.
//
// The reference's stopping rule is global: rec_error = sum_k ||V - newV_k||^2 where newV_k (a ROW) is subtracted
// from the whole matrix by broadcasting (nnls.py:150: tl.norm(V - newV) ** 2) — restated as is, from running sums
// S1 = sum_l v_l, S2 = sum_l v_l^2 per column; the per-CTA sums are combined in a fixed order after a grid-wide
// barrier (cooperative launch), so every CTA takes the same decision and the result is deterministic.
#include "common.cuh"

#include <cooperative_groups.h>
#include <type_traits>

namespace tlb200 {
namespace {

namespace cg = cooperative_groups;

constexpr int HT = 128;            // threads (columns) per CTA

template <typename T>
struct HalsGrams {
    const T* g[TLB200_MAX_NDIM];
    int n;
};

template <typename T>
struct HalsParams {
    HalsGrams<T> gl;           // UtU = (w w^T) o prod_{i != mode} G_i  (mode < 0: UtU = g[0] as is)
    int mode;
    const T* w;
    const T* m;                // UtM^T: element (column j, coordinate k) at m[j * m_rs + k * m_cs]
    int64_t m_rs, m_cs;
    T* f;                      // V^T, updated in place: f[j * f_rs + k * f_cs]
    int64_t f_rs, f_cs;
    int64_t n;                 // columns of V = rows of the factor
    int R;
    int n_iter_max;
    double tol;
    int has_sparsity, has_ridge;
    T sparsity, ridge, epsilon;
    double* partial;           // [2][gridDim.x]
    int* iters_out;            // inner iterations actually run (device int, optional)
};

template <typename T, int RM>
__global__ void __launch_bounds__(HT)
hals_kernel(const HalsParams<T> p) {
    extern __shared__ __align__(16) unsigned char hals_smem[];
    // UtU is kept in DOUBLE in shared memory: the gradient update reads R entries of it per coordinate and thread, and
    // converting them from fp32 every time (F2F.F64.F32 issues at a quarter of the DFMA rate) cost more than the FMAs
    double* Urot = reinterpret_cast<double*>(hals_smem);   // [RM][RM]: Urot[k][j] = UtU[(k + j) % RM][k]
    T* vs = reinterpret_cast<T*>(Urot + RM * RM);          // [RM][HT]: v_k of this thread's column at vs[k * HT + tid]
    __shared__ T diag[RM];
    __shared__ double red[HT / 32];
    __shared__ double s_total;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, R = p.R;
    const int64_t col = (int64_t)blockIdx.x * HT + tid;
    const bool live = col < p.n;

    // UtU in the reference's evaluation order (ones * G_a * G_b ..., then (w[:,None] * V) * w[None,:]), rotated
    for (int e = tid; e < RM * RM; e += HT) {
        const int k = e / RM, j = e - k * RM;
        const int l = (k + j) % RM;
        T v = T(0);
        if (l < R && k < R) {
            if (p.mode < 0) {
                v = p.gl.g[0][(int64_t)l * R + k];
            } else {
                v = T(1);
                for (int i = 0; i < p.gl.n; ++i)
                    if (i != p.mode) v = v * p.gl.g[i][(int64_t)l * R + k];
                if (p.w) v = (p.w[l] * v) * p.w[k];
            }
        }
        Urot[e] = (double)v;
        if (j == 0) diag[k] = v;
    }
    // this column of V, its running sums, and the gradient g = UtM - UtU v (rotation 0: g[j] <-> coordinate j)
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < RM; ++k) {
        const T v = (live && k < R) ? p.f[col * p.f_rs + k * p.f_cs] : T(0);
        vs[k * HT + tid] = v;
        s1 += (double)v;
        s2 += (double)v * (double)v;
    }
    __syncthreads();
    // The gradient is carried in DOUBLE whatever T is: it is updated incrementally R x n_iter times, and in fp32 the
    // rounding drift of those updates (a difference of large terms) reached 2e-4 of the solution after 100 passes,
    // whereas the reference recomputes every dot product from scratch.
    double g[RM];
#pragma unroll
    for (int j = 0; j < RM; ++j) g[j] = (live && j < R) ? (double)p.m[col * p.m_rs + j * p.m_cs] : 0.0;
    for (int k = 0; k < R; ++k) {          // g[j] -= UtU[j][k] * v_k, with UtU[j][k] = Urot[k][(j - k) mod RM]
        const double vk = (double)vs[k * HT + tid];
        const double* c = Urot + k * RM;
#pragma unroll
        for (int j = 0; j < RM; ++j) g[j] -= c[(j - k) & (RM - 1)] * vk;
    }

    const double Rd = (double)R;
    int it = 0;
    double err0 = 0.0;
    for (; it < p.n_iter_max; ++it) {
        double rec = 0.0;
        for (int k = 0; k < RM; ++k) {
            const T dkk = diag[k];
            double delta = 0.0;
            if (k < R && dkk != T(0)) {
                const T vk = vs[k * HT + tid];
                T num = (T)(g[0] + (double)dkk * (double)vk);
                T den = dkk;
                if (p.has_sparsity) num -= p.sparsity;
                if (p.has_ridge) den += T(2) * p.ridge;
                T nv = num / den;
                nv = nv > p.epsilon ? nv : p.epsilon;
                if (live) {
                    const double nd = (double)nv;
                    rec += s2 - 2.0 * nd * s1 + Rd * nd * nd;          // sum_l (v_l - new)^2, old v
                    s1 += nd - (double)vk;
                    s2 += nd * nd - (double)vk * (double)vk;
                }
                delta = (double)nv - (double)vk;
                vs[k * HT + tid] = nv;
            }
            // g <- rotate_left(g - delta * UtU[:, k]): coordinate k + 1 moves to register 0
            const double* c = Urot + k * RM;
            const double g0 = g[0] - delta * c[0];
#pragma unroll
            for (int j = 1; j < RM; ++j) g[j - 1] = g[j] - delta * c[j];
            g[RM - 1] = g0;
        }
        // global stopping statistic, fixed summation order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rec += __shfl_xor_sync(0xffffffffu, rec, o);
        if ((tid & 31) == 0) red[tid >> 5] = rec;
        __syncthreads();
        double* slot = p.partial + (size_t)(it & 1) * gridDim.x;
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < HT / 32; ++i) t += red[i];
            slot[blockIdx.x] = t;
        }
        grid.sync();
        if (tid < 32) {
            double t = 0.0;
            for (unsigned b = tid; b < gridDim.x; b += 32) t += __ldcg(slot + b);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (tid == 0) s_total = t;
        }
        __syncthreads();
        const double total = s_total;
        if (it == 0) err0 = total;
        if (total < p.tol * err0) { ++it; break; }
    }
    if (live)
        for (int k = 0; k < R; ++k) p.f[col * p.f_rs + k * p.f_cs] = vs[k * HT + tid];
    if (p.iters_out && blockIdx.x == 0 && tid == 0) *p.iters_out = it;
}

template <typename T, int RM>
int launch_rm(const HalsParams<T>& p, cudaStream_t stream) {
    const int smem = (int)sizeof(double) * RM * RM + (int)sizeof(T) * RM * HT;
    static std::atomic<uint64_t> attr_done{0};
    if (ensure_dynamic_smem(hals_kernel<T, RM>, smem, attr_done)) return TLB200_ECUDA;
    const int64_t nblk = ceil_div(p.n, HT);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hals_kernel<T, RM>, HT, smem) != cudaSuccess) return TLB200_ECUDA;
    if (nblk > (int64_t)per_sm * kNumSMs) return TLB200_EUNSUPPORTED;      // the grid barrier needs every CTA resident
    void* args[] = {const_cast<HalsParams<T>*>(&p)};
    if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(hals_kernel<T, RM>), dim3((unsigned)nblk), dim3(HT), args, smem,
                                    stream) != cudaSuccess)
        return TLB200_ECUDA;
    count_launch();
    return TLB200_OK;
}

template <typename T>
int launch(const HalsParams<T>& p, cudaStream_t stream) {
    if (p.R <= 8) return launch_rm<T, 8>(p, stream);
    if (p.R <= 16) return launch_rm<T, 16>(p, stream);
    if (p.R <= 32) return launch_rm<T, 32>(p, stream);
    if (sizeof(T) == 4 && p.R <= 64) return launch_rm<T, 64>(p, stream);
    return TLB200_EUNSUPPORTED;
}

}  // namespace
}  // namespace tlb200

using namespace tlb200;

extern "C" size_t tlb200_hals_workspace_bytes(int64_t rows) {
    if (rows < 1) return 0;
    return align_up((size_t)2 * ceil_div(rows, HT) * sizeof(double), 256) + 256;
}

// grams / nmodes / mode / weights as in tlb200_cp_update; mode < 0: grams[0] IS UtU (plain hals_nnls).
extern "C" int tlb200_hals_update(const void* const* grams, int nmodes, int mode, int64_t rank, const void* weights,
                                  const void* m, int64_t m_row_stride, int64_t m_col_stride, void* f, int64_t f_row_stride,
                                  int64_t f_col_stride, int64_t rows, int n_iter_max, double tol, const double* sparsity,
                                  const double* ridge, double epsilon, int dtype, void* iters_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    if (!grams || !m || !f || !workspace || rank < 1 || rows < 1 || nmodes < 1 || nmodes > TLB200_MAX_NDIM || mode >= nmodes ||
        n_iter_max < 0 || !dtype_valid(dtype))
        return TLB200_EINVAL;
    if (workspace_bytes < tlb200_hals_workspace_bytes(rows)) return TLB200_EWORKSPACE;
    for (int i = 0; i < nmodes; ++i)
        if (i != mode && !grams[i]) return TLB200_EINVAL;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto fill = [&](auto& p) {
        using T = typename std::remove_pointer<decltype(p.f)>::type;
        p.gl.n = mode < 0 ? 1 : nmodes;
        for (int i = 0; i < TLB200_MAX_NDIM; ++i) p.gl.g[i] = i < nmodes ? static_cast<const T*>(grams[i]) : nullptr;
        p.mode = mode;
        p.w = static_cast<const T*>(weights);
        p.m = static_cast<const T*>(m); p.m_rs = m_row_stride; p.m_cs = m_col_stride;
        p.f = static_cast<T*>(f); p.f_rs = f_row_stride; p.f_cs = f_col_stride;
        p.n = rows; p.R = (int)rank; p.n_iter_max = n_iter_max; p.tol = tol;
        p.has_sparsity = sparsity != nullptr; p.has_ridge = ridge != nullptr;
        p.sparsity = sparsity ? (T)*sparsity : T(0); p.ridge = ridge ? (T)*ridge : T(0); p.epsilon = (T)epsilon;
        p.partial = static_cast<double*>(workspace);
        p.iters_out = static_cast<int*>(iters_out);
    };
    if (dtype == TLB200_F32) { HalsParams<float> p; fill(p); return launch<float>(p, s); }
    HalsParams<double> p; fill(p); return launch<double>(p, s);
}
