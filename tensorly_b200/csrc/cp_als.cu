// CP-ALS normal-equation kernels: Gram, Gram-Hadamard + LU solve, multiplicative update,
// fast reconstruction error, sum of squares.
//
// Reference (caller side of the hot path): tensorly/decomposition/_cp.py:411-428 (V and
// solve), :217-225 + tensorly/cp_tensor.py:614-644 (error), _nn_cp.py:114-136 (MU),
// _cp.py:350 (tl.norm).  In the reference these are ~25 tiny array-library calls per mode
// (and a host sync inside `solve`); here each is one launch, never synchronising, so a
// whole sweep can sit in a CUDA graph.  All of it is latency-bound R x R work.
#include "common.cuh"

#include <cooperative_groups.h>

namespace tlb200 {
namespace {

constexpr int kMaxRank = 128;

#ifdef TLB_CP_TRACE   // probes/cp_update_probe.cu: phase timestamps of CTA 0
__device__ long long g_cp_trace[16];
#define CP_TRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_cp_trace[i] = clock64(); } while (0)
#else
#define CP_TRACE(i) do { } while (0)
#endif

template <typename T>
struct GramList {
    const T* g[TLB200_MAX_NDIM];
    int n;
};

// V[r][s] = w_r * (prod_{i != mode} G_i[r][s] + l2 * delta_rs) * w_s, evaluated in the
// reference's order: ones * G_a * G_b ..., += Id, then (w[:,None] * V) * w[None,:].
template <typename T>
__device__ __forceinline__ T form_v(const GramList<T>& gl, int mode, int64_t R, const T* __restrict__ w, T l2,
                                    int r, int s) {
    T v = T(1);
    for (int i = 0; i < gl.n; ++i)
        if (i != mode) v = v * gl.g[i][(int64_t)r * R + s];
    if (r == s) v += l2;
    if (w) v = (w[r] * v) * w[s];
    return v;
}

// ---- Gram: partial sums over row blocks, then an ordered reduction -----------------
constexpr int kGramRows = 128;  // rows per CTA

template <typename T, int RB>  // RB = ceil(R / 16)
__global__ void __launch_bounds__(256)
gram_partial_kernel(const T* __restrict__ f, int64_t rows, int R, int64_t rs, int64_t cs, T* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Fs = reinterpret_cast<T*>(smem_raw);   // [32][R + 1]
    const int ld = R + 1;
    const int tid = threadIdx.x, tr = tid >> 4, ts = tid & 15;
    const int64_t row0 = (int64_t)blockIdx.x * kGramRows;
    T acc[RB][RB];
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[i][j] = T(0);
    for (int base = 0; base < kGramRows; base += 32) {
        __syncthreads();
        for (int e = tid; e < 32 * R; e += 256) {
            // pick the lane-fastest index along the contiguous dim of F
            int rr, c;
            if (cs <= rs) { rr = e / R; c = e - rr * R; } else { c = e / 32; rr = e - c * 32; }
            const int64_t gr = row0 + base + rr;
            Fs[rr * ld + c] = gr < rows ? f[gr * rs + c * cs] : T(0);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            T a[RB], b[RB];
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                a[i] = (tr + 16 * i) < R ? Fs[k * ld + tr + 16 * i] : T(0);
                b[i] = (ts + 16 * i) < R ? Fs[k * ld + ts + 16 * i] : T(0);
            }
#pragma unroll
            for (int i = 0; i < RB; ++i)
#pragma unroll
                for (int j = 0; j < RB; ++j) acc[i][j] += a[i] * b[j];
        }
    }
    T* out = partial + (int64_t)blockIdx.x * R * R;
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const int r = tr + 16 * i, s = ts + 16 * j;
            if (r < R && s < R) out[r * R + s] = acc[i][j];
        }
}

template <typename T>
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const T* __restrict__ partial, int nblk, int RR, T* __restrict__ gram) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < RR; e += gridDim.x * blockDim.x) {
        T s = T(0);
        for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * RR + e];
        gram[e] = s;
    }
}

// ---- CP update: V, LU with partial pivoting, solve for a block of rows ---------------
constexpr int kSolveRows = 64;   // rows of M per CTA (general path)
constexpr int kFastRows = 32;    // rows of M per CTA on the register-LU path: two passes of 16 rows x 16 lanes
constexpr int kSolveThreads = 256;

// A = V^T in shared memory (ld = R + 1); perm[] row permutation.  All 256 threads of the CTA, arranged
// as 8 warps: lane = column offset, warp = row offset (no integer divisions in the O(R^3) part).
template <typename T>
__device__ void lu_factor_smem(T* A, int* perm, int R, int ld, int* s_piv) {
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int k = 0; k < R; ++k) {
        // pivot search by warp 0 (partial pivoting, first maximum like LAPACK's idamax)
        if (ty == 0) {
            T best = T(-1);
            int bi = k;
            for (int i = k + tx; i < R; i += 32) {
                T v = A[i * ld + k];
                v = v < T(0) ? -v : v;
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                T ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tx == 0) *s_piv = bi;
        }
        __syncthreads();
        const int piv = *s_piv;
        if (piv != k && ty == 0) {
            for (int c = tx; c < R; c += 32) {
                T t = A[k * ld + c]; A[k * ld + c] = A[piv * ld + c]; A[piv * ld + c] = t;
            }
            if (tx == 0) { int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
        }
        __syncthreads();
        const T inv = T(1) / A[k * ld + k];
        // row i is owned by one warp: all lanes read A[i][k], then lane 0 overwrites it with the multiplier
        for (int i = k + 1 + ty; i < R; i += 8) {
            const T m = A[i * ld + k] * inv;
            for (int j = k + 1 + tx; j < R; j += 32) A[i * ld + j] -= m * A[k * ld + j];
            __syncwarp();
            if (tx == 0) A[i * ld + k] = m;
        }
        __syncthreads();
    }
}

// Optional tail of cp_update: Gram of the freshly solved factor rows (still in shared memory) without a second
// pass over the factor.  Every CTA writes its R x R partial; the last CTA to arrive (ticket counter, left at zero
// again) adds the partials in block order, so the result does not depend on scheduling.
template <typename T>
struct GramTail {
    T* partial;          // [gridDim.x][R*R] or null = no Gram requested
    T* gram;             // [R][R]
    unsigned* counter;   // counter[0], counter[1]: zero on entry, zero on exit
    int parallel;        // 1: every CTA is resident (grid <= #SMs), so they may wait for each other: each CTA then
                         //    sums a slice of the R x R entries over all partials instead of the last CTA summing all
};

// Where the right-hand sides come from: the MTTKRP itself (splits == 1), or the split-K partials of the MTTKRP
// kernel [splits][rows][ld], which the solve sums in split order while it loads them (the separate reduction
// launch disappears; on the register-LU path the 192 threads that do not factor do the summing meanwhile).
template <typename T>
struct MSource {
    const T* m;
    int64_t ld;
    int splits;
    int64_t split_stride;
    T* m_out;            // optional: the summed MTTKRP (rows x R, row stride m_out_ld) for callers that need it
    int64_t m_out_ld;
    double* iprod_partial;   // optional [gridDim.x]: per-CTA sum of M o F_new (the <X, model> term of the fast error)
    T* iprod_out;            // device scalar receiving the ordered sum of the above
    // optional: finish the fast reconstruction error (tensorly/decomposition/_cp.py:217-225) in the tail of this
    // launch — err_out = [sqrt(|norm_x2 + ||cp||^2 - 2 <M, F>|) / sqrt(norm_x2), <M, F>, ||cp||^2] — for the LAST mode
    T* err_out;
    const T* norm_x2;
    double* ncp_partial;     // [gridDim.x]
};

template <typename T>
__device__ __forceinline__ T msource_load(const MSource<T>& ms, int64_t row, int c) {
    const T* p = ms.m + row * ms.ld + c;
    if (ms.splits == 1) return *p;
    return ordered_sum_strided<T>(p, ms.splits, (size_t)ms.split_stride);
}

// Y^T Y of the CTA's rows, 256 threads as a 16 x 16 grid, thread (tr, tc) owning the TS x TS outputs
// (tr + 16a, tc + 16b): per row TS + TS shared-memory reads (conflict-free / broadcast) feed TS*TS FMAs.
template <typename T, int TS>
__device__ __forceinline__ void gram_tile(const T* Y, int ld, int R, int nrows, T* __restrict__ dst) {
    const int tc = threadIdx.x & 15, tr = threadIdx.x >> 4;
    T acc[TS][TS];
#pragma unroll
    for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b) acc[a][b] = T(0);
    for (int i = 0; i < nrows; ++i) {
        const T* y = Y + i * ld;
        T yr[TS], yc[TS];
#pragma unroll
        for (int a = 0; a < TS; ++a) {
            yr[a] = tr + 16 * a < R ? y[tr + 16 * a] : T(0);
            yc[a] = tc + 16 * a < R ? y[tc + 16 * a] : T(0);
        }
#pragma unroll
        for (int a = 0; a < TS; ++a)
#pragma unroll
            for (int b = 0; b < TS; ++b) acc[a][b] += yr[a] * yc[b];
    }
#pragma unroll
    for (int a = 0; a < TS; ++a)
#pragma unroll
        for (int b = 0; b < TS; ++b)
            if (tr + 16 * a < R && tc + 16 * b < R) dst[(tr + 16 * a) * R + tc + 16 * b] = acc[a][b];
}

template <typename T>
__device__ __forceinline__ void iprod_finish(const MSource<T>& ms) {      // one warp: ordered sum of the per-CTA terms
    if (threadIdx.x >= 32) return;
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) t += __ldcg(ms.iprod_partial + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) *ms.iprod_out = (T)t;
}

template <typename T>
__device__ __forceinline__ void error_finish(const MSource<T>& ms) {      // one warp of the CTA that finishes last
    if (threadIdx.x >= 32) return;
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) t += __ldcg(ms.ncp_partial + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
        const double ip = (double)__ldcg(ms.iprod_out), nx2 = (double)ms.norm_x2[0];
        double d = nx2 + t - 2.0 * ip;
        d = d < 0 ? -d : d;
        ms.err_out[0] = (T)(sqrt(d) / sqrt(nx2));
        ms.err_out[1] = (T)ip;
        ms.err_out[2] = (T)t;
    }
}

// entry e = (r, s) of w_r w_s prod_{i != mode} G_i[r, s], times the freshly summed Gram entry of `mode`
template <typename T>
__device__ __forceinline__ double ncp_term(const GramList<T>& gl, int mode, const T* __restrict__ w, int R, int e, T g_new) {
    T v = T(1);
    for (int i = 0; i < gl.n; ++i)
        if (i != mode) v = v * gl.g[i][e];
    v = v * g_new;
    if (w) { const int r = e / R, s2 = e - r * R; v = v * (w[r] * w[s2]); }
    return (double)v;
}

template <typename T>
__device__ __forceinline__ void gram_tail(const GramTail<T>& gt, const MSource<T>& ms, const GramList<T>& gl, int mode,
                                          const T* __restrict__ w, const T* Y, int ld, int R, int nrows) {
    if (gt.partial == nullptr) return;
    __shared__ double ncp_red[8];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    T* mine = gt.partial + (size_t)blockIdx.x * R * R;
    if (R <= 16) gram_tile<T, 1>(Y, ld, R, nrows, mine);
    else if (R <= 32) gram_tile<T, 2>(Y, ld, R, nrows, mine);
    else if (R <= 64) gram_tile<T, 4>(Y, ld, R, nrows, mine);
    else gram_tile<T, 8>(Y, ld, R, nrows, mine);
    __threadfence();
    __syncthreads();
    if (gt.parallel) {
        // grid-wide rendezvous (all CTAs resident by construction), then a parallel, still rank-ordered, sum: the
        // serial tail of one CTA walking gridDim.x partials per entry cost more than the solve it finished
        if (tid == 0) {
            atomicAdd(gt.counter, 1u);
            unsigned spins = 0;
            while (atomicAdd(gt.counter, 0u) < gridDim.x) {
                if (++spins > (1u << 26)) asm volatile("trap;");
            }
        }
        __syncthreads();
        __threadfence();
        const int per = (R * R + (int)gridDim.x - 1) / (int)gridDim.x;
        const int e0 = (int)blockIdx.x * per, e1 = min(R * R, e0 + per);
        double ncp = 0.0;
        for (int e = e0 + tid; e < e1; e += blockDim.x) {
            const T g = ordered_sum_strided<T>(gt.partial + e, (int)gridDim.x, (size_t)R * R);
            gt.gram[e] = g;
            if (ms.err_out != nullptr) ncp += ncp_term<T>(gl, mode, w, R, e, g);
        }
        if (ms.iprod_out != nullptr && blockIdx.x == 0) iprod_finish<T>(ms);
        if (ms.err_out != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ncp += __shfl_xor_sync(0xffffffffu, ncp, o);
            if ((tid & 31) == 0) ncp_red[tid >> 5] = ncp;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ncp_red[i];
                ms.ncp_partial[blockIdx.x] = t;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(gt.counter + 1, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (ms.err_out != nullptr) error_finish<T>(ms);
            if (tid == 0) { gt.counter[0] = 0u; gt.counter[1] = 0u; }
        }
        return;
    }
    if (tid == 0) s_last = atomicAdd(gt.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double ncp = 0.0;
    for (int e = tid; e < R * R; e += blockDim.x) {
        const T g = ordered_sum_strided<T>(gt.partial + e, (int)gridDim.x, (size_t)R * R);
        gt.gram[e] = g;
        if (ms.err_out != nullptr) ncp += ncp_term<T>(gl, mode, w, R, e, g);
    }
    if (ms.iprod_out != nullptr) iprod_finish<T>(ms);
    if (ms.err_out != nullptr) {          // this CTA is alone here: its sum is the whole ||cp||^2
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ncp += __shfl_xor_sync(0xffffffffu, ncp, o);
        if ((tid & 31) == 0) ncp_red[tid >> 5] = ncp;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ncp_red[i];
            for (unsigned b = 0; b < gridDim.x; ++b) ms.ncp_partial[b] = b == 0 ? t : 0.0;
        }
        __threadfence();
        __syncthreads();
        error_finish<T>(ms);
    }
    if (tid == 0) *gt.counter = 0u;
}

// ---- fast solve for R <= RM (RM = 32, or 64 in fp32): everything latency-critical lives in registers ----------
//
// LU: thread t of the first RM threads owns row t of A = V^T in registers.  Partial pivoting is implicit (a pivot
// row is marked used instead of being swapped: same pivots and arithmetic as the swapping form).  The row is kept
// ROTATED so that the current column is always register 0: the step loop stays a real loop (~100 instructions
// of code) while every register index is static.  Step k: the unused row with the largest |a[0]| (one redux +
// ballot in fp32) publishes its rotated row as Urot[k] through shared memory (vector stores); everyone else
// eliminates against it and shifts left.  Multipliers go to Lraw[original row][k].
// Substitution: 16 lanes per right-hand side (axpy form, like trsm), see solve_fast.
template <int RM>
__device__ __forceinline__ void lu_barrier() {
    if constexpr (RM > 32) asm volatile("bar.sync 1, %0;" ::"n"(RM) : "memory");
    else __syncwarp();
}

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename T, int RM>
__device__ __forceinline__ void store_row(T* dst, const T (&a)[RM]) {
    using V = typename Vec4<T>::type;
#pragma unroll
    for (int j = 0; j < RM / 4; ++j) {
        V v; v.x = a[4 * j]; v.y = a[4 * j + 1]; v.z = a[4 * j + 2]; v.w = a[4 * j + 3];
        reinterpret_cast<V*>(dst)[j] = v;
    }
}
// b[j-1] = b[j] - c[j] * s for j = 1..RM-1 (c read with vector loads), b[RM-1] = 0; `keep`: rotate only
template <typename T, int RM>
__device__ __forceinline__ void axpy_rotate(T (&b)[RM], const T* c, T s, bool keep) {
    using V = typename Vec4<T>::type;
    T cj[RM];
#pragma unroll
    for (int j = 0; j < RM / 4; ++j) {
        const V v = reinterpret_cast<const V*>(c)[j];
        cj[4 * j] = v.x; cj[4 * j + 1] = v.y; cj[4 * j + 2] = v.z; cj[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int j = 1; j < RM; ++j) b[j - 1] = keep ? b[j] : b[j] - cj[j] * s;
    b[RM - 1] = T(0);
}

template <typename T, int RM>
struct SolveSmem {
    T* A;       // [R][ld]   V^T staging
    T* Y;       // [kSolveRows][ld] right-hand sides, then solutions
    T* Lraw;    // [R][ld]   multipliers by original row
    T* Urot;    // [R][RM]   Urot[k][j] = U[k][k+j]
    T* Lcol;    // [R][RM]   Lcol[k][i] = L[i][k] for i > k, else 0   (column k of L, un-rotated)
    T* Ucol;    // [R][RM]   Ucol[k][i] = U[i][k] for i < k, else 0   (column k of U)
    T* inv;     // [RM]      1 / U[k][k]
    T* s_val;   // [4]
    int* perm;  // [R]
    int* s_idx; // [4]
    unsigned* s_key;  // [4]
    __host__ __device__ static size_t elems(int R) { return (size_t)(2 * R + 64) * (R + 1) + 3 + (size_t)3 * R * RM + RM + 4; }
    __host__ __device__ static size_t bytes(int R) { return sizeof(T) * elems(R) + sizeof(int) * (R + 8); }
    __device__ SolveSmem(unsigned char* raw, int R) {
        const int ld = R + 1;
        A = reinterpret_cast<T*>(raw);
        Y = A + R * ld;
        Lraw = Y + 64 * ld;
        size_t off = (size_t)(2 * R + 64) * ld;
        off = (off + 3) & ~(size_t)3;                 // vector rows: 4-element aligned
        Urot = A + off;
        Lcol = Urot + R * RM;
        Ucol = Lcol + R * RM;
        inv = Ucol + R * RM;
        s_val = inv + RM;
        perm = reinterpret_cast<int*>(s_val + 4);
        s_idx = perm + R;
        s_key = reinterpret_cast<unsigned*>(s_idx + 4);
    }
};

template <typename T, int RM>
__device__ __forceinline__ void lu_factor_rot(int R, const SolveSmem<T, RM>& sm, int ld) {
    const int row = threadIdx.x;            // < RM
    const int lane = row & 31;
    T a[RM];
#pragma unroll
    for (int j = 0; j < RM; ++j) a[j] = (row < R && j < R) ? sm.A[row * ld + j] : T(0);
    bool used = row >= R;
#pragma unroll 2
    for (int k = 0; k < R; ++k) {
        // ---- pivot: unused row with the largest |a[0]|, lowest row on ties; every thread learns its value
        int bi;
        T piv;
        if constexpr (sizeof(T) == 4) {
            // non-negative floats order like their bit patterns; +1 so that even a zero beats a used row
            const unsigned key = used ? 0u : (__float_as_uint(fabsf((float)a[0])) + 1u);
            unsigned mx = __reduce_max_sync(0xffffffffu, key);
            const unsigned bal = __ballot_sync(0xffffffffu, key == mx);
            const unsigned neg = __ballot_sync(0xffffffffu, a[0] < T(0));
            const int pl = __ffs(bal) - 1;
            bi = (row & ~31) | pl;
            float pv = __uint_as_float(mx - 1u);
            if ((neg >> pl) & 1u) pv = -pv;
            if constexpr (RM > 32) {
                const int slot = (k & 1) * 2;
                if (lane == 0) { sm.s_key[slot + (row >> 5)] = mx; sm.s_idx[slot + (row >> 5)] = bi; sm.s_val[slot + (row >> 5)] = (T)pv; }
                lu_barrier<RM>();
                const int pick = sm.s_key[slot + 1] > sm.s_key[slot] ? 1 : 0;      // ties: the lower row (warp 0)
                bi = sm.s_idx[slot + pick];
                pv = (float)sm.s_val[slot + pick];
            }
            piv = (T)pv;
        } else {
            T best = used ? T(-1) : (a[0] < T(0) ? -a[0] : a[0]);
            bi = row;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const T ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            piv = __shfl_sync(0xffffffffu, a[0], bi & 31);      // fp64: RM == 32, one warp
        }
        const T m = a[0] * (T(1) / piv);
        T* pr = sm.Urot + k * RM;
        if (row == bi) {
            used = true;
            sm.perm[k] = row;
            store_row<T, RM>(pr, a);
        }
        lu_barrier<RM>();
        if (!used) sm.Lraw[row * ld + k] = m;
        // eliminate and rotate left in one pass (used rows only rotate; their values are never read again)
        axpy_rotate<T, RM>(a, pr, m, used);
    }
}

// One CTA: factor V^T, then solve kFastRows right-hand sides (rows of M).  `tmp` = this thread's share of the
// right-hand sides, fetched by the caller before the factorisation so the loads are long done.
template <typename T, int RM>
__device__ __forceinline__ void solve_fast(unsigned char* raw, const GramList<T>& gl, int mode, int R, const T* __restrict__ w,
                                           T l2, const MSource<T>& ms, int64_t rows, T* __restrict__ out,
                                           int64_t out_ld, const GramTail<T>& gtail) {
    constexpr int kRows = kFastRows;
    const SolveSmem<T, RM> sm(raw, R);
    const int ld = R + 1;
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * kRows;
    CP_TRACE(0);
    {   // V^T into shared memory: same evaluation order as form_v, Gram pointers in registers, loads batched
        const T* gp[TLB200_MAX_NDIM];
#pragma unroll
        for (int i = 0; i < TLB200_MAX_NDIM; ++i) gp[i] = (i < gl.n && i != mode) ? gl.g[i] : nullptr;
#pragma unroll 4
        for (int e = tid; e < R * R; e += 256) {
            const int r = e / R, s2 = e - r * R;
            T v = T(1);
#pragma unroll
            for (int i = 0; i < TLB200_MAX_NDIM; ++i)
                if (gp[i] != nullptr) v = v * __ldg(gp[i] + e);
            if (r == s2) v += l2;
            if (w) v = (__ldg(w + r) * v) * __ldg(w + s2);
            sm.A[s2 * ld + r] = v;
        }
    }
    __syncthreads();
    CP_TRACE(1);
    if (tid < RM) {
        lu_factor_rot<T, RM>(R, sm, ld);
    } else {
        // the 256 - RM threads that do not factor fetch the right-hand sides meanwhile — summing the MTTKRP's
        // split-K partials in split order when that is what they were given
        for (int e = tid - RM; e < kRows * R; e += 256 - RM) {
            const int rr = e / R, c = e - rr * R;
            const int64_t gr = row0 + rr;
            const T v = gr < rows ? msource_load<T>(ms, gr, c) : T(0);
            sm.Y[rr * ld + c] = v;
            if (ms.m_out != nullptr && gr < rows) ms.m_out[gr * ms.m_out_ld + c] = v;
        }
    }
    CP_TRACE(2);
    __syncthreads();
    // substitution operands: whole columns of L (below the diagonal) and of U (above it), zero elsewhere, so the
    // updates below need no bounds: Lcol[k][i] = L[i][k], Ucol[k][i] = U[i][k]; inv[k] = 1 / U[k][k]
    for (int e = tid; e < R * RM; e += 256) {
        const int k = e / RM, i = e - k * RM;
        sm.Lcol[e] = (i > k && i < R) ? sm.Lraw[sm.perm[i] * ld + k] : T(0);
        sm.Ucol[e] = i < k ? sm.Urot[i * RM + (k - i)] : T(0);
    }
    for (int k = tid; k < RM; k += 256) sm.inv[k] = k < R ? T(1) / sm.Urot[k * RM] : T(0);
    __syncthreads();
    CP_TRACE(3);
    // 16 lanes per right-hand side: lane l of a group owns the entries i = l, l + 16, ... of its vector.  Step k:
    // the owner of entry k broadcasts it inside the group (one shuffle), everyone applies column k to its entries
    // (axpy form, like trsm): RM / 16 FMAs per lane and step, the shuffle is the only thing on the dependency chain.
    constexpr int NS = RM / 16;
    const int grp = tid >> 4, l16 = tid & 15;
    double ip_acc = 0.0;
#pragma unroll 1
    for (int pass = 0; pass < kRows / 16; ++pass) {
        T* y = sm.Y + (pass * 16 + grp) * ld;
        T b[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int i = t * 16 + l16;
            b[t] = i < R ? y[sm.perm[i]] : T(0);                 // P * rhs
        }
        __syncwarp();
        // forward substitution, unit lower triangular
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
#pragma unroll 4
            for (int kk = 0; kk < 16; ++kk) {
                const int k = sl * 16 + kk;
                if (k >= R) break;
                const T yk = __shfl_sync(0xffffffffu, b[sl], kk, 16);
                const T* c = sm.Lcol + k * RM + l16;
#pragma unroll
                for (int t = 0; t < NS; ++t) b[t] = fma(-c[t * 16], yk, b[t]);
            }
        }
        CP_TRACE(4);
        // back substitution: x_k = b_k / U_kk, then b_i -= U[i][k] x_k for i < k
#pragma unroll
        for (int sl = NS - 1; sl >= 0; --sl) {
#pragma unroll 4
            for (int kk = 15; kk >= 0; --kk) {
                const int k = sl * 16 + kk;
                if (k >= R) continue;
                const T xk = __shfl_sync(0xffffffffu, b[sl] * sm.inv[k], kk, 16);
                if (l16 == kk) b[sl] = xk;
                const T* c = sm.Ucol + k * RM + l16;
#pragma unroll
                for (int t = 0; t < NS; ++t) b[t] = fma(-c[t * 16], xk, b[t]);
            }
        }
        __syncwarp();
        if (ms.iprod_partial != nullptr) {          // <M_row, F_row>: y still holds the right-hand side here
#pragma unroll
            for (int t = 0; t < NS; ++t) {
                const int i = t * 16 + l16;
                if (i < R) ip_acc += (double)y[i] * (double)b[t];
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NS; ++t) {
            const int i = t * 16 + l16;
            if (i < R) y[i] = b[t];
        }
    }
    CP_TRACE(5);
    __syncthreads();
    if (ms.iprod_partial != nullptr) {
        __shared__ double ip_red[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ip_acc += __shfl_xor_sync(0xffffffffu, ip_acc, o);
        if ((tid & 31) == 0) ip_red[tid >> 5] = ip_acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < 8; ++i) t += ip_red[i];
            ms.iprod_partial[blockIdx.x] = t;
        }
    }
    for (int e = tid; e < kRows * R; e += 256) {
        const int r2 = e / R, c = e - r2 * R;
        const int64_t gr = row0 + r2;
        if (gr < rows) out[gr * out_ld + c] = sm.Y[r2 * ld + c];
    }
    CP_TRACE(6);
    gram_tail<T>(gtail, ms, gl, mode, w, sm.Y, ld, R, (int)min((int64_t)kRows, rows - row0));
}

template <typename T>
__global__ void __launch_bounds__(kSolveThreads, 1)
cp_update_kernel(GramList<T> gl, int mode, int R, const T* __restrict__ w, T l2, MSource<T> ms,
                 int64_t rows, T* __restrict__ out, int64_t out_ld, GramTail<T> gtail) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = R + 1;
    T* A = reinterpret_cast<T*>(smem_raw);            // [R][ld]   = V^T, then its LU factors
    T* Y = A + R * ld;                                // [kSolveRows][ld] right-hand sides / solutions
    int* perm = reinterpret_cast<int*>(Y + kSolveRows * ld);
    int* s_piv = perm + R;
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * kSolveRows;

    if (R <= 32) { solve_fast<T, 32>(smem_raw, gl, mode, R, w, l2, ms, rows, out, out_ld, gtail); return; }
    if constexpr (sizeof(T) == 4) {
        if (R <= 64) { solve_fast<T, 64>(smem_raw, gl, mode, R, w, l2, ms, rows, out, out_ld, gtail); return; }
    }
    {
        for (int e = tid; e < R * R; e += blockDim.x) {
            const int r = e / R, s = e - r * R;
            A[s * ld + r] = form_v<T>(gl, mode, R, w, l2, r, s);   // store V^T
        }
        for (int i = tid; i < R; i += blockDim.x) perm[i] = i;
        __syncthreads();
        lu_factor_smem<T>(A, perm, R, ld, s_piv);
        // load permuted right-hand sides: Y[row][i] = M[row][perm[i]]
        for (int e = tid; e < kSolveRows * R; e += blockDim.x) {
            const int rr = e / R, c = e - rr * R;
            const int64_t gr = row0 + rr;
            const T v = gr < rows ? msource_load<T>(ms, gr, perm[c]) : T(0);
            Y[rr * ld + c] = v;
            if (ms.m_out != nullptr && gr < rows) ms.m_out[gr * ms.m_out_ld + perm[c]] = v;
        }
        __syncthreads();
    }
    CP_TRACE(3);
    // 4 threads per row: thread q of a row owns the partial dot products of columns
    // j = q, q+4, ... ; the running solution lives in shared memory.
    const int rr = tid >> 2, q = tid & 3;
    T* y = Y + rr * ld;
    // forward substitution, unit lower triangular
    for (int i = 1; i < R; ++i) {
        T s = T(0);
        for (int j = q; j < i; j += 4) s += A[i * ld + j] * y[j];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0) y[i] -= s;
        __syncwarp();
    }
    CP_TRACE(4);
    // back substitution
    for (int i = R - 1; i >= 0; --i) {
        T s = T(0);
        for (int j = i + 1 + q; j < R; j += 4) s += A[i * ld + j] * y[j];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0) y[i] = (y[i] - s) / A[i * ld + i];
        __syncwarp();
    }
    CP_TRACE(5);
    __syncthreads();
    for (int e = tid; e < kSolveRows * R; e += blockDim.x) {
        const int r2 = e / R, c = e - r2 * R;
        const int64_t gr = row0 + r2;
        if (gr < rows) out[gr * out_ld + c] = Y[r2 * ld + c];
    }
    CP_TRACE(6);
    gram_tail<T>(gtail, ms, gl, mode, w, Y, ld, R, (int)min((int64_t)kSolveRows, rows - row0));
}

// ---- NN-CP multiplicative update -------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
nncp_update_kernel(GramList<T> gl, int mode, int R, const T* __restrict__ w, const T* __restrict__ m, int64_t m_ld,
                   T* __restrict__ f, int64_t f_ld, int64_t rows, T eps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = R + 1;
    T* V = reinterpret_cast<T*>(smem_raw);  // [R][ld]
    T* Fs = V + R * ld;                     // [32][ld]
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    for (int e = tid; e < R * R; e += 256) {
        const int r = e / R, s = e - r * R;
        V[r * ld + s] = form_v<T>(gl, mode, R, w, T(0), r, s);
    }
    for (int e = tid; e < 32 * R; e += 256) {
        const int rr = e / R, c = e - rr * R;
        const int64_t gr = row0 + rr;
        Fs[rr * ld + c] = gr < rows ? f[gr * f_ld + c] : T(0);
    }
    __syncthreads();
    const int rr = tid >> 3, cg = tid & 7;
    const int64_t gr = row0 + rr;
    if (gr >= rows) return;
    for (int c = cg; c < R; c += 8) {
        T den = T(0);
        for (int s = 0; s < R; ++s) den += Fs[rr * ld + s] * V[s * ld + c];
        T num = m[gr * m_ld + c];
        num = num < eps ? eps : num;
        den = den < eps ? eps : den;
        f[gr * f_ld + c] = Fs[rr * ld + c] * num / den;
    }
}

// ---- error ----------------------------------------------------------------------------
// One thread-block cluster of 8 CTAs: each sums <M, F> over an eighth of the rows, the partial sums meet in CTA 0
// through distributed shared memory and are added in CTA order (deterministic, no global scratch).  A single CTA
// walking all of M and F was 33 us of pure latency at the end of every C5 sweep.
constexpr int kErrCluster = 8;

template <typename T>
__global__ void __cluster_dims__(kErrCluster, 1, 1) __launch_bounds__(1024)
cp_error_kernel(GramList<T> gl, int R, const T* __restrict__ w, const T* __restrict__ m, int64_t m_ld,
                const T* __restrict__ f, int64_t frs, int64_t fcs, int64_t rows, const T* __restrict__ norm_x2,
                T* __restrict__ err_out) {
    __shared__ double red[2][32];
    __shared__ double cta_sum[2];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    double iprod = 0.0, ncp = 0.0;
    const int64_t rows_per = (rows + kErrCluster - 1) / kErrCluster;
    const int64_t r0 = (int64_t)crank * rows_per, r1 = min(rows, r0 + rows_per);
    const int64_t total = max((int64_t)0, r1 - r0) * R;
    if (total < (1LL << 31)) {
        // 32-bit index arithmetic and four independent loads in flight
        const int tot = (int)total, step = (int)blockDim.x;
#pragma unroll 4
        for (int e = tid; e < tot; e += step) {
            const int i = e / R, r = e - i * R;
            iprod += (double)__ldg(m + (r0 + i) * m_ld + r) * (double)__ldg(f + (r0 + i) * frs + (int64_t)r * fcs);
        }
    } else {
        for (int64_t e = tid; e < total; e += blockDim.x) {
            const int64_t i = e / R, r = e - i * R;
            iprod += (double)m[(r0 + i) * m_ld + r] * (double)f[(r0 + i) * frs + r * fcs];
        }
    }
    if (crank == 0) {
        for (int e = tid; e < R * R; e += blockDim.x) {
            const int r = e / R, s = e - r * R;
            T v = T(1);
            for (int i = 0; i < gl.n; ++i) v = v * gl.g[i][(int64_t)r * R + s];
            if (w) v = v * (w[r] * w[s]);
            ncp += (double)v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        iprod += __shfl_xor_sync(0xffffffffu, iprod, o);
        ncp += __shfl_xor_sync(0xffffffffu, ncp, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = iprod; red[1][tid >> 5] = ncp; }
    __syncthreads();
    if (tid < 32) {
        iprod = tid < (int)(blockDim.x >> 5) ? red[0][tid] : 0.0;
        ncp = tid < (int)(blockDim.x >> 5) ? red[1][tid] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            iprod += __shfl_xor_sync(0xffffffffu, iprod, o);
            ncp += __shfl_xor_sync(0xffffffffu, ncp, o);
        }
        if (tid == 0) { cta_sum[0] = iprod; cta_sum[1] = ncp; }
    }
    cluster.sync();
    if (crank == 0 && tid == 0) {
        double ip = 0.0;
        for (int c = 0; c < kErrCluster; ++c) ip += cluster.map_shared_rank(cta_sum, c)[0];
        const double nc = cta_sum[1];
        const double nx2 = (double)norm_x2[0];
        double d = nx2 + nc - 2.0 * ip;
        d = d < 0 ? -d : d;
        err_out[0] = (T)(sqrt(d) / sqrt(nx2));
        err_out[1] = (T)ip;
        err_out[2] = (T)nc;
    }
    cluster.sync();          // nobody leaves while CTA 0 may still read its shared memory
}

// ---- sum of squares ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const T* __restrict__ x, int64_t n, double* __restrict__ partial) {
    __shared__ double red[8];
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = (double)__ldg(x + i);
        s += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}

// float4 variant for aligned fp32 arrays (the 4 GB tensor of config 2)
__global__ void __launch_bounds__(256)
sumsq_partial_f4_kernel(const float4* __restrict__ x, int64_t n4, double* __restrict__ partial) {
    __shared__ double red[8];
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(x + i);
        s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}

template <typename T>
__global__ void sumsq_final_kernel(const double* __restrict__ partial, int n, const T* __restrict__ tail, int ntail,
                                   T* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += partial[i];
        for (int i = 0; i < ntail; ++i) t += (double)tail[i] * (double)tail[i];
        out[0] = (T)t;
    }
}

constexpr int kSumsqBlocks = kNumSMs * 8;

template <typename T>
int gram_launch(const T* f, int64_t rows, int64_t R, int64_t rs, int64_t cs, T* gram, void* workspace,
                cudaStream_t stream) {
    const int nblk = (int)ceil_div(rows, kGramRows);
    T* partial = reinterpret_cast<T*>(workspace);
    const size_t smem = sizeof(T) * 32 * (R + 1);
    const int RB = (int)ceil_div(R, 16);
    if (RB <= 1) gram_partial_kernel<T, 1><<<nblk, 256, smem, stream>>>(f, rows, (int)R, rs, cs, partial);
    else if (RB <= 2) gram_partial_kernel<T, 2><<<nblk, 256, smem, stream>>>(f, rows, (int)R, rs, cs, partial);
    else if (RB <= 4) gram_partial_kernel<T, 4><<<nblk, 256, smem, stream>>>(f, rows, (int)R, rs, cs, partial);
    else gram_partial_kernel<T, 8><<<nblk, 256, smem, stream>>>(f, rows, (int)R, rs, cs, partial);
    TLB_CHECK_LAUNCH();
    const int RR = (int)(R * R);
    gram_reduce_kernel<T><<<(RR + 255) / 256, 256, 0, stream>>>(partial, nblk, RR, gram);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <typename T>
int fill_grams(GramList<T>* gl, const void* const* grams, int nmodes, int skip) {
    if (!grams || nmodes < 1 || nmodes > TLB200_MAX_NDIM) return TLB200_EINVAL;
    gl->n = nmodes;
    for (int i = 0; i < TLB200_MAX_NDIM; ++i) gl->g[i] = nullptr;
    for (int i = 0; i < nmodes; ++i) {
        if (i != skip && !grams[i]) return TLB200_EINVAL;
        gl->g[i] = static_cast<const T*>(grams[i]);
    }
    return TLB200_OK;
}

template <typename T>
int cp_update_launch(const void* const* grams, int nmodes, int mode, int64_t R, const T* w, double l2, MSource<T> ms,
                     int64_t rows, T* out, int64_t out_ld, T* gram_out, void* workspace, cudaStream_t stream) {
    GramList<T> gl;
    int st = fill_grams<T>(&gl, grams, nmodes, mode);
    if (st) return st;
    const int ld = (int)R + 1;
    size_t smem = sizeof(T) * ((size_t)R * ld + (size_t)kSolveRows * ld) + sizeof(int) * (R + 1);
    if (R <= 32) smem = SolveSmem<T, 32>::bytes((int)R);
    else if (sizeof(T) == 4 && R <= 64) smem = SolveSmem<T, 64>::bytes((int)R);
    static std::atomic<uint64_t> attr_done{0};        // one per instantiation (T), one bit per device
    if (ensure_dynamic_smem(cp_update_kernel<T>, 200 * 1024, attr_done)) return TLB200_ECUDA;
    if (smem > 200 * 1024) return TLB200_EUNSUPPORTED;
    const bool fast = R <= 32 || (sizeof(T) == 4 && R <= 64);          // must match the dispatch in cp_update_kernel
    const int nblk = (int)ceil_div(rows, fast ? kFastRows : kSolveRows);
    if (nblk == 0) return TLB200_OK;
    GramTail<T> gt;
    gt.partial = nullptr; gt.gram = gram_out; gt.counter = nullptr;
    gt.parallel = nblk > 1 && nblk <= kNumSMs;          // one CTA per SM at most: all of them are resident
    if (gram_out != nullptr) {      // workspace: [ticket counters, 256 bytes][nblk][R*R][nblk doubles: <M, F> terms]
        gt.counter = static_cast<unsigned*>(workspace);
        gt.partial = reinterpret_cast<T*>(static_cast<char*>(workspace) + 256);
        double* scal = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256 + align_up((size_t)nblk * R * R * sizeof(T), 256));
        ms.iprod_partial = ms.iprod_out ? scal : nullptr;
        ms.ncp_partial = ms.err_out ? scal + nblk : nullptr;
    } else {
        ms.iprod_partial = nullptr; ms.iprod_out = nullptr; ms.err_out = nullptr; ms.ncp_partial = nullptr;
    }
    if (ms.err_out != nullptr && (ms.iprod_out == nullptr || ms.norm_x2 == nullptr)) return TLB200_EINVAL;
    if (ms.iprod_out != nullptr && !fast) return TLB200_EUNSUPPORTED;      // only the register-LU path forms <M, F>
    cp_update_kernel<T><<<nblk, kSolveThreads, smem, stream>>>(gl, mode, (int)R, w, (T)l2, ms, rows, out, out_ld, gt);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <typename T>
int nncp_update_launch(const void* const* grams, int nmodes, int mode, int64_t R, const T* w, const T* m, int64_t m_ld,
                       T* f, int64_t f_ld, int64_t rows, double eps, cudaStream_t stream) {
    GramList<T> gl;
    int st = fill_grams<T>(&gl, grams, nmodes, mode);
    if (st) return st;
    const int ld = (int)R + 1;
    const size_t smem = sizeof(T) * ((size_t)R * ld + 32 * (size_t)ld);
    static std::atomic<uint64_t> attr_done{0};
    if (ensure_dynamic_smem(nncp_update_kernel<T>, 200 * 1024, attr_done)) return TLB200_ECUDA;
    const int nblk = (int)ceil_div(rows, 32);
    if (nblk == 0) return TLB200_OK;
    nncp_update_kernel<T><<<nblk, 256, smem, stream>>>(gl, mode, (int)R, w, m, m_ld, f, f_ld, rows, (T)eps);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

}  // namespace
}  // namespace tlb200

using namespace tlb200;

extern "C" size_t tlb200_gram_workspace_bytes(int64_t rows, int64_t rank, int dtype) {
    if (rows < 0 || rank < 1 || !dtype_valid(dtype)) return 0;
    return align_up((size_t)ceil_div(rows > 0 ? rows : 1, kGramRows) * rank * rank * dtype_size(dtype), 256);
}

extern "C" int tlb200_gram(const void* f, int64_t rows, int64_t rank, int64_t row_stride, int64_t col_stride, int dtype,
                           void* gram, void* workspace, size_t workspace_bytes, void* stream) {
    if (!f || !gram || !workspace || rows < 1 || rank < 1 || rank > kMaxRank || !dtype_valid(dtype)) return TLB200_EINVAL;
    if (workspace_bytes < tlb200_gram_workspace_bytes(rows, rank, dtype)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32) return gram_launch<float>((const float*)f, rows, rank, row_stride, col_stride, (float*)gram, workspace, s);
    return gram_launch<double>((const double*)f, rows, rank, row_stride, col_stride, (double*)gram, workspace, s);
}

template <typename T>
static MSource<T> plain_source(const void* m, int64_t m_ld) {
    MSource<T> ms;
    ms.m = static_cast<const T*>(m); ms.ld = m_ld; ms.splits = 1; ms.split_stride = 0;
    ms.m_out = nullptr; ms.m_out_ld = 0; ms.iprod_partial = nullptr; ms.iprod_out = nullptr;
    ms.err_out = nullptr; ms.norm_x2 = nullptr; ms.ncp_partial = nullptr;
    return ms;
}

extern "C" int tlb200_cp_update(const void* const* grams, int nmodes, int mode, int64_t rank, const void* weights,
                                double l2_reg, const void* m, int64_t m_ld, int64_t rows, int dtype, void* out,
                                int64_t out_ld, void* stream) {
    if (!m || !out || rank < 1 || rank > kMaxRank || rows < 0 || mode < 0 || mode >= nmodes || m_ld < rank ||
        out_ld < rank || !dtype_valid(dtype))
        return TLB200_EINVAL;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return cp_update_launch<float>(grams, nmodes, mode, rank, (const float*)weights, l2_reg, plain_source<float>(m, m_ld), rows,
                                       (float*)out, out_ld, nullptr, nullptr, s);
    return cp_update_launch<double>(grams, nmodes, mode, rank, (const double*)weights, l2_reg, plain_source<double>(m, m_ld), rows,
                                    (double*)out, out_ld, nullptr, nullptr, s);
}

extern "C" size_t tlb200_cp_update_gram_workspace_bytes(int64_t rows, int64_t rank, int dtype) {
    if (rows < 0 || rank < 1 || !dtype_valid(dtype)) return 0;
    const size_t nblk = (size_t)ceil_div(rows > 0 ? rows : 1, kFastRows);
    return 256 + align_up(nblk * rank * rank * dtype_size(dtype), 256) + align_up(2 * nblk * sizeof(double), 256);
}

extern "C" int tlb200_cp_update_gram(const void* const* grams, int nmodes, int mode, int64_t rank, const void* weights,
                                     double l2_reg, const void* m, int64_t m_ld, int64_t rows, int dtype, void* out,
                                     int64_t out_ld, void* gram_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!m || !out || !gram_out || !workspace || rank < 1 || rank > kMaxRank || rows < 1 || mode < 0 || mode >= nmodes ||
        m_ld < rank || out_ld < rank || !dtype_valid(dtype))
        return TLB200_EINVAL;
    if (workspace_bytes < tlb200_cp_update_gram_workspace_bytes(rows, rank, dtype)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return cp_update_launch<float>(grams, nmodes, mode, rank, (const float*)weights, l2_reg, plain_source<float>(m, m_ld), rows,
                                       (float*)out, out_ld, (float*)gram_out, workspace, s);
    return cp_update_launch<double>(grams, nmodes, mode, rank, (const double*)weights, l2_reg, plain_source<double>(m, m_ld), rows,
                                    (double*)out, out_ld, (double*)gram_out, workspace, s);
}

// The fused form the own driver uses: right-hand sides = split-K partials of the MTTKRP (summed in split order while
// they are loaded), plus optionally the summed MTTKRP itself (m_out) and <M, F_new> (iprod_out, one device scalar).
extern "C" int tlb200_cp_update_fused(const void* const* grams, int nmodes, int mode, int64_t rank, const void* weights,
                                      double l2_reg, const tlb200_partials_t* m, int dtype, void* out, int64_t out_ld,
                                      void* gram_out, void* m_out, int64_t m_out_ld, void* iprod_out, const void* norm_x2,
                                      void* err_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!m || !m->data || !out || !gram_out || !workspace || rank < 1 || rank > kMaxRank || m->rows < 1 || mode < 0 ||
        mode >= nmodes || m->ld < rank || out_ld < rank || m->splits < 1 || !dtype_valid(dtype) ||
        (m_out && m_out_ld < rank))
        return TLB200_EINVAL;
    if (workspace_bytes < tlb200_cp_update_gram_workspace_bytes(m->rows, rank, dtype)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) {
        using T = decltype(tag);
        MSource<T> ms;
        ms.m = static_cast<const T*>(m->data); ms.ld = m->ld; ms.splits = (int)m->splits; ms.split_stride = m->split_stride;
        ms.m_out = static_cast<T*>(m_out); ms.m_out_ld = m_out_ld;
        ms.iprod_partial = nullptr; ms.iprod_out = static_cast<T*>(iprod_out);
        ms.err_out = static_cast<T*>(err_out); ms.norm_x2 = static_cast<const T*>(norm_x2); ms.ncp_partial = nullptr;
        return cp_update_launch<T>(grams, nmodes, mode, rank, static_cast<const T*>(weights), l2_reg, ms, m->rows,
                                   static_cast<T*>(out), out_ld, static_cast<T*>(gram_out), workspace, s);
    };
    return dtype == TLB200_F32 ? go(float()) : go(double());
}

// err_out = [sqrt(|norm_x2 + norm_cp^2 - 2 iprod|) / sqrt(norm_x2), iprod, norm_cp^2] from a device-resident <M, F>
// (tlb200_cp_update_fused's iprod_out): the R x R part of tlb200_cp_error alone.
template <typename T>
__global__ void __launch_bounds__(256)
cp_error_iprod_kernel(GramList<T> gl, int R, const T* __restrict__ w, const T* __restrict__ iprod,
                      const T* __restrict__ norm_x2, T* __restrict__ err_out) {
    __shared__ double red[8];
    double ncp = 0.0;
    for (int e = threadIdx.x; e < R * R; e += 256) {
        const int r = e / R, s2 = e - r * R;
        T v = T(1);
        for (int i = 0; i < gl.n; ++i) v = v * gl.g[i][e];
        if (w) v = v * (w[r] * w[s2]);
        ncp += (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ncp += __shfl_xor_sync(0xffffffffu, ncp, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ncp;
    __syncthreads();
    if (threadIdx.x == 0) {
        double nc = 0.0;
        for (int i = 0; i < 8; ++i) nc += red[i];
        const double ip = (double)iprod[0], nx2 = (double)norm_x2[0];
        double d = nx2 + nc - 2.0 * ip;
        d = d < 0 ? -d : d;
        err_out[0] = (T)(sqrt(d) / sqrt(nx2));
        err_out[1] = (T)ip;
        err_out[2] = (T)nc;
    }
}

extern "C" int tlb200_cp_error_iprod(const void* const* grams, int nmodes, int64_t rank, const void* weights,
                                     const void* iprod, const void* norm_x2, int dtype, void* err_out, void* stream) {
    if (!iprod || !norm_x2 || !err_out || rank < 1 || rank > kMaxRank || !dtype_valid(dtype)) return TLB200_EINVAL;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32) {
        GramList<float> gl;
        int st = fill_grams<float>(&gl, grams, nmodes, -1);
        if (st) return st;
        cp_error_iprod_kernel<float><<<1, 256, 0, s>>>(gl, (int)rank, (const float*)weights, (const float*)iprod,
                                                       (const float*)norm_x2, (float*)err_out);
    } else {
        GramList<double> gl;
        int st = fill_grams<double>(&gl, grams, nmodes, -1);
        if (st) return st;
        cp_error_iprod_kernel<double><<<1, 256, 0, s>>>(gl, (int)rank, (const double*)weights, (const double*)iprod,
                                                        (const double*)norm_x2, (double*)err_out);
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

extern "C" int tlb200_nncp_update(const void* const* grams, int nmodes, int mode, int64_t rank, const void* weights,
                                  const void* m, int64_t m_ld, void* f, int64_t f_ld, int64_t rows, double eps, int dtype,
                                  void* stream) {
    if (!m || !f || rank < 1 || rank > kMaxRank || rows < 0 || mode < 0 || mode >= nmodes || m_ld < rank || f_ld < rank ||
        !dtype_valid(dtype))
        return TLB200_EINVAL;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return nncp_update_launch<float>(grams, nmodes, mode, rank, (const float*)weights, (const float*)m, m_ld, (float*)f,
                                         f_ld, rows, eps, s);
    return nncp_update_launch<double>(grams, nmodes, mode, rank, (const double*)weights, (const double*)m, m_ld, (double*)f,
                                      f_ld, rows, eps, s);
}

extern "C" int tlb200_cp_error(const void* const* grams, int nmodes, int64_t rank, const void* weights, const void* m_last,
                               int64_t m_ld, const void* f_last, int64_t f_row_stride, int64_t f_col_stride, int64_t rows,
                               const void* norm_x2, int dtype, void* err_out, void* stream) {
    if (!m_last || !f_last || !norm_x2 || !err_out || rank < 1 || rank > kMaxRank || rows < 1 || !dtype_valid(dtype))
        return TLB200_EINVAL;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32) {
        GramList<float> gl;
        int st = fill_grams<float>(&gl, grams, nmodes, -1);
        if (st) return st;
        cp_error_kernel<float><<<kErrCluster, 1024, 0, s>>>(gl, (int)rank, (const float*)weights, (const float*)m_last, m_ld,
                                                  (const float*)f_last, f_row_stride, f_col_stride, rows,
                                                  (const float*)norm_x2, (float*)err_out);
    } else {
        GramList<double> gl;
        int st = fill_grams<double>(&gl, grams, nmodes, -1);
        if (st) return st;
        cp_error_kernel<double><<<kErrCluster, 1024, 0, s>>>(gl, (int)rank, (const double*)weights, (const double*)m_last, m_ld,
                                                   (const double*)f_last, f_row_stride, f_col_stride, rows,
                                                   (const double*)norm_x2, (double*)err_out);
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

extern "C" size_t tlb200_sumsq_workspace_bytes(int64_t, int) { return align_up(sizeof(double) * kSumsqBlocks, 256); }

extern "C" int tlb200_sumsq(const void* x, int64_t n, int dtype, void* out, void* workspace, size_t workspace_bytes,
                            void* stream) {
    if (!x || !out || !workspace || n < 0 || !dtype_valid(dtype)) return TLB200_EINVAL;
    if (workspace_bytes < tlb200_sumsq_workspace_bytes(n, dtype)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(workspace);
    if (dtype == TLB200_F32) {
        const float* xf = static_cast<const float*>(x);
        if (reinterpret_cast<uintptr_t>(x) % 16 == 0) {
            const int64_t n4 = n / 4;
            sumsq_partial_f4_kernel<<<kSumsqBlocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x), n4, partial);
            TLB_CHECK_LAUNCH();
            sumsq_final_kernel<float><<<1, 32, 0, s>>>(partial, kSumsqBlocks, xf + n4 * 4, (int)(n - n4 * 4), (float*)out);
        } else {
            sumsq_partial_kernel<float><<<kSumsqBlocks, 256, 0, s>>>(xf, n, partial);
            TLB_CHECK_LAUNCH();
            sumsq_final_kernel<float><<<1, 32, 0, s>>>(partial, kSumsqBlocks, xf, 0, (float*)out);
        }
    } else {
        const double* xd = static_cast<const double*>(x);
        sumsq_partial_kernel<double><<<kSumsqBlocks, 256, 0, s>>>(xd, n, partial);
        TLB_CHECK_LAUNCH();
        sumsq_final_kernel<double><<<1, 32, 0, s>>>(partial, kSumsqBlocks, xd, 0, (double*)out);
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}
