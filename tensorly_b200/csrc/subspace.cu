// Orthonormalisation of a tall block of vectors — the "SVD" of the own HOOI driver.
//
// The reference's HOOI takes the leading left singular vectors of every mode unfolding of the projected tensor
// with a full LAPACK/cuSOLVER SVD (tensorly/decomposition/_tucker.py:197-201 -> tenalg/svd.py:211-235).  HOOI
// only needs an orthonormal basis of the dominant subspace, so the own driver runs warm-started subspace
// iteration  U <- orth(G U)  on the small Gram matrix G = Y_(k) Y_(k)^T (both products on the existing TTM
// kernels) and this file supplies `orth`: Cholesky-QR with the R x R Gram, its Cholesky factor and the
// triangular inverse in fp64 — exact enough for blocks whose condition number squared exceeds 1/eps(fp32),
// which G U reaches after a single power step on tensors with a dominant mean component.
//
//   kernel 1  partial Grams of 64-row blocks (fp64 accumulation); the last CTA to arrive sums them in block
//             order, factors S = R^T R in shared memory and writes R^{-1} (upper triangular, fp64).
//   kernel 2  Q = Z R^{-1}: one thread per row, the row in registers, R^{-1} broadcast from shared memory.
// Latency-bound R x R work (R <= 64); deterministic.
#include "common.cuh"

namespace tlb200 {
namespace {

__device__ long long g_ps_trace[8];        // perf triage: phase timestamps of the factoring CTA of the last call
#define PS_TRACE(i) do { if (threadIdx.x == 0) g_ps_trace[i] = clock64(); } while (0)

constexpr int OR_MAX = 64;        // widest block
constexpr int OR_ROWS = 64;       // rows per CTA in kernel 1
constexpr int OR_THREADS = 256;

// Shifted Cholesky S + shift I = R^T R in shared memory, then R^{-1} (upper triangular); blockDim = 256, S is
// [64][65], R <= 64.  shift = 2^-43 x mean diagonal: invisible for a well-conditioned block (orthogonality defect
// 1e-13), and what keeps the factorisation alive when the block is numerically rank deficient (G U on a tensor
// with a dominant mean direction reaches cond^2 > 1e16 after two steps; the reference's SVD does not care) — the
// weak directions then come out scaled down instead of as inf/nan, and the next pass repairs them.
__device__ void chol_inverse_64(double (*S)[OR_MAX + 1], double (*Ri)[OR_MAX + 1], int R, int* bad_out) {
    const int tid = threadIdx.x;
    __shared__ double s_shift;
    if (tid == 0) {
        double tr = 0.0;
        for (int i = 0; i < R; ++i) tr += S[i][i];
        s_shift = ldexp(tr / R, -43);
    }
    __syncthreads();
    const double shift = s_shift;
    if (tid < R) S[tid][tid] += shift;
    __syncthreads();
    // thread t updates the 16 elements (i, j) = (ti + 16 a, tj + 16 b) — fixed, no divisions inside the k loop
    const int ti = tid >> 4, tj = tid & 15;
    int bad = 0;
    for (int k = 0; k < R; ++k) {
        if (tid == 0) {
            double d = S[k][k];
            if (!(d > shift)) { d = shift > 0.0 ? shift : 1e-300; bad = 1; }
            S[k][k] = sqrt(d);
        }
        __syncthreads();
        const double inv = 1.0 / S[k][k];
        if (tid > k && tid < R) S[k][tid] *= inv;
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = ti + 16 * a;
            if (i <= k || i >= R) continue;
            const double rki = S[k][i];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = tj + 16 * b;
                if (j >= i && j < R) S[i][j] -= rki * S[k][j];
            }
        }
        __syncthreads();
    }
    PS_TRACE(3);
    // R^{-1} column by column: 4 lanes per column (256 threads = 64 columns x 4), partial sums combined by shuffles
    {
        const int c = tid >> 2, q = tid & 3;
        for (int i = R - 1; i >= 0; --i) {
            double part = 0.0;
            if (c < R && i < c)
                for (int j = i + 1 + q; j <= c; j += 4) part += S[i][j] * Ri[j][c];
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (q == 0 && c < R) Ri[i][c] = i > c ? 0.0 : ((i == c ? 1.0 : 0.0) - part) / S[i][i];
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid == 0 && bad_out) *bad_out = bad;
}

template <typename T>
__global__ void __launch_bounds__(OR_THREADS)
orth_gram_chol_kernel(const T* __restrict__ z, int64_t rows, int R, int64_t rs, int64_t cs, double* __restrict__ partial,
                      unsigned* __restrict__ counter, double* __restrict__ rinv, int* __restrict__ status) {
    // dynamic shared memory: [Zs | S]; R^{-1} later reuses the Zs block (the row tile is dead by then)
    extern __shared__ __align__(16) unsigned char orth_smem[];
    typedef double Row[OR_MAX + 1];
    Row* Zs = reinterpret_cast<Row*>(orth_smem);
    Row* S = Zs + OR_ROWS;
    Row* Ri = Zs;
    static_assert(OR_ROWS >= OR_MAX, "R^{-1} aliases the row tile");
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * OR_ROWS;
    for (int e = tid; e < OR_ROWS * R; e += OR_THREADS) {
        int rr, c;
        if (cs <= rs) { rr = e / R; c = e - rr * R; } else { c = e / OR_ROWS; rr = e - c * OR_ROWS; }
        const int64_t gr = row0 + rr;
        Zs[rr][c] = gr < rows ? (double)z[gr * rs + c * cs] : 0.0;
    }
    __syncthreads();
    double* mine = partial + (size_t)blockIdx.x * R * R;
    for (int e = tid; e < R * R; e += OR_THREADS) {
        const int a = e / R, b = e - a * R;
        double acc = 0.0;
#pragma unroll 8
        for (int i = 0; i < OR_ROWS; ++i) acc = fma(Zs[i][a], Zs[i][b], acc);
        mine[e] = acc;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int e = tid; e < R * R; e += OR_THREADS)
        S[e / R][e % R] = ordered_sum_strided<double>(partial + e, (int)gridDim.x, (size_t)R * R);
    __syncthreads();
    chol_inverse_64(S, Ri, R, status);
    for (int e = tid; e < R * R; e += OR_THREADS) rinv[e] = Ri[e / R][e % R];
    if (tid == 0) *counter = 0u;
}

template <typename T, int RM>
__global__ void __launch_bounds__(128)
orth_apply_kernel(const T* __restrict__ z, int64_t rows, int R, int64_t rs, int64_t cs, const double* __restrict__ rinv,
                  T* __restrict__ out, int64_t out_ld) {
    __shared__ double Ri[RM * RM];
    for (int e = threadIdx.x; e < RM * RM; e += 128) {
        const int i = e / RM, j = e - i * RM;
        Ri[e] = (i < R && j < R) ? rinv[i * R + j] : 0.0;
    }
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (row >= rows) return;
    double zr[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) zr[i] = i < R ? (double)z[row * rs + i * cs] : 0.0;
#pragma unroll 4
    for (int j = 0; j < RM; ++j) {
        if (j >= R) break;
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < RM; ++i) acc = fma(zr[i], Ri[i * RM + j], acc);     // R^{-1} is upper triangular: zeros below
        out[row * out_ld + j] = (T)acc;
    }
}

// ---- fused power step of the HOOI subspace iteration (fp64): U <- orth(G U) ------------------------------------
// kernel A (n / 16 CTAs): a 16-row block of Z = G U, its p x p Gram partial; the last CTA to arrive sums the
//           partials in block order, factors S = R^T R and writes R^{-1}.
// kernel B (n / 16 CTAs): U = Z R^{-1}.
// G (n x n) and U (n x p, p <= 64) are read from L2; nothing here is bandwidth-relevant — the step is a latency chain
// (GEMM block ~4 us, Cholesky + triangular inverse of a 64 x 64 matrix ~20 us, apply ~3 us), which is why it is two
// launches instead of the five (TTM prep + TTM + 2 orth kernels + copies) the generic entry points would take.
constexpr int PS_ROWS = 16;
constexpr int PS_KC = 64;
// tile row strides of 68 doubles (544 bytes, 16-byte aligned): 136 words = 8 mod 32, so the 16 fragment loads of a
// half-warp (4 k x 4 rows / columns) hit 32 distinct banks
constexpr int PS_GLD = 68;
constexpr int PS_ULD = 68;
constexpr int PS_STAGES = 3;
constexpr int PS_STAGE_DOUBLES = PS_KC * PS_ULD + PS_ROWS * PS_GLD;  // 5440 doubles = 43 520 bytes

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;                               // 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(bytes) : "memory");
}


// Shifted Cholesky S + shift I = R^T R for the power step, S held in REGISTERS: thread (ti, tj) owns the 4 x 4 entries
// S[ti + 16a][tj + 16b].  Pivot k: the 16 threads that own row k publish it through a double-buffered shared row
// (one barrier per pivot), everyone forms 1 / d from a float rsqrt seed + two Newton steps and updates its 16
// entries with FMAs — no shared-memory traffic in the trailing update (the all-in-shared-memory version moved ~50
// words per thread and pivot: 625 clk per pivot).  No explicit inverse: kernel B solves the triangular system.
// Rm[k][j] = R[k][j] for j >= k (untouched below the diagonal), invd[k] = 1 / R[k][k].  blockDim = 256.
__device__ __forceinline__ double fast_rsqrt(double d) {
    if (!(d > 1e-30 && d < 1e30)) return rsqrt(d);
    double r = (double)rsqrtf((float)d);
    r = r * (1.5 - 0.5 * d * r * r);
    r = r * (1.5 - 0.5 * d * r * r);
    return r;
}
// One Newton step on the 22-bit float seed: 2^-43 relative.  Enough for the power steps (their orthonormalisation
// only stabilises the iteration — the shift already leaves a 1e-13 defect, and a final two-pass
// tlb200_orthonormalize makes the basis exact); it takes ~100 clk off every pivot's dependency chain.
__device__ __forceinline__ double fast_rsqrt1(double d) {
    if (!(d > 1e-30 && d < 1e30)) return rsqrt(d);
    const double r = (double)rsqrtf((float)d);
    const double e = fma(-d * r, r, 1.0);
    return fma(0.5 * r, e, r);
}
// 1 / sqrt(d) and 1 / d from the hardware's double-precision seeds (MUFU.RSQ64H / RCP64H, ~2^-20) + one Newton step
// each (~2^-40), as two INDEPENDENT chains: the trailing update needs only 1 / d, which is two dependent fp64 operations
// after its seed, while the scaled row (which needs 1 / sqrt) leaves the critical path.  fp64 operations have a long
// latency here, and the pivot loop of the Cholesky is one dependency chain.
__device__ __forceinline__ void pivot_scales(double d, double& rs, double& id) {
    if (!(d > 1e-280 && d < 1e280)) { rs = rsqrt(d); id = rs * rs; return; }
    double r0, i0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(i0) : "d"(d));
    id = fma(i0, fma(-d, i0, 1.0), i0);
    rs = fma(0.5 * r0, fma(-d * r0, r0, 1.0), r0);
}

// pivots 16 KB .. 16 KB + 15.  KB is static, so everything a pivot of this block cannot touch is pruned at compile
// time: register rows x < KB are finished, pairs with y < x lie below the diagonal (10 / 6 / 3 / 1 live pairs for
// KB = 0..3 instead of 16 predicated ones — the loop was ~200 instructions per pivot and issue-bound).
template <int KB>
__device__ __forceinline__ void chol_block(double (&a)[4][4], double (*rowk)[OR_MAX], double* __restrict__ Rm,
                                           double* __restrict__ invd, int R, double shift, int& bad, int ti, int tj, int tid) {
#pragma unroll 1
    for (int kl = 0; kl < 16; ++kl) {
        const int k = 16 * KB + kl;
        if (k >= R) break;
        const double* rk = rowk[k & 1];
        double d = rk[k];
        if (!(d > shift)) { d = shift > 0.0 ? shift : 1e-300; bad = 1; }
        double rs, id;
        pivot_scales(d, rs, id);
        // row k of R goes out (threads 0..63, one column each)
        if (tid >= k && tid < R) Rm[k * R + tid] = tid == k ? d * rs : rk[tid] * rs;
        if (tid == k) invd[k] = rs;
        double ri[4], rj[4];
#pragma unroll
        for (int x = KB; x < 4; ++x) { ri[x] = rk[ti + 16 * x]; rj[x] = rk[tj + 16 * x]; }
#pragma unroll
        for (int x = KB; x < 4; ++x)
#pragma unroll
            for (int y = x; y < 4; ++y) {
                const bool on = (x > KB || ti > kl) && (y > x || tj >= ti);
                // the products do not wait for the pivot's reciprocal
                if (on) a[x][y] = fma(-(ri[x] * rj[y]), id, a[x][y]);
            }
        // publish row k + 1 into the other buffer: its owners are the 16 threads with ti == (k + 1) % 16
        if (k + 1 < R) {
            if (kl < 15) {
                if (ti == kl + 1) {
#pragma unroll
                    for (int y = KB; y < 4; ++y) rowk[(k + 1) & 1][tj + 16 * y] = a[KB][y];
                }
            } else if (KB < 3) {
                if (ti == 0) {
#pragma unroll
                    for (int y = KB + 1; y < 4; ++y) rowk[(k + 1) & 1][tj + 16 * y] = a[KB < 3 ? KB + 1 : 3][y];
                }
            }
        }
        __syncthreads();
    }
}

__device__ void chol_upper_64(double (*S)[OR_MAX + 1], double* __restrict__ Rm, double* __restrict__ invd, int R,
                              int* bad_out) {
    const int tid = threadIdx.x;
    __shared__ double s_shift2;
    __shared__ double rowk[2][OR_MAX];
    if (tid == 0) {
        double tr = 0.0;
        for (int i = 0; i < R; ++i) tr += S[i][i];
        s_shift2 = ldexp(tr / R, -43);
    }
    __syncthreads();
    const double shift = s_shift2;
    const int ti = tid >> 4, tj = tid & 15;
    double a[4][4];
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int i = ti + 16 * x, j = tj + 16 * y;
            a[x][y] = (i < R && j < R) ? S[i][j] + (i == j ? shift : 0.0) : (i == j ? 1.0 : 0.0);
        }
    int bad = 0;
    // publish row 0
    if (ti == 0) {
#pragma unroll
        for (int y = 0; y < 4; ++y) rowk[0][tj + 16 * y] = a[0][y];
    }
    __syncthreads();
    chol_block<0>(a, rowk, Rm, invd, R, shift, bad, ti, tj, tid);
    if (R > 16) chol_block<1>(a, rowk, Rm, invd, R, shift, bad, ti, tj, tid);
    if (R > 32) chol_block<2>(a, rowk, Rm, invd, R, shift, bad, ti, tj, tid);
    if (R > 48) chol_block<3>(a, rowk, Rm, invd, R, shift, bad, ti, tj, tid);
    if (tid == 0 && bad_out) *bad_out = bad;
}

__device__ __forceinline__ void ps_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
power_step_a_kernel(const double* __restrict__ G, int64_t n, int64_t g_ld, const double* U, int p,
                    int64_t u_ld, double* __restrict__ Z, double* __restrict__ partial, double* __restrict__ ssum,
                    unsigned* __restrict__ counter, double* __restrict__ rinv, int* __restrict__ status, int parallel,
                    int pipelined, double* __restrict__ U_out) {
    extern __shared__ __align__(16) unsigned char ps_smem[];
    typedef double Row[OR_MAX + 1];
    // fused solve (parallel mode): the epoch word is advanced by the factoring CTA once R is in global memory; every
    // CTA reads it before the rendezvous (nobody can advance it before all have passed it)
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(counter + 8);
    // GEMM phase: PS_STAGES stages of [Us: PS_KC x 64][Gs: 16 x PS_GLD]; later, in the factoring CTA only:
    // [S: 64 x 65][Ri: 64 x 65]
    __shared__ int s_last;
    const int tid = threadIdx.x;
    // thread (r, l): row r of the block, columns {2l, 2l+1, 32+2l, 33+2l} — a 16-lane group reads 256 contiguous
    // bytes per LDS.128 (columns 4l..4l+3 per thread put the lanes 32 bytes apart: 4-way bank conflicts)
    const int r = tid >> 4, cl = (tid & 15) * 2, ch = 32 + cl;
    const int64_t row0 = (int64_t)blockIdx.x * PS_ROWS;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, q4 = lane & 3;      // DMMA fragment coordinates
    double cacc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};     // [k4 parity][row tile][2]
    const long long t_start = clock64();
    long long t_loop = t_start, t_z = t_start;
    double* stage0 = reinterpret_cast<double*>(ps_smem);
    const int nchunks = (int)((n + PS_KC - 1) / PS_KC);
    if (pipelined) {
        // Every CTA streams all of U (n x p) and 16 rows of G out of L2; with plain loads the phase was a chain of
        // exposed L2 round trips (70 us of a 150 us step at n = 512).  cp.async keeps PS_STAGES - 1 chunks in flight.
        const int pc = (p + 1) / 2;                                  // 16-byte pieces per row of U
        for (int e = tid; e < PS_STAGES * PS_STAGE_DOUBLES; e += 256) stage0[e] = 0.0;     // columns >= p stay zero
        __syncthreads();
        auto issue = [&](int chunk) {
            double* us = stage0 + (chunk % PS_STAGES) * PS_STAGE_DOUBLES;
            double* gs = us + PS_KC * PS_ULD;
            const int64_t k0 = (int64_t)chunk * PS_KC;
            for (int e = tid; e < PS_KC * pc; e += 256) {
                const int kk = e / pc, c2 = e - kk * pc;
                const bool ok = k0 + kk < n;
                cp_async16(us + kk * PS_ULD + 2 * c2, U + (ok ? (k0 + kk) * u_ld + 2 * c2 : 0), ok);
            }
            for (int e = tid; e < PS_ROWS * (PS_KC / 2); e += 256) {
                const int rr = e / (PS_KC / 2), k2 = e - rr * (PS_KC / 2);
                const bool ok = row0 + rr < n && k0 + 2 * k2 < n;
                cp_async16(gs + rr * PS_GLD + 2 * k2, G + (ok ? (row0 + rr) * g_ld + k0 + 2 * k2 : 0), ok);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int c = 0; c < PS_STAGES - 1; ++c) {
            if (c < nchunks) issue(c); else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int c = 0; c < nchunks; ++c) {
            asm volatile("cp.async.wait_group %0;" ::"n"(PS_STAGES - 2) : "memory");
            __syncthreads();                       // chunk c has landed for everyone; chunk c - 1 is fully consumed
            if (c + PS_STAGES - 1 < nchunks && pipelined != 2) issue(c + PS_STAGES - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
            if (pipelined == 3) continue;
            const double* us = stage0 + (c % PS_STAGES) * PS_STAGE_DOUBLES;
            const double* gs = us + PS_KC * PS_ULD;
            // warp w: columns [8 w, 8 w + 8) of both 8-row tiles, all of k — DMMA.8x8x4 fragments straight from the
            // tiles (3 LDS.64 per 512 FMAs; the FMA version needed 12 shared-memory wavefronts per 256 and was bound
            // by them: 27 of the 35 kclk of this loop)
#pragma unroll 4
            for (int k4 = 0; k4 < PS_KC / 4; ++k4) {
                const int kk = k4 * 4 + q4;
                const double b = us[kk * PS_ULD + warp * 8 + g8];
                const double a0 = gs[g8 * PS_GLD + kk], a1 = gs[(8 + g8) * PS_GLD + kk];
                ps_dmma(cacc[k4 & 1][0][0], cacc[k4 & 1][0][1], a0, b);
                ps_dmma(cacc[k4 & 1][1][0], cacc[k4 & 1][1][1], a1, b);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        t_loop = clock64();
        __syncthreads();                       // nobody reads the stages any more: the Z tile goes over stage 0
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double v = cacc[0][mt][h] + cacc[1][mt][h];
                const int rr = mt * 8 + g8, c = warp * 8 + 2 * q4 + h;
                stage0[rr * OR_MAX + c] = v;
                if (row0 + rr < n && c < p) Z[(row0 + rr) * p + c] = v;
            }
    } else {
        double* Us0 = stage0;                                        // [PS_KC][64]
        double* Gs0 = Us0 + PS_KC * OR_MAX;                          // [PS_ROWS][PS_GLD]
        for (int64_t k0 = 0; k0 < n; k0 += PS_KC) {
            __syncthreads();
            for (int e = tid; e < PS_ROWS * PS_KC; e += 256) {
                const int rr = e >> 6, kk = e & 63;
                Gs0[rr * PS_GLD + kk] = (row0 + rr < n && k0 + kk < n) ? G[(row0 + rr) * g_ld + k0 + kk] : 0.0;
            }
            for (int e = tid; e < PS_KC * OR_MAX; e += 256) {
                const int kk = e >> 6, c = e & 63;
                Us0[e] = (k0 + kk < n && c < p) ? U[(k0 + kk) * u_ld + c] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < PS_KC; ++kk) {
                const double g = Gs0[r * PS_GLD + kk];
                const double2 u0 = *reinterpret_cast<const double2*>(Us0 + kk * OR_MAX + cl);
                const double2 u1 = *reinterpret_cast<const double2*>(Us0 + kk * OR_MAX + ch);
                acc[0] = fma(g, u0.x, acc[0]); acc[1] = fma(g, u0.y, acc[1]);
                acc[2] = fma(g, u1.x, acc[2]); acc[3] = fma(g, u1.y, acc[3]);
            }
        }
    }
    double* Us = stage0;
    __syncthreads();
    // the Z block: to global, and into shared memory (over the U tile) for its Gram partial
    t_z = clock64();
    double* Zs = Us;                                                 // [PS_ROWS][64]
    if (!pipelined) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int c = (t < 2 ? cl : ch) + (t & 1);
            Zs[r * OR_MAX + c] = acc[t];
            if (row0 + r < n && c < p) Z[(row0 + r) * p + c] = acc[t];
        }
    }
    __syncthreads();
    {
        double* mine = partial + (size_t)blockIdx.x * p * p;
        const int a0 = tid >> 4, b0 = tid & 15;                      // outputs (a0 + 16 x, b0 + 16 y)
        // 16 independent accumulators per thread (the 16-deep chains one after the other cost more than the GEMM)
        double sacc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) sacc[x][y] = 0.0;
#pragma unroll 4
        for (int i = 0; i < PS_ROWS; ++i) {
            double za[4], zb[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) { za[x] = Zs[i * OR_MAX + a0 + 16 * x]; zb[x] = Zs[i * OR_MAX + b0 + 16 * x]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) sacc[x][y] = fma(za[x], zb[y], sacc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int a = a0 + 16 * x, b = b0 + 16 * y;
                if (a < p && b < p) mine[a * p + b] = sacc[x][y];
            }
    }
    __threadfence();
    __syncthreads();
    const long long t_gemm = clock64();
    Row* S = reinterpret_cast<Row*>(ps_smem);
    Row* Ri = S + OR_MAX;
    if (parallel) {
        // every CTA is resident (grid <= #SMs): rendezvous, then each CTA sums a slice of the p x p entries over all
        // partials (in block order) into `ssum`; the last CTA to finish its slice gathers S and factors it
        if (tid == 0) {
            atomicAdd(counter, 1u);
            unsigned spins = 0;
            while (atomicAdd(counter, 0u) < gridDim.x) {
                if (++spins > (1u << 26)) asm volatile("trap;");
            }
        }
        __syncthreads();
        __threadfence();
        const int per = (p * p + (int)gridDim.x - 1) / (int)gridDim.x;
        const int e0 = (int)blockIdx.x * per, e1 = min(p * p, e0 + per);
        for (int e = e0 + tid; e < e1; e += 256)
            ssum[e] = ordered_sum_strided<double>(partial + e, (int)gridDim.x, (size_t)p * p);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(counter + 2, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int e = tid; e < p * p; e += 256) S[e / p][e % p] = __ldcg(ssum + e);
            if (tid == 0) { counter[0] = 0u; counter[2] = 0u; }
            __syncthreads();
            if (tid == 0) { g_ps_trace[0] = t_start; g_ps_trace[1] = t_gemm; g_ps_trace[6] = t_loop; g_ps_trace[7] = t_z; }
            PS_TRACE(2);
            chol_upper_64(S, rinv, rinv + (size_t)p * p, p, status);
            PS_TRACE(3);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                unsigned next = s_epoch + 1u;
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(counter + 8), "r"(next) : "memory");
            }
        } else if (tid == 0) {
            // wait for the factor (bounded: trap instead of hanging)
            const unsigned want = s_epoch + 1u;
            unsigned seen = 0, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter + 8) : "memory");
                if (++spins > (1u << 26)) asm volatile("trap;");
            } while (seen != want);
        }
        __syncthreads();
        // U = Z R^{-1} for this CTA's 16 rows, here instead of in a second launch: 16 lanes per row, axpy form (see
        // power_step_b_kernel).  The Z block is still in shared memory — except in the factoring CTA, whose tile was
        // overwritten by S: it reads its rows back from global memory.
        {
            double* Rs = stage0 + 9216;                       // [64][64]: R[k][i] for i > k, else 0 (past S / R^-1 scratch)
            double* idg = Rs + OR_MAX * OR_MAX;
            for (int e = tid; e < OR_MAX * OR_MAX; e += 256) {
                const int k = e >> 6, i = e & 63;
                Rs[e] = (k < p && i < p && i > k) ? __ldcg(rinv + k * p + i) : 0.0;
            }
            if (tid < OR_MAX) idg[tid] = tid < p ? __ldcg(rinv + (size_t)p * p + tid) : 0.0;
            const int lr = tid >> 4, l16 = tid & 15;
            const int64_t grow = row0 + lr;
            double b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = t * 16 + l16;
                if (s_last) b[t] = (grow < n && i < p) ? __ldcg(Z + grow * p + i) : 0.0;
                else b[t] = stage0[lr * OR_MAX + i];
            }
            __syncthreads();
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
#pragma unroll 4
                for (int kk = 0; kk < 16; ++kk) {
                    const int k = sl * 16 + kk;
                    if (k >= p) break;
                    const double qk = __shfl_sync(0xffffffffu, b[sl] * idg[k], kk, 16);
                    if (l16 == kk) b[sl] = qk;
                    const double* c = Rs + k * OR_MAX + l16;
#pragma unroll
                    for (int t = 0; t < 4; ++t) b[t] = fma(-c[t * 16], qk, b[t]);
                }
            }
            if (grow < n)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int i = t * 16 + l16;
                    if (i < p) U_out[grow * u_ld + i] = b[t];
                }
        }
        PS_TRACE(4);
        PS_TRACE(5);
        return;
    } else {
        if (tid == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int e = tid; e < p * p; e += 256)
            S[e / p][e % p] = ordered_sum_strided<double>(partial + e, (int)gridDim.x, (size_t)p * p);
        if (tid == 0) *counter = 0u;
    }
    __syncthreads();
    if (tid == 0) { g_ps_trace[0] = t_start; g_ps_trace[1] = t_gemm; g_ps_trace[6] = t_loop; g_ps_trace[7] = t_z; }
    PS_TRACE(2);
    chol_upper_64(S, rinv, rinv + (size_t)p * p, p, status);          // rinv buffer: [R factor, p x p][1 / diagonal, p]
    PS_TRACE(3);
    PS_TRACE(4);
    PS_TRACE(5);
}

// U = Z R^{-1} without forming R^{-1}: rows of U solve q R = z.  16 lanes per row (axpy form, like the substitution of
// cp_update_kernel): step k — the owner of entry k scales it by 1 / R_kk and broadcasts it inside the group, every
// lane subtracts R[k][i] q_k from the entries i > k it owns.  One shuffle per step on the dependency chain.
__global__ void __launch_bounds__(256)
power_step_b_kernel(const double* __restrict__ Z, int64_t n, int p, const double* __restrict__ rfac, double* __restrict__ U,
                    int64_t u_ld) {
    __shared__ double Rs[OR_MAX * OR_MAX];       // Rs[k][i] = R[k][i] for i > k, else 0       (32 KB)
    __shared__ double invd[OR_MAX];
    const int tid = threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * PS_ROWS + (tid >> 4);
    const int l16 = tid & 15;
    for (int e = tid; e < OR_MAX * OR_MAX; e += 256) {
        const int k = e >> 6, i = e & 63;
        Rs[e] = (k < p && i < p && i > k) ? rfac[k * p + i] : 0.0;
    }
    if (tid < OR_MAX) invd[tid] = tid < p ? rfac[(size_t)p * p + tid] : 0.0;
    double b[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int i = t * 16 + l16;
        b[t] = (row < n && i < p) ? Z[row * p + i] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk) {
            const int k = sl * 16 + kk;
            if (k >= p) break;
            const double qk = __shfl_sync(0xffffffffu, b[sl] * invd[k], kk, 16);
            if (l16 == kk) b[sl] = qk;
            const double* c = Rs + k * OR_MAX + l16;
#pragma unroll
            for (int t = 0; t < 4; ++t) b[t] = fma(-c[t * 16], qk, b[t]);
        }
    }
    if (row < n)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = t * 16 + l16;
            if (i < p) U[row * u_ld + i] = b[t];
        }
}

// ---- symmetric eigendecomposition of a small matrix (n <= 64): cyclic Jacobi, parallel ordering --------------
// One CTA.  A round rotates n/2 disjoint index pairs (round-robin "circle" schedule, n - 1 rounds per sweep):
// angles from the current A, rows of all pairs, then columns of A and of the eigenvector matrix V.  Sweeps repeat
// until off(A)^2 <= 1e-30 ||A||^2 (fp64) or 30 sweeps.  Output: eigenvalues in DESCENDING order, eigenvectors as
// columns, in the caller's dtype.  Used for the Rayleigh-Ritz step of the HOOI subspace iteration.
constexpr int EIG_MAX = 64;
constexpr int EIG_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(EIG_THREADS)
symeig_jacobi_kernel(const T* __restrict__ a, int n, int64_t lda, T* __restrict__ evals, T* __restrict__ evecs,
                     int64_t ldv) {
    extern __shared__ __align__(16) unsigned char eig_smem[];       // A | V, (EIG_MAX x (EIG_MAX + 1)) doubles each
    typedef double ERow[EIG_MAX + 1];
    ERow* A = reinterpret_cast<ERow*>(eig_smem);
    ERow* V = A + EIG_MAX;
    __shared__ double cs_c[EIG_MAX / 2], cs_s[EIG_MAX / 2];
    __shared__ int pr_p[EIG_MAX / 2], pr_q[EIG_MAX / 2];
    __shared__ double red[EIG_THREADS / 32];
    __shared__ double s_off, s_tot;
    __shared__ int order[EIG_MAX];
    const int tid = threadIdx.x;
    const int m = (n + 1) & ~1;                  // even working size; the padding index never rotates (zero couplings)
    for (int e = tid; e < m * m; e += EIG_THREADS) {
        const int i = e / m, j = e - i * m;
        // symmetrise the input: the caller's matrix is U^T G U up to rounding
        A[i][j] = (i < n && j < n) ? 0.5 * ((double)a[i * lda + j] + (double)a[j * lda + i]) : 0.0;
        V[i][j] = i == j ? 1.0 : 0.0;
    }
    __syncthreads();
    const int half = m / 2;
    for (int sweep = 0; sweep < 30; ++sweep) {
        // convergence test
        double off = 0.0, tot = 0.0;
        for (int e = tid; e < m * m; e += EIG_THREADS) {
            const int i = e / m, j = e - i * m;
            const double v = A[i][j] * A[i][j];
            tot += v;
            if (i != j) off += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
        if ((tid & 31) == 0) red[tid >> 5] = off;
        __syncthreads();
        if (tid == 0) { double t = 0.0; for (int i = 0; i < EIG_THREADS / 32; ++i) t += red[i]; s_off = t; }
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = tot;
        __syncthreads();
        if (tid == 0) { double t = 0.0; for (int i = 0; i < EIG_THREADS / 32; ++i) t += red[i]; s_tot = t; }
        __syncthreads();
        if (s_off <= 1e-30 * s_tot) break;
        for (int r = 0; r < m - 1; ++r) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = r; }
                else { p = (r + tid) % (m - 1); q = (r - tid + (m - 1)) % (m - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, sn = 0.0;
                const double apq = A[p][q];
                if (fabs(apq) > 1e-300) {
                    const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    sn = t * c;
                }
                pr_p[tid] = p; pr_q[tid] = q; cs_c[tid] = c; cs_s[tid] = sn;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int e = tid; e < half * m; e += EIG_THREADS) {
                const int pi = e / m, k = e - pi * m;
                const int p = pr_p[pi], q = pr_q[pi];
                const double c = cs_c[pi], sn = cs_s[pi];
                const double ap = A[p][k], aq = A[q][k];
                A[p][k] = c * ap - sn * aq;
                A[q][k] = sn * ap + c * aq;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int e = tid; e < half * m; e += EIG_THREADS) {
                const int pi = e / m, k = e - pi * m;
                const int p = pr_p[pi], q = pr_q[pi];
                const double c = cs_c[pi], sn = cs_s[pi];
                const double ap = A[k][p], aq = A[k][q];
                A[k][p] = c * ap - sn * aq;
                A[k][q] = sn * ap + c * aq;
                const double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - sn * vq;
                V[k][q] = sn * vp + c * vq;
            }
            __syncthreads();
        }
    }
    // descending order (ties: lower index first), then write
    if (tid < n) {
        const double mine = A[tid][tid];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double o = A[j][j];
            if (o > mine || (o == mine && j < tid)) ++rank;
        }
        order[rank] = tid;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += EIG_THREADS) {
        const int i = e / n, j = e - i * n;
        evecs[i * ldv + j] = (T)V[i][order[j]];
    }
    if (tid < n) evals[tid] = (T)A[order[tid]][order[tid]];
}

template <typename T>
int run_pass(const T* z, int64_t rows, int64_t R, int64_t rs, int64_t cs, T* out, int64_t out_ld, void* workspace,
             cudaStream_t stream) {
    // workspace: [counter + status, 256 B][R^{-1}, fp64][partials, fp64]
    unsigned* counter = static_cast<unsigned*>(workspace);
    int* status = reinterpret_cast<int*>(counter + 1);
    double* rinv = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    double* partial = rinv + align_up((size_t)R * R, 32);
    const int nblk = (int)ceil_div(rows, OR_ROWS);
    constexpr int smem = (OR_ROWS + OR_MAX) * (OR_MAX + 1) * (int)sizeof(double);
    static std::atomic<uint64_t> attr_done{0};
    if (ensure_dynamic_smem(orth_gram_chol_kernel<T>, smem, attr_done)) return TLB200_ECUDA;
    orth_gram_chol_kernel<T><<<nblk, OR_THREADS, smem, stream>>>(z, rows, (int)R, rs, cs, partial, counter, rinv, status);
    TLB_CHECK_LAUNCH();
    const int nb2 = (int)ceil_div(rows, 128);
    if (R <= 16) orth_apply_kernel<T, 16><<<nb2, 128, 0, stream>>>(z, rows, (int)R, rs, cs, rinv, out, out_ld);
    else if (R <= 32) orth_apply_kernel<T, 32><<<nb2, 128, 0, stream>>>(z, rows, (int)R, rs, cs, rinv, out, out_ld);
    else orth_apply_kernel<T, 64><<<nb2, 128, 0, stream>>>(z, rows, (int)R, rs, cs, rinv, out, out_ld);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

}  // namespace
}  // namespace tlb200

using namespace tlb200;

// passes = 1: plain Cholesky-QR (orthogonality ~ cond(Z)^2 x 1e-16); passes = 2: a second pass on the result
// ("CholQR2": orthogonal to the rounding of the output type whenever cond(Z)^2 < 1e16).
template <typename T>
int run(const T* z, int64_t rows, int64_t R, int64_t rs, int64_t cs, T* out, int64_t out_ld, void* workspace, int passes,
        cudaStream_t stream) {
    int st = run_pass<T>(z, rows, R, rs, cs, out, out_ld, workspace, stream);
    if (st || passes < 2) return st;
    return run_pass<T>(out, rows, R, out_ld, 1, out, out_ld, workspace, stream);
}

extern "C" size_t tlb200_subspace_iterate_workspace_bytes(int64_t n, int64_t p) {
    if (n < 1 || p < 1 || p > OR_MAX) return 0;
    return 256 + sizeof(double) * ((size_t)n * p + 2 * align_up((size_t)p * p + OR_MAX, 32) + (size_t)ceil_div(n, PS_ROWS) * p * p) + 256;
}

extern "C" int tlb200_subspace_iterate(const void* g, int64_t n, int64_t g_ld, void* u, int64_t p, int64_t u_ld, int steps,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    if (!g || !u || !workspace || n < 1 || p < 1 || g_ld < n || u_ld < p || steps < 0) return TLB200_EINVAL;
    if (p > OR_MAX || n < p) return TLB200_EUNSUPPORTED;
    if (workspace_bytes < tlb200_subspace_iterate_workspace_bytes(n, p)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned* counter = static_cast<unsigned*>(workspace);
    int* status = reinterpret_cast<int*>(counter + 1);
    double* rinv = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
    double* ssum = rinv + align_up((size_t)p * p + OR_MAX, 32);
    double* z = ssum + align_up((size_t)p * p + OR_MAX, 32);
    double* partial = z + (size_t)n * p;
    constexpr int smem_a = PS_STAGES * PS_STAGE_DOUBLES * (int)sizeof(double);    // 123 648 bytes: 3 GEMM stages
    static_assert(smem_a >= 2 * OR_MAX * (OR_MAX + 1) * (int)sizeof(double), "the factoring phase reuses the stages");
    // cp.async moves 16-byte pieces: rows of G and U must start 16-byte aligned
    int pipelined = g_ld % 2 == 0 && u_ld % 2 == 0 && reinterpret_cast<uintptr_t>(g) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(u) % 16 == 0;
    if (pipelined) {     // perf triage: 2 = load only the first stages (compute-only timing), 3 = loads only
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("TLB200_PS_DEBUG"); dbg = e ? atoi(e) : 0; }
        if (dbg == 2 || dbg == 3) pipelined = dbg;
    }
    static std::atomic<uint64_t> attr_done{0};
    if (ensure_dynamic_smem(power_step_a_kernel, smem_a, attr_done)) return TLB200_ECUDA;
    const int nblk = (int)ceil_div(n, PS_ROWS);
    for (int it = 0; it < steps; ++it) {
        const int parallel = nblk > 1 && nblk <= kNumSMs;
        power_step_a_kernel<<<nblk, 256, smem_a, s>>>((const double*)g, n, g_ld, (const double*)u, (int)p, u_ld, z, partial,
                                                      ssum, counter, rinv, status, parallel, pipelined, (double*)u);
        TLB_CHECK_LAUNCH();
        if (!parallel) {           // (parallel: every CTA solved its own rows after the factoring CTA published R)
            power_step_b_kernel<<<nblk, 256, 0, s>>>(z, n, (int)p, rinv, (double*)u, u_ld);
            TLB_CHECK_LAUNCH();
        }
    }
    return TLB200_OK;
}

extern "C" int tlb200_symeig(const void* a, int64_t n, int64_t lda, int dtype, void* evals, void* evecs, int64_t ldv,
                             void* stream) {
    if (!a || !evals || !evecs || n < 1 || lda < n || ldv < n || !dtype_valid(dtype)) return TLB200_EINVAL;
    if (n > EIG_MAX) return TLB200_EUNSUPPORTED;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    constexpr int smem = 2 * EIG_MAX * (EIG_MAX + 1) * (int)sizeof(double);
    static std::atomic<uint64_t> done_f{0}, done_d{0};
    if (dtype == TLB200_F32) {
        if (ensure_dynamic_smem(symeig_jacobi_kernel<float>, smem, done_f)) return TLB200_ECUDA;
        symeig_jacobi_kernel<float><<<1, EIG_THREADS, smem, s>>>((const float*)a, (int)n, lda, (float*)evals, (float*)evecs, ldv);
    } else {
        if (ensure_dynamic_smem(symeig_jacobi_kernel<double>, smem, done_d)) return TLB200_ECUDA;
        symeig_jacobi_kernel<double><<<1, EIG_THREADS, smem, s>>>((const double*)a, (int)n, lda, (double*)evals, (double*)evecs, ldv);
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

extern "C" size_t tlb200_orthonormalize_workspace_bytes(int64_t rows, int64_t rank) {
    if (rows < 1 || rank < 1 || rank > OR_MAX) return 0;
    return 256 + sizeof(double) * (align_up((size_t)rank * rank, 32) + (size_t)ceil_div(rows, OR_ROWS) * rank * rank) + 256;
}

extern "C" int tlb200_orthonormalize(const void* z, int64_t rows, int64_t rank, int64_t row_stride, int64_t col_stride,
                                     int dtype, void* out, int64_t out_ld, int passes, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    if (!z || !out || !workspace || rows < 1 || rank < 1 || out_ld < rank || !dtype_valid(dtype)) return TLB200_EINVAL;
    if (rank > OR_MAX || rows < rank) return TLB200_EUNSUPPORTED;
    if (workspace_bytes < tlb200_orthonormalize_workspace_bytes(rows, rank)) return TLB200_EWORKSPACE;
    set_last_path("simt");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return run<float>((const float*)z, rows, rank, row_stride, col_stride, (float*)out, out_ld, workspace, passes, s);
    return run<double>((const double*)z, rows, rank, row_stride, col_stride, (double*)out, out_ld, workspace, passes, s);
}

// perf triage hook (not part of the public ABI): copies the 8 phase timestamps of the last factoring CTA to the host
extern "C" int tlb200_debug_subspace_trace(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, tlb200::g_ps_trace, sizeof(long long) * 8) == cudaSuccess ? 0 : -3;
}
