// unfold / fold: bit-exact permuting copies.
//
// Reference: tensorly/base.py:39-53 (unfold) and :56-79 (fold).  For a C-contiguous
// N-way tensor viewed as X[L, J, T] (L = prod of modes before `mode`, J = shape[mode],
// T = prod of modes after) unfold is out[J, L, T] = X[L, J, T]; fold is the inverse.
// Both are "swap the two leading dims of a 3-D array whose innermost run of T elements
// stays contiguous", so one pair of kernels serves both.
#include "common.cuh"

namespace tlb200 {
namespace {

// ---- long inner runs: copy rows of T elements, 128-bit when aligned --------------
// in  [D0, D1, T] -> out [D1, D0, T].  blockDim = (TX, TY): TY rows per CTA.
template <typename V>
__global__ void __launch_bounds__(256)
swap01_rows_kernel(const V* __restrict__ in, V* __restrict__ out, int64_t D0, int64_t D1,
                   int64_t Tv /* run length in units of V */, int64_t nrows) {
    int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (row >= nrows) return;
    int64_t d0 = row / D1, d1 = row - d0 * D1;
    const V* src = in + row * Tv;
    V* dst = out + (d1 * D0 + d0) * Tv;
    for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < Tv;
         t += (int64_t)gridDim.y * blockDim.x)
        dst[t] = __ldg(src + t);
}

// ---- short inner runs (T < 32, incl. T == 1: a plain transpose) -------------------
// Tile: 32 values of d0  x  TJ values of d1 (TJ*T >= 32) staged through shared memory so
// that both the global reads (runs of TJ*T elements) and writes (runs of 32*T elements)
// are contiguous.
template <typename T>
__global__ void __launch_bounds__(256)
swap01_tile_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t D0, int64_t D1,
                   int Tt, int TJ) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    const int W = TJ * Tt;       // elements per tile row (one d0)
    const int ldw = W + 1;
    const int64_t d0_0 = (int64_t)blockIdx.y * 32;
    const int64_t d1_0 = (int64_t)blockIdx.x * TJ;
    const int tid = threadIdx.x;
    const int64_t w_valid = min((int64_t)TJ, D1 - d1_0) * Tt;
    // load: 32 rows x W
    for (int e = tid; e < 32 * W; e += 256) {
        int r = e / W, c = e - r * W;
        int64_t d0 = d0_0 + r;
        if (d0 < D0 && c < w_valid) tile[r * ldw + c] = in[(d0 * D1 + d1_0) * Tt + c];
    }
    __syncthreads();
    // store: for each d1 in tile a run of 32*T elements
    const int run = 32 * Tt;
    const int64_t d0_valid = min((int64_t)32, D0 - d0_0);
    for (int e = tid; e < TJ * run; e += 256) {
        int j = e / run, q = e - j * run;
        int r = q / Tt, t = q - r * Tt;
        int64_t d1 = d1_0 + j;
        if (d1 < D1 && r < d0_valid) out[(d1 * D0 + d0_0) * Tt + q] = tile[r * ldw + j * Tt + t];
    }
}

// ---- T == 1: plain 2-D transpose, 64 x 64 tiles through shared memory -------------------
// in [D0][D1] -> out [D1][D0]; both sides move whole 128/256-byte row pieces.  One-dimensional grid (tiles of the
// longer dimension can exceed the 65535 limit of grid.y).
template <typename T>
__global__ void __launch_bounds__(256)
transpose64_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t D0, int64_t D1, int64_t tiles1) {
    __shared__ T tile[64][65];
    const int64_t b0 = blockIdx.x / tiles1, b1 = blockIdx.x - b0 * tiles1;
    const int64_t d0_0 = b0 * 64, d1_0 = b1 * 64;
    const int c = threadIdx.x & 63, r0 = threadIdx.x >> 6;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int r = r0 + 4 * k;
        if (d0_0 + r < D0 && d1_0 + c < D1) tile[r][c] = __ldg(in + (d0_0 + r) * D1 + d1_0 + c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int r = r0 + 4 * k;
        if (d1_0 + r < D1 && d0_0 + c < D0) out[(d1_0 + r) * D0 + d0_0 + c] = tile[c][r];
    }
}

// ---- medium inner runs (rows of 128 B .. 2 KB): 8 x 8 rows per CTA ------------------------
// Copying row by row scatters 256-byte pieces on the write side (output rows of consecutive input rows are
// D0*T apart).  A CTA that takes 8 values of d0 x 8 values of d1 reads 8 contiguous rows per d0 and writes 8
// contiguous rows per d1: both sides see runs of 8 rows.  V = 16-byte vector.
template <typename V>
__global__ void __launch_bounds__(256)
swap01_rowtile_kernel(const V* __restrict__ in, V* __restrict__ out, int64_t D0, int64_t D1, int Tv, int64_t tiles1) {
    const int64_t b0 = blockIdx.x / tiles1, b1 = blockIdx.x - b0 * tiles1;
    const int64_t d0_0 = b0 * 8, d1_0 = b1 * 8;
    const int n0 = (int)min((int64_t)8, D0 - d0_0), n1 = (int)min((int64_t)8, D1 - d1_0);
    const int per0 = n1 * Tv;                  // vectors per d0 (contiguous in the input)
    for (int e = threadIdx.x; e < n0 * per0; e += 256) {
        const int i0 = e / per0, q = e - i0 * per0;
        const int i1 = q / Tv, v = q - i1 * Tv;
        out[((d1_0 + i1) * D0 + d0_0 + i0) * Tv + v] = __ldg(in + ((d0_0 + i0) * D1 + d1_0) * Tv + q);
    }
}

template <typename T>
int swap01(const T* in, T* out, int64_t D0, int64_t D1, int64_t Tt, cudaStream_t stream) {
    if (D0 == 0 || D1 == 0 || Tt == 0) return TLB200_OK;
    if (D0 == 1 || D1 == 1) {  // no permutation at all
        if (in != out) {
            if (cudaMemcpyAsync(out, in, sizeof(T) * D0 * D1 * Tt, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
                return TLB200_ECUDA;
        }
        return TLB200_OK;
    }
    if (Tt == 1) {
        const int64_t tiles0 = ceil_div(D0, 64), tiles1 = ceil_div(D1, 64);
        if (tiles0 * tiles1 > 0x7fffffffLL) return TLB200_EUNSUPPORTED;
        transpose64_kernel<T><<<(unsigned)(tiles0 * tiles1), 256, 0, stream>>>(in, out, D0, D1, tiles1);
        TLB_CHECK_LAUNCH();
        return TLB200_OK;
    }
    if (Tt >= 32 && Tt * sizeof(T) < 2048 && (Tt * sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0 &&
        reinterpret_cast<uintptr_t>(out) % 16 == 0) {
        const int64_t tiles0 = ceil_div(D0, 8), tiles1 = ceil_div(D1, 8);
        if (tiles0 * tiles1 <= 0x7fffffffLL) {
            swap01_rowtile_kernel<int4><<<(unsigned)(tiles0 * tiles1), 256, 0, stream>>>(
                reinterpret_cast<const int4*>(in), reinterpret_cast<int4*>(out), D0, D1, (int)(Tt * sizeof(T) / 16), tiles1);
            TLB_CHECK_LAUNCH();
            return TLB200_OK;
        }
    }
    if (Tt >= 32) {
        const int64_t nrows = D0 * D1;
        if (nrows > 0x7fffffffLL) return TLB200_EUNSUPPORTED;
        const bool vec16 = (Tt * sizeof(T)) % 16 == 0 && (reinterpret_cast<uintptr_t>(in) % 16 == 0) &&
                           (reinterpret_cast<uintptr_t>(out) % 16 == 0);
        int64_t Tv = vec16 ? (int64_t)(Tt * sizeof(T) / 16) : Tt;
        int tx = 32;
        while (tx < 256 && tx < Tv) tx <<= 1;
        int ty = 256 / tx;
        dim3 block(tx, ty);
        int64_t gy = ceil_div(Tv, (int64_t)tx * 8);
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        dim3 grid((unsigned)ceil_div(nrows, ty), (unsigned)gy);
        if (vec16)
            swap01_rows_kernel<int4><<<grid, block, 0, stream>>>(reinterpret_cast<const int4*>(in),
                                                               reinterpret_cast<int4*>(out), D0, D1, Tv, nrows);
        else
            swap01_rows_kernel<T><<<grid, block, 0, stream>>>(in, out, D0, D1, Tv, nrows);
    } else {
        int TJ = (int)ceil_div(32, Tt);
        int W = TJ * (int)Tt;
        size_t smem = sizeof(T) * 32 * (W + 1);
        int64_t gx = ceil_div(D1, TJ), gy = ceil_div(D0, 32);
        if (gx > 0x7fffffffLL || gy > 65535) {
            // fall back to flipping the roles so that the long dim is on grid.x: not needed
            // for supported shapes (D0 tiles of 32 beyond 65535*32 = 2M rows) -> reject.
            return TLB200_EUNSUPPORTED;
        }
        dim3 grid((unsigned)gx, (unsigned)gy);
        swap01_tile_kernel<T><<<grid, 256, smem, stream>>>(in, out, D0, D1, (int)Tt, TJ);
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

int split_shape(const int64_t* shape, int ndim, int mode, int64_t* L, int64_t* J, int64_t* T) {
    if (!shape || ndim < 1 || ndim > TLB200_MAX_NDIM || mode < 0 || mode >= ndim) return TLB200_EINVAL;
    int64_t l = 1, t = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] < 0) return TLB200_EINVAL;
        if (i < mode) l *= shape[i];
        if (i > mode) t *= shape[i];
    }
    *L = l; *J = shape[mode]; *T = t;
    return TLB200_OK;
}

}  // namespace
}  // namespace tlb200

using namespace tlb200;

extern "C" int tlb200_unfold(const void* x, const int64_t* shape, int ndim, int mode, int dtype,
                             void* out, void* stream) {
    int64_t L, J, T;
    int st = split_shape(shape, ndim, mode, &L, &J, &T);
    if (st) return st;
    if (!dtype_valid(dtype) || (!x && L * J * T) || (!out && L * J * T)) return TLB200_EINVAL;
    set_last_path("copy");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // in [L, J, T] -> out [J, L, T]
    if (dtype == TLB200_F32) return swap01<float>((const float*)x, (float*)out, L, J, T, s);
    return swap01<double>((const double*)x, (double*)out, L, J, T, s);
}

extern "C" int tlb200_fold(const void* unfolded, const int64_t* shape, int ndim, int mode, int dtype,
                           void* out, void* stream) {
    int64_t L, J, T;
    int st = split_shape(shape, ndim, mode, &L, &J, &T);
    if (st) return st;
    if (!dtype_valid(dtype) || (!unfolded && L * J * T) || (!out && L * J * T)) return TLB200_EINVAL;
    set_last_path("copy");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // in [J, L, T] -> out [L, J, T]
    if (dtype == TLB200_F32) return swap01<float>((const float*)unfolded, (float*)out, J, L, T, s);
    return swap01<double>((const double*)unfolded, (double*)out, J, L, T, s);
}
