// Shared helpers for the tlb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stddef.h>
#include "../../include/tlb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ != 1000)
#error "tlb200 kernels are written for sm_100a (B200) only"
#endif

namespace tlb200 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Name of the kernel family the last API call on this thread dispatched to.
void set_last_path(const char* name);

inline int cuda_ok(cudaError_t e) { return e == cudaSuccess ? TLB200_OK : TLB200_ECUDA; }

void count_launch();

// After every kernel launch: count it and surface launch errors.
#define TLB_CHECK_LAUNCH()                                  \
    do {                                                    \
        ::tlb200::count_launch();                           \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return TLB200_ECUDA;        \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (kernel, device)
// pair, so it is tracked per device — one bit per ordinal in a per-kernel atomic mask (a process may drive several
// GPUs from several threads); ordinals >= 64 simply set it on every launch.
template <typename K>
inline int ensure_dynamic_smem(K kernel, int bytes, std::atomic<uint64_t>& done) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return TLB200_ECUDA;
    const uint64_t bit = dev >= 0 && dev < 64 ? (1ull << dev) : 0;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return TLB200_OK;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return TLB200_ECUDA;
    if (bit) done.fetch_or(bit, std::memory_order_release);
    return TLB200_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t dtype_size(int dtype) { return dtype == TLB200_F64 ? 8 : 4; }
inline bool dtype_valid(int dtype) { return dtype == TLB200_F32 || dtype == TLB200_F64; }

// Carve sub-buffers out of a caller-provided workspace.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    size_t used() const { return align_up(off, 256); }
};

// A Khatri-Rao table that would be a plain copy of ONE unweighted factor matrix already in the table's layout
// (row-major, row stride = padded rank, 16-byte aligned): use the factor in place and skip the prep launch.
template <typename T>
inline const T* table_is_factor(const T* const* factors, const int64_t* frs, const int64_t* fcs, int first, int count,
                                const T* weights, int64_t rank, int64_t rank_padded) {
    if (count != 1 || weights != nullptr || rank != rank_padded) return nullptr;
    if (fcs[first] != 1 || frs[first] != rank_padded) return nullptr;
    if (reinterpret_cast<uintptr_t>(factors[first]) % 16) return nullptr;
    return factors[first];
}

#ifdef __CUDACC__
// Sum of `n` values spaced `stride` elements apart (per-CTA partial results in global memory), in index order, with
// the loads issued 16 at a time: a plain `acc += ldcg(p[b])` loop is a chain of L2 round trips (~0.5 us each) that
// made the "last CTA sums the partials" tails of several kernels cost more than the work they finish.
template <typename T>
__device__ __forceinline__ T ordered_sum_strided(const T* __restrict__ p, int n, size_t stride) {
    T acc = T(0);
    for (int b0 = 0; b0 < n; b0 += 16) {
        T v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = b0 + i < n ? __ldcg(p + (size_t)(b0 + i) * stride) : T(0);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i];
    }
    return acc;
}
#endif

// ---- internal launchers shared between translation units ------------------
// Khatri-Rao of `nmats` matrices into out[(rows), ld]; columns [rank, pad_cols) are
// written as zero.  Unlike the public entry point, weights are applied even for a
// single matrix.
template <typename T>
int launch_khatri_rao(const T* const* mats, const int64_t* rows, const int64_t* row_stride,
                      const int64_t* col_stride, int nmats, int64_t rank, const T* weights,
                      const T* mask, T* out, int64_t out_ld, int64_t pad_cols,
                      cudaStream_t stream);

// Transposed + zero-padded: out[c * rows_padded + row], c < pad_cols, row < rows_padded (tcgen05 engine's Q table).
template <typename T>
int launch_khatri_rao_t(const T* const* mats, const int64_t* rows, const int64_t* row_stride,
                        const int64_t* col_stride, int nmats, int64_t rank, const T* weights, T* out,
                        int64_t rows_padded, int64_t pad_cols, T* out_lo, cudaStream_t stream);

}  // namespace tlb200
