// One-shot all-reduce over NVLink peer memory for the small per-mode partials of the sharded CP-ALS sweep
// (an R x R Gram partial and an I_n x R MTTKRP partial per mode, 16 KB - 0.5 MB).
//
// The reference has no distributed path; the round-1 driver used two NCCL all-reduces per sweep.  At these sizes a
// collective is pure latency (launch + protocol + rank skew), and NCCL kernels inside a captured CUDA graph tie the
// graph's lifetime to the communicator's.  Here every rank owns one "symmetric" buffer that its peers map through
// CUDA IPC, and ONE kernel per all-reduce does everything:
//
//   push    every CTA copies its part of the local vector into slot [parity][my rank] of EVERY rank's buffer
//           (plain 16-byte stores over NVLink / NVSwitch; the local copy is one of them);
//   publish the last CTA to finish pushing (ticket) fences to system scope and writes the epoch number into
//           flag [parity][my rank] of every peer (st.release.sys);
//   wait    every CTA spins (ld.acquire.sys, bounded: it traps instead of hanging) until all `world` local flags
//           of this parity carry the epoch;
//   reduce  out[i] = slot[0][i] + slot[1][i] + ... in RANK ORDER from local memory: every rank computes the same
//           bits, and the result does not depend on arrival order.
//
// The epoch lives in device memory and is advanced by the kernel itself, so a captured graph can be replayed.
// Slots alternate with the epoch's parity: a rank can only be two epochs ahead of a peer's reads if the peer has
// published the epoch in between, i.e. finished reading — no extra barrier is needed.
#include "common.cuh"

#include <cstring>

namespace tlb200 {
namespace {

constexpr int kCommMaxWorld = 16;
constexpr int kCommThreads = 256;
constexpr size_t kCommHeader = 1024;        // epoch, tickets, flags[2][16]

struct CommHeader {
    unsigned long long epoch;               // all-reduces completed on this rank
    unsigned int push_ticket;
    unsigned int done_ticket;
    unsigned long long pad;
    unsigned long long flags[2][kCommMaxWorld];
};
static_assert(sizeof(CommHeader) <= kCommHeader, "header");

struct CommPeers {
    unsigned char* buf[kCommMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// What this rank contributes: a plain vector, or — fused split-K reduction — the unsummed partials of an MTTKRP
// ([splits][rows][ld], summed in split order while they are pushed) optionally followed by a plain tail (the R x R
// Gram partial that travels in the same exchange).
template <typename T>
struct CommSource {
    const T* in;            // plain vector of `count` elements, or the partials
    int splits;             // 1: plain
    int64_t split_stride;
    int64_t cols, ld;       // partials: element e <-> (row e / cols, column e % cols) at row * ld + column
    int64_t count1;         // elements coming from `in`
    const T* in2;           // plain tail (count - count1 elements) or null
};

template <typename T>
__device__ __forceinline__ T comm_fetch(const CommSource<T>& src, int64_t e) {
    if (e >= src.count1) return src.in2[e - src.count1];
    if (src.splits == 1 && src.ld == src.cols) return src.in[e];
    const int64_t row = e / src.cols, c = e - row * src.cols;
    const T* p = src.in + row * src.ld + c;
    return src.splits == 1 ? *p : ordered_sum_strided<T>(p, src.splits, (size_t)src.split_stride);
}

// slot_bytes: size of one rank's slot; data region of a buffer = [2 parities][world slots][slot_bytes]
template <typename T>
__global__ void __launch_bounds__(kCommThreads)
allreduce_oneshot_kernel(const CommSource<T> src, T* __restrict__ out, int64_t count, CommPeers peers, int world, int rank,
                         size_t slot_bytes) {
    const T* __restrict__ in = src.in;
    __shared__ unsigned long long s_epoch;
    __shared__ int s_last;
    CommHeader* me = reinterpret_cast<CommHeader*>(peers.buf[rank]);
    const int tid = threadIdx.x;
    if (tid == 0) s_epoch = me->epoch + 1;          // advanced only by the last CTA of this launch, at the very end
    __syncthreads();
    const unsigned long long epoch = s_epoch;
    const int par = (int)(epoch & 1ull);
    const size_t my_slot = kCommHeader + ((size_t)par * world + rank) * slot_bytes;
    constexpr int VW = 16 / sizeof(T);
    const int64_t nvec = count / VW;
    const bool plain = src.splits == 1 && src.ld == src.cols && src.in2 == nullptr;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    const bool vec_in = plain && (reinterpret_cast<uintptr_t>(in) % 16 == 0);
    const int64_t stride = (int64_t)gridDim.x * kCommThreads;
    const int64_t first = (int64_t)blockIdx.x * kCommThreads + tid;

    // ---- push ----
    if (vec_ok) {
        for (int64_t v = first; v < nvec; v += stride) {
            int4 x;
            if (vec_in) {
                x = reinterpret_cast<const int4*>(in)[v];
            } else {
                T t[VW];
#pragma unroll
                for (int k = 0; k < VW; ++k) t[k] = comm_fetch<T>(src, v * VW + k);
                memcpy(&x, t, 16);
            }
            for (int p = 0; p < world; ++p) reinterpret_cast<int4*>(peers.buf[p] + my_slot)[v] = x;
        }
        for (int64_t e = nvec * VW + first; e < count; e += stride) {
            const T x = comm_fetch<T>(src, e);
            for (int p = 0; p < world; ++p) reinterpret_cast<T*>(peers.buf[p] + my_slot)[e] = x;
        }
    } else {
        for (int64_t e = first; e < count; e += stride) {
            const T x = comm_fetch<T>(src, e);
            for (int p = 0; p < world; ++p) reinterpret_cast<T*>(peers.buf[p] + my_slot)[e] = x;
        }
    }
    __threadfence_system();
    __syncthreads();
    // ---- publish (last CTA of this rank) ----
    if (tid == 0) s_last = atomicAdd(&me->push_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (tid < world) st_release_sys(&reinterpret_cast<CommHeader*>(peers.buf[tid])->flags[par][rank], epoch);
    }
    // ---- wait: all ranks' slots of this parity are complete in MY buffer ----
    if (tid < world) {
        const unsigned long long* f = &me->flags[par][tid];
        unsigned long long spins = 0;
        while (ld_acquire_sys(f) < epoch) {
            if (++spins > (1ull << 27)) asm volatile("trap;");     // a peer never arrived: fail loudly, never hang
        }
    }
    __syncthreads();
    // ---- reduce in rank order, from local memory ----
    const unsigned char* base = peers.buf[rank] + kCommHeader + (size_t)par * world * slot_bytes;
    if (vec_ok) {
        for (int64_t v = first; v < nvec; v += stride) {
            T acc[VW];
            {
                const int4 x = __ldcg(reinterpret_cast<const int4*>(base) + v);
                memcpy(acc, &x, 16);
            }
            for (int p = 1; p < world; ++p) {
                const int4 x = __ldcg(reinterpret_cast<const int4*>(base + (size_t)p * slot_bytes) + v);
                T t[VW];
                memcpy(t, &x, 16);
#pragma unroll
                for (int k = 0; k < VW; ++k) acc[k] += t[k];
            }
            int4 o;
            memcpy(&o, acc, 16);
            reinterpret_cast<int4*>(out)[v] = o;
        }
        for (int64_t e = nvec * VW + first; e < count; e += stride) {
            T acc = __ldcg(reinterpret_cast<const T*>(base) + e);
            for (int p = 1; p < world; ++p) acc += __ldcg(reinterpret_cast<const T*>(base + (size_t)p * slot_bytes) + e);
            out[e] = acc;
        }
    } else {
        for (int64_t e = first; e < count; e += stride) {
            T acc = __ldcg(reinterpret_cast<const T*>(base) + e);
            for (int p = 1; p < world; ++p) acc += __ldcg(reinterpret_cast<const T*>(base + (size_t)p * slot_bytes) + e);
            out[e] = acc;
        }
    }
    // ---- the last CTA to finish advances the epoch and re-arms the tickets ----
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(&me->done_ticket, 1u) == gridDim.x - 1) {
            me->push_ticket = 0u;
            me->done_ticket = 0u;
            me->epoch = epoch;
            __threadfence();
        }
    }
}

}  // namespace
}  // namespace tlb200

using namespace tlb200;

extern "C" size_t tlb200_comm_buffer_bytes(int world, size_t max_payload_bytes) {
    if (world < 1 || world > kCommMaxWorld) return 0;
    const size_t slot = align_up(max_payload_bytes, 256);
    return kCommHeader + (size_t)2 * world * slot;
}

// Allocate this rank's symmetric buffer (zero-filled) and export its CUDA IPC handle (64 bytes).
extern "C" int tlb200_comm_alloc(size_t bytes, void** ptr, void* ipc_handle_out) {
    if (!ptr || !ipc_handle_out || bytes < kCommHeader) return TLB200_EINVAL;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return TLB200_ECUDA;
    if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return TLB200_ECUDA; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return TLB200_ECUDA; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &h, sizeof(h));
    *ptr = p;
    return TLB200_OK;
}

extern "C" int tlb200_comm_open(const void* ipc_handle, void** peer_ptr) {
    if (!ipc_handle || !peer_ptr) return TLB200_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return TLB200_ECUDA;
    *peer_ptr = p;
    return TLB200_OK;
}

extern "C" int tlb200_comm_close(void* peer_ptr) {
    return peer_ptr && cudaIpcCloseMemHandle(peer_ptr) == cudaSuccess ? TLB200_OK : TLB200_ECUDA;
}

extern "C" int tlb200_comm_free(void* ptr) { return ptr && cudaFree(ptr) == cudaSuccess ? TLB200_OK : TLB200_ECUDA; }

namespace tlb200 {
namespace {
template <typename T>
int launch_allreduce(const CommSource<T>& src, T* out, int64_t count, void* const* bufs, int world, int rank,
                     size_t max_payload_bytes, cudaStream_t s) {
    const size_t slot = align_up(max_payload_bytes, 256);
    if ((size_t)count * sizeof(T) > slot) return TLB200_EWORKSPACE;
    if (count == 0) return TLB200_OK;
    CommPeers peers;
    for (int p = 0; p < kCommMaxWorld; ++p) peers.buf[p] = p < world ? static_cast<unsigned char*>(bufs[p]) : nullptr;
    for (int p = 0; p < world; ++p)
        if (!peers.buf[p]) return TLB200_EINVAL;
    set_last_path("p2p");
    const int64_t vecs = ceil_div((int64_t)count * (int64_t)sizeof(T), 16);
    // every CTA must be resident (they wait for each other): at most 128.  A fused split-K sum wants the threads
    // (one vector each, `splits` loads behind it), a plain push wants few CTAs (>= 4 vectors per thread)
    int grid = (int)ceil_div(vecs, kCommThreads * (src.splits > 1 ? 1 : 4));
    if (grid < 1) grid = 1;
    const int cap = src.splits > 1 ? 128 : 32;
    if (grid > cap) grid = cap;
    allreduce_oneshot_kernel<T><<<grid, kCommThreads, 0, s>>>(src, out, count, peers, world, rank, slot);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}
}  // namespace
}  // namespace tlb200

// Sum `count` elements over `world` ranks; in may equal out.  bufs[p] = rank p's symmetric buffer as mapped in THIS
// process (bufs[rank] = the local allocation), each of tlb200_comm_buffer_bytes(world, max_payload_bytes) bytes.
// Every rank must issue the same sequence of calls on its buffer.
extern "C" int tlb200_allreduce_oneshot(const void* in, void* out, int64_t count, int dtype, void* const* bufs, int world,
                                        int rank, size_t max_payload_bytes, void* stream) {
    if (!in || !out || !bufs || count < 0 || world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world ||
        !dtype_valid(dtype))
        return TLB200_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto go = [&](auto tag) {
        using T = decltype(tag);
        CommSource<T> src;
        src.in = static_cast<const T*>(in); src.splits = 1; src.split_stride = 0; src.cols = count; src.ld = count;
        src.count1 = count; src.in2 = nullptr;
        return launch_allreduce<T>(src, static_cast<T*>(out), count, bufs, world, rank, max_payload_bytes, s);
    };
    return dtype == TLB200_F32 ? go(float()) : go(double());
}

// The same exchange with the split-K reduction of an MTTKRP fused into its push phase: rank-local contribution =
// sum over splits of `m` (tlb200_partials_t), followed by `tail_count` plain elements of `tail` (may be null: the
// R x R Gram partial that shares the exchange).  out: contiguous, m->rows * m->rank + tail_count elements.
extern "C" int tlb200_allreduce_partials(const tlb200_partials_t* m, const void* tail, int64_t tail_count, void* out,
                                         int dtype, void* const* bufs, int world, int rank, size_t max_payload_bytes,
                                         void* stream) {
    if (!m || !m->data || !out || !bufs || tail_count < 0 || (tail_count > 0 && !tail) || world < 1 || world > kCommMaxWorld ||
        rank < 0 || rank >= world || !dtype_valid(dtype) || m->splits < 1 || m->rows < 1 || m->rank < 1 || m->ld < m->rank)
        return TLB200_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t count1 = m->rows * m->rank;
    auto go = [&](auto tag) {
        using T = decltype(tag);
        CommSource<T> src;
        src.in = static_cast<const T*>(m->data); src.splits = (int)m->splits; src.split_stride = m->split_stride;
        src.cols = m->rank; src.ld = m->ld; src.count1 = count1; src.in2 = tail_count > 0 ? static_cast<const T*>(tail) : nullptr;
        return launch_allreduce<T>(src, static_cast<T*>(out), count1 + tail_count, bufs, world, rank, max_payload_bytes, s);
    };
    return dtype == TLB200_F32 ? go(float()) : go(double());
}
