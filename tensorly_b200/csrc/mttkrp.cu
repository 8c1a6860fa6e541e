// MTTKRP: out[i_mode, r] = sum x[i_0..i_{N-1}] * w[r] * prod_{n != mode} F_n[i_n, r]
//
// Reference: tensorly/tenalg/core_tenalg/mttkrp.py:47-49 materialises the Khatri-Rao
// matrix (prod_{n != mode} I_n x R) and a permuted copy of the tensor, then calls one
// GEMM.  Here the tensor is viewed — without any copy — as X[a, j, b] (j = the kept
// mode) and the Khatri-Rao row of contraction index (a, b) is P[a, :] * Q[b, :], where
// P and Q are two small tables (Khatri-Rao products of the factors of the modes folded
// into `a` resp. `b`, weights folded into the first).  KR tiles are formed in shared
// memory as the tensor streams by; the tensor is read from HBM exactly once.
//
//   mode 0      : X[j, (a, b)]   P = KR(F_1..F_s),   Q = KR(F_{s+1}..F_{N-1})
//   middle mode : X[a, j, b]     P = KR(F_0..F_{n-1}), Q = KR(F_{n+1}..F_{N-1})
//   last mode   : X[(a, b), j]   P = KR(F_0..F_s),   Q = KR(F_{s+1}..F_{N-2})
//
// For a 3-way tensor P and Q are simply the two other factor matrices.
#include "common.cuh"
#include "stream_gemm.cuh"
#include "mttkrp_tc.cuh"
#include "hf_split.cuh"

namespace tlb200 {

static int make_plan_one(const int64_t* shape, int ndim, int mode, int64_t rank, int dtype, int path,
                         tlb200_mttkrp_plan_t* pl, bool f16 = false) {
    if (!shape || !pl || ndim < 2 || ndim > TLB200_MAX_NDIM || mode < 0 || mode >= ndim || rank < 1 ||
        !dtype_valid(dtype) || path < TLB200_PATH_AUTO || path > TLB200_PATH_TCGEN05)
        return TLB200_EINVAL;
    for (int i = 0; i < ndim; ++i)
        if (shape[i] < 1) return TLB200_EINVAL;
    const int64_t J = shape[mode];
    auto prod = [&](int first, int count) {
        int64_t p = 1;
        for (int i = first; i < first + count; ++i) p *= shape[i];
        return p;
    };
    int pf = 0, pc = 0, qf = 0, qc = 0;
    if (mode > 0 && mode < ndim - 1) {
        pf = 0; pc = mode; qf = mode + 1; qc = ndim - 1 - mode;
    } else {
        // all contracted modes sit on one side: split them into two contiguous groups so
        // that both tables stay small (never the full Khatri-Rao matrix).
        const int first = mode == 0 ? 1 : 0;
        const int count = ndim - 1;
        if (count == 1) {
            pf = first; pc = 0; qf = first; qc = 1;
        } else {
            int best = 1;
            int64_t best_cost = -1;
            for (int s = 1; s < count; ++s) {
                const int64_t a = prod(first, s), b = prod(first + s, count - s);
                const int64_t cost = a + b + (b < 32 ? (int64_t)1 << 40 : 0);  // keep the inner run >= one K tile
                if (best_cost < 0 || cost < best_cost || (cost == best_cost && b > prod(first + best, count - best))) {
                    best = s; best_cost = cost;
                }
            }
            pf = first; pc = best; qf = first + best; qc = count - best;
        }
    }
    pl->A = prod(pf, pc);
    pl->B = prod(qf, qc);
    pl->J = J;
    if (mode == ndim - 1) {          // j is the contiguous index
        pl->sj = 1; pl->sb = J; pl->sa = pl->B * J;
    } else if (mode == 0) {
        pl->sb = 1; pl->sa = pl->B; pl->sj = pl->A * pl->B;
    } else {
        pl->sb = 1; pl->sj = pl->B; pl->sa = J * pl->B;
    }
    pl->p_first = pf; pl->p_count = pc; pl->q_first = qf; pl->q_count = qc;

    int resolved = TLB200_PATH_SIMT;
    if (path != TLB200_PATH_SIMT && mttkrp_tc_supported(*pl, rank, dtype)) resolved = TLB200_PATH_TCGEN05;
    if (path == TLB200_PATH_TCGEN05 && resolved != TLB200_PATH_TCGEN05) return TLB200_EUNSUPPORTED;
    pl->path = resolved;
    pl->rank_passes = 1;
    pl->f16 = 0;
    pl->reserved_ = 0;

    if (resolved == TLB200_PATH_TCGEN05) {
        pl->f16 = f16 && mttkrp_tc_hf_ok(*pl) ? 1 : 0;
        mttkrp_tc_fill_plan(pl, rank);
    } else {
        const int TR = stream_gemm_tr_for(rank, dtype);
        const int TN = 8 * TR;
        const int KT = dtype == TLB200_F64 ? 16 : 32;
        pl->rank_padded = ceil_div(rank, TN) * TN;
        const int64_t tiles = ceil_div(J, 128) * (pl->rank_padded / TN);
        const int64_t total_chunks = pl->A * ceil_div(pl->B, KT);
        int64_t splits = ceil_div((int64_t)kNumSMs * 4, tiles);
        if (splits > total_chunks / 4) splits = total_chunks / 4;
        if (splits < 1) splits = 1;
        if (splits > 4096) splits = 4096;
        const int64_t per = ceil_div(total_chunks, splits);
        pl->splits = ceil_div(total_chunks, per);
    }
    return TLB200_OK;
}

constexpr int64_t kTcRankChunk = 64;      // widest column block the tcgen05 engine takes in one pass

// Rank > 64 on the tensor-core engine: ceil(rank / 64) passes over the tensor, each producing a block of 64
// columns (factor column slices are pointer offsets).  Two to four passes at ~5 TB/s still beat the SIMT kernel,
// which is FMA-bound at such ranks (0.25 TB/s at rank 100).  The returned plan describes the first pass.
static int make_plan(const int64_t* shape, int ndim, int mode, int64_t rank, int dtype, int path,
                     tlb200_mttkrp_plan_t* pl, bool f16 = false) {
    if (rank > kTcRankChunk && dtype == TLB200_F32 && path != TLB200_PATH_SIMT) {
        tlb200_mttkrp_plan_t chunk;
        const int st = make_plan_one(shape, ndim, mode, kTcRankChunk, dtype, path, &chunk, f16);
        if (st == TLB200_OK && chunk.path == TLB200_PATH_TCGEN05) {
            *pl = chunk;
            pl->rank_passes = (int)ceil_div(rank, kTcRankChunk);
            return TLB200_OK;
        }
    }
    return make_plan_one(shape, ndim, mode, rank, dtype, path, pl, f16);
}

// a registered range hint (tlb200_hint_tensor_absmax) selects the fp16-split engine for this tensor
static bool wants_f16(const void* x, int dtype) { return dtype == TLB200_F32 && x != nullptr && tc_range_hint(x) != nullptr; }

static size_t workspace_for(const tlb200_mttkrp_plan_t& pl, int dtype) {
    const size_t es = dtype_size(dtype);
    size_t total = 0;
    if (pl.p_count > 0) total += align_up((size_t)pl.A * pl.rank_padded * es, 256);
    total += align_up((size_t)2 * (pl.B + 64) * pl.rank_padded * es, 256);   // Q (tcgen05: transposed hi + lo, B padded to 64)
    total += align_up((size_t)pl.splits * pl.J * pl.rank_padded * es, 256);
    if (pl.path == TLB200_PATH_TCGEN05) total += mttkrp_tc_extra_workspace(pl);
    return total + 256;
}

template <typename T>
static int run(const T* x, const int64_t* shape, int ndim, int mode, const T* const* factors,
               const int64_t* frs, const int64_t* fcs, int64_t rank, const T* weights, T* out,
               int64_t out_ld, void* workspace, const tlb200_mttkrp_plan_t& pl, cudaStream_t stream,
               tlb200_partials_t* info = nullptr) {
    Carver ws(workspace);
    const T* P = pl.p_count > 0 ? ws.take<T>((size_t)pl.A * pl.rank_padded) : nullptr;
    T* Q = ws.take<T>((size_t)2 * (pl.B + 64) * pl.rank_padded);
    T* partial = ws.take<T>((size_t)pl.splits * pl.J * pl.rank_padded);

    // The reference's khatri_rao returns a single remaining matrix untouched, i.e. it
    // ignores the weights for 2-way tensors (_khatri_rao.py:68-69).  Mirror that.
    const T* w = ndim == 2 ? nullptr : weights;
    int st;
    if (pl.p_count > 0) {
        // 3-way tensors with unit weights: the outer table IS the factor matrix — no prep launch
        if (const T* direct = table_is_factor<T>(factors, frs, fcs, pl.p_first, pl.p_count, w, rank, pl.rank_padded)) {
            P = direct;
        } else {
            st = launch_khatri_rao<T>(factors + pl.p_first, shape + pl.p_first, frs + pl.p_first, fcs + pl.p_first,
                                      pl.p_count, rank, w, nullptr, const_cast<T*>(P), pl.rank_padded, pl.rank_padded, stream);
            if (st) return st;
            w = nullptr;  // weights go to the first non-skipped factor only
        }
    }
    const float* x_absmax = nullptr;
    if (pl.f16) {
        x_absmax = tc_range_hint(x);
        if (x_absmax == nullptr) return TLB200_EINVAL;       // the hint was withdrawn between planning and launch
    }
    if (x_absmax != nullptr) {
        // Q transposed as fp16 hi / lo tables [rank_padded][Bpad] with one power-of-two scale per column, then the
        // inverse scales [rank_padded] — all inside the region sized for the fp32 tables
        const int64_t bpad = ceil_div(pl.B, 64) * 64;
        __half* qhi = reinterpret_cast<__half*>(Q);
        float* col_inv = reinterpret_cast<float*>(Q) + pl.rank_padded * bpad;
        st = launch_khatri_rao_t_f16(reinterpret_cast<const float* const*>(factors) + pl.q_first, shape + pl.q_first,
                                     frs + pl.q_first, fcs + pl.q_first, pl.q_count, rank,
                                     reinterpret_cast<const float*>(w), qhi, qhi + pl.rank_padded * bpad, bpad,
                                     pl.rank_padded, col_inv, stream);
    } else if (pl.path == TLB200_PATH_TCGEN05) {
        // the tensor-core engine takes Q transposed ([rank_padded][Bpad], K-major rows, zero padded) and already
        // split into tf32 hi / lo tables, which TMA streams straight into the B-operand ring
        const int64_t bpad = ceil_div(pl.B, 64) * 64;
        float* qhi = reinterpret_cast<float*>(Q);
        st = launch_khatri_rao_t<float>(reinterpret_cast<const float* const*>(factors) + pl.q_first, shape + pl.q_first,
                                        frs + pl.q_first, fcs + pl.q_first, pl.q_count, rank,
                                        reinterpret_cast<const float*>(w), qhi, bpad, pl.rank_padded,
                                        qhi + pl.rank_padded * bpad, stream);
    } else {
        st = launch_khatri_rao<T>(factors + pl.q_first, shape + pl.q_first, frs + pl.q_first, fcs + pl.q_first,
                                  pl.q_count, rank, w, nullptr, Q, pl.rank_padded, pl.rank_padded, stream);
    }
    if (st) return st;

    if (pl.path == TLB200_PATH_TCGEN05) {
        set_last_path(x_absmax ? "tcgen05-f16" : "tcgen05");
        st = mttkrp_tc_launch(reinterpret_cast<const float*>(x), pl, rank, reinterpret_cast<const float*>(P),
                              reinterpret_cast<const float*>(Q), reinterpret_cast<float*>(partial),
                              ws.base + ws.used(), stream, x_absmax);
        if (st) return st;
    } else {
        set_last_path("simt");
        const int dtype = sizeof(T) == 8 ? TLB200_F64 : TLB200_F32;
        const int KT = dtype == TLB200_F64 ? 16 : 32;
        StreamGemmParams<T> p;
        p.X = x; p.M = pl.J; p.KA = pl.A; p.KB = pl.B;
        p.sXm = pl.sj; p.sXa = pl.sa; p.sXb = pl.sb; p.sXbatch = 0;
        p.P = P; p.ldP = pl.rank_padded;
        p.Q = Q; p.sQb = pl.rank_padded; p.sQn = 1;
        p.N = pl.rank_padded;
        p.C = partial; p.sCm = pl.rank_padded; p.sCn = 1; p.sCbatch = 0;
        p.sCsplit = pl.J * pl.rank_padded;
        p.nbatch = 1;
        p.chunks_per_a = ceil_div(pl.B, KT);
        p.total_chunks = pl.A * p.chunks_per_a;
        p.chunks_per_split = ceil_div(p.total_chunks, pl.splits);
        p.nsplit = pl.splits;
        st = launch_stream_gemm<T>(p, stream_gemm_tr_for(rank, dtype), pl.sb == 1, stream);
        if (st) return st;
    }
    if (info != nullptr) {          // leave the split-K partials unsummed for tlb200_cp_update_fused
        info->data = partial;
        info->splits = pl.splits;
        info->split_stride = pl.J * pl.rank_padded;
        info->ld = pl.rank_padded;
        info->rows = pl.J;
        info->rank = rank;
        return TLB200_OK;
    }
    const int64_t total = pl.J * rank;
    if constexpr (sizeof(T) == 4) {
        if (rank % 4 == 0 && pl.rank_padded % 4 == 0 && out_ld % 4 == 0 && pl.splits >= 8 && pl.splits < (1 << 30) &&
            reinterpret_cast<uintptr_t>(partial) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
            int64_t blocks = ceil_div(total / 4, 32);
            if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
            splitk_reduce_f4_kernel<0><<<(unsigned)blocks, 256, 0, stream>>>(
                reinterpret_cast<const float4*>(partial), (int)pl.splits, pl.J, (int)(rank / 4), (int)(pl.rank_padded / 4),
                reinterpret_cast<float4*>(out), out_ld / 4);
            TLB_CHECK_LAUNCH();
            return TLB200_OK;
        }
    }
    int64_t blocks = ceil_div(total, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    splitk_reduce_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(partial, pl.splits, pl.J, rank, pl.rank_padded, out, out_ld);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

}  // namespace tlb200

using namespace tlb200;

extern "C" int tlb200_mttkrp_plan(const int64_t* shape, int ndim, int mode, int64_t rank, int dtype, int path,
                                  tlb200_mttkrp_plan_t* plan) {
    return make_plan(shape, ndim, mode, rank, dtype, path, plan);
}

extern "C" size_t tlb200_mttkrp_workspace_bytes(const int64_t* shape, int ndim, int mode, int64_t rank, int dtype,
                                                int path) {
    tlb200_mttkrp_plan_t pl;
    if (make_plan(shape, ndim, mode, rank, dtype, path, &pl)) return 0;
    size_t need = workspace_for(pl, dtype);
    if (pl.path == TLB200_PATH_TCGEN05) {                  // the fp16-split engine blocks the contraction differently
        tlb200_mttkrp_plan_t hf;
        if (!make_plan(shape, ndim, mode, rank, dtype, path, &hf, true)) {
            const size_t h = workspace_for(hf, dtype);
            if (h > need) need = h;
        }
    }
    if (pl.rank_passes > 1 && rank % kTcRankChunk) {       // the last, narrower pass may plan differently
        tlb200_mttkrp_plan_t tail;
        if (!make_plan_one(shape, ndim, mode, rank % kTcRankChunk, dtype, path, &tail)) {
            const size_t t = workspace_for(tail, dtype);
            if (t > need) need = t;
        }
    }
    if (path == TLB200_PATH_AUTO && pl.path == TLB200_PATH_TCGEN05) {
        // AUTO may still have to take the SIMT path at launch time (e.g. a tensor pointer that
        // is not 16-byte aligned cannot be described to TMA): size for both.
        tlb200_mttkrp_plan_t simt;
        if (!make_plan(shape, ndim, mode, rank, dtype, TLB200_PATH_SIMT, &simt)) {
            const size_t s = workspace_for(simt, dtype);
            if (s > need) need = s;
        }
    }
    return need;
}

extern "C" int tlb200_mttkrp(const void* x, const int64_t* shape, int ndim, int mode, const void* const* factors,
                             const int64_t* f_row_stride, const int64_t* f_col_stride, int64_t rank,
                             const void* weights, int dtype, void* out, int64_t out_ld, void* workspace,
                             size_t workspace_bytes, int path, void* stream) {
    tlb200_mttkrp_plan_t pl;
    const bool f16 = wants_f16(x, dtype);
    int st = make_plan(shape, ndim, mode, rank, dtype, path, &pl, f16);
    if (st) return st;
    if (!x || !factors || !f_row_stride || !f_col_stride || !out || out_ld < rank || !workspace) return TLB200_EINVAL;
    for (int i = 0; i < ndim; ++i)
        if (i != mode && !factors[i]) return TLB200_EINVAL;
    if (pl.path == TLB200_PATH_TCGEN05 && reinterpret_cast<uintptr_t>(x) % 16) {
        if (path == TLB200_PATH_TCGEN05) return TLB200_EUNSUPPORTED;
        st = make_plan(shape, ndim, mode, rank, dtype, TLB200_PATH_SIMT, &pl);
        if (st) return st;
    }
    if (workspace_bytes < workspace_for(pl, dtype)) return TLB200_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return TLB200_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (pl.rank_passes > 1) {
        // column blocks of 64: same tensor, factor columns [c0, c0 + rc) (pointer offsets), output columns alike
        const float* fptr[TLB200_MAX_NDIM];
        for (int64_t c0 = 0; c0 < rank; c0 += kTcRankChunk) {
            const int64_t rc = rank - c0 < kTcRankChunk ? rank - c0 : kTcRankChunk;
            tlb200_mttkrp_plan_t cpl;
            st = make_plan_one(shape, ndim, mode, rc, dtype, path, &cpl, f16);
            if (st) return st;
            if (workspace_bytes < workspace_for(cpl, dtype)) return TLB200_EWORKSPACE;
            for (int i = 0; i < ndim; ++i)
                fptr[i] = i == mode ? nullptr : static_cast<const float*>(factors[i]) + c0 * f_col_stride[i];
            st = run<float>((const float*)x, shape, ndim, mode, fptr, f_row_stride, f_col_stride, rc,
                            weights ? (const float*)weights + c0 : nullptr, (float*)out + c0, out_ld, workspace, cpl, s);
            if (st) return st;
        }
        return TLB200_OK;
    }
    if (dtype == TLB200_F32)
        return run<float>((const float*)x, shape, ndim, mode, reinterpret_cast<const float* const*>(factors), f_row_stride,
                          f_col_stride, rank, (const float*)weights, (float*)out, out_ld, workspace, pl, s);
    return run<double>((const double*)x, shape, ndim, mode, reinterpret_cast<const double* const*>(factors), f_row_stride,
                       f_col_stride, rank, (const double*)weights, (double*)out, out_ld, workspace, pl, s);
}

extern "C" int tlb200_mttkrp_partials(const void* x, const int64_t* shape, int ndim, int mode, const void* const* factors,
                                      const int64_t* f_row_stride, const int64_t* f_col_stride, int64_t rank,
                                      const void* weights, int dtype, void* workspace, size_t workspace_bytes, int path,
                                      tlb200_partials_t* partials, void* stream) {
    tlb200_mttkrp_plan_t pl;
    if (!partials) return TLB200_EINVAL;
    int st = make_plan(shape, ndim, mode, rank, dtype, path, &pl, wants_f16(x, dtype));
    if (st) return st;
    if (pl.rank_passes > 1) return TLB200_EUNSUPPORTED;
    if (!x || !factors || !f_row_stride || !f_col_stride || !workspace) return TLB200_EINVAL;
    for (int i = 0; i < ndim; ++i)
        if (i != mode && !factors[i]) return TLB200_EINVAL;
    if (pl.path == TLB200_PATH_TCGEN05 && reinterpret_cast<uintptr_t>(x) % 16) {
        if (path == TLB200_PATH_TCGEN05) return TLB200_EUNSUPPORTED;
        st = make_plan(shape, ndim, mode, rank, dtype, TLB200_PATH_SIMT, &pl);
        if (st) return st;
    }
    if (workspace_bytes < workspace_for(pl, dtype)) return TLB200_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return TLB200_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == TLB200_F32)
        return run<float>((const float*)x, shape, ndim, mode, reinterpret_cast<const float* const*>(factors), f_row_stride,
                          f_col_stride, rank, (const float*)weights, nullptr, 0, workspace, pl, s, partials);
    return run<double>((const double*)x, shape, ndim, mode, reinterpret_cast<const double* const*>(factors), f_row_stride,
                       f_col_stride, rank, (const double*)weights, nullptr, 0, workspace, pl, s, partials);
}
