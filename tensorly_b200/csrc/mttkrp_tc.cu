// MTTKRP on the tcgen05 stream-GEMM engine (tc_stream.cu): maps the (A, J, B) plan of
// mttkrp.cu onto work items and TMA tensor maps.
//
//   kept mode not last : X[a][j][b], b contiguous  -> rows = j, K-major tiles.  When B is a
//                        multiple of 32 each TMA box fetches two adjacent 128-byte lines per row
//                        (4-D map {32, B/32, J, A}), the DRAM-friendly shape; otherwise one line.
//   kept mode last     : X[a][b][j], j contiguous  -> rows = j, tiles of 64 b-rows x 128 j.
// Split-K: item (j-tile, b block, a range): the b block's rows of the inner Khatri-Rao table stay resident in
// shared memory while the item streams X over its `a` range (no per-tile reload through L2); every item
// writes one partial, summed by the deterministic reduction kernel of mttkrp.cu.
#include "mttkrp_tc.cuh"
#include "tc_stream.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

namespace tlb200 {

static int layout_for(const tlb200_mttkrp_plan_t& pl) {
    if (pl.sb != 1) return TC_X_MMAJOR;
    return (pl.B % 32 == 0) ? TC_X_KMAJOR_2 : TC_X_KMAJOR_1;
}

bool mttkrp_tc_supported(const tlb200_mttkrp_plan_t& pl, int64_t rank, int dtype) {
    if (dtype != TLB200_F32 || rank > 64) return false;
    // TMA: global strides must be multiples of 16 bytes, extents must fit 32 bits
    if (pl.sb == 1) { if ((pl.sj % 4) || (pl.sa % 4)) return false; }
    else            { if ((pl.sb % 4) || (pl.sa % 4)) return false; }
    if (pl.A >= (1LL << 31) || pl.B >= (1LL << 31) || pl.J >= (1LL << 31)) return false;
    // tiny problems are launch-bound: the SIMT kernel has the shorter prologue
    if (pl.A * ceil_div(pl.B, 64) < 8 || pl.J < 32) return false;
    return tc_available();
}

bool mttkrp_tc_hf_ok(const tlb200_mttkrp_plan_t& pl) { return layout_for(pl) != TC_X_KMAJOR_1; }

// chunks of the inner Khatri-Rao table one work item keeps resident in shared memory
static int block_chunks(const tlb200_mttkrp_plan_t& pl, int64_t rank_padded) {
    const int ks = tc_chunk_k(layout_for(pl));
    if (pl.f16) {
        // fp16 engine: every slot holds a whole 64-element tile of the table, so twice the block fits — and the
        // accumulation groups (cut at every `a` boundary) are twice as long: half the drains per tile
        const int cap = tc_b_slots((int)rank_padded);
        if (const char* e = getenv("TLB200_TC_NB_F16")) { const int v = atoi(e); if (v >= 1 && v <= cap) return v; }
        // equal blocks: the largest size in [cap / 2, cap] that wastes the fewest chunk slots over the b range
        // (2048 / 64 = 32 chunks: 8 x 4 rather than 5 x 6 + 2)
        const int64_t cpa = ceil_div(pl.B, ks);
        if (cpa <= cap) return (int)cpa;
        int best = cap;
        int64_t best_waste = -1;
        for (int nb = cap; nb >= (cap + 1) / 2; --nb) {
            const int64_t waste = ceil_div(cpa, nb) * nb - cpa;
            if (best_waste < 0 || waste < best_waste) { best = nb; best_waste = waste; }
        }
        return best;
    }
    // as many chunks as fit in shared memory: a shorter block would mean shorter accumulation groups, i.e. more
    // epilogue drains per tile (measured: equal blocks of 2 chunks lose to blocks of 3 + 1 at B = 256, rank 64)
    return tc_b_slots((int)rank_padded) * 32 / ks;
}

void mttkrp_tc_fill_plan(tlb200_mttkrp_plan_t* pl, int64_t rank) {
    pl->rank_padded = rank <= 32 ? 32 : 64;
    const int ks = tc_chunk_k(layout_for(*pl));
    const int nb = block_chunks(*pl, pl->rank_padded);
    const int64_t m_tiles = ceil_div(pl->J, 128);
    const int64_t cpa = ceil_div(pl->B, ks);
    const int64_t n_bblocks = ceil_div(cpa, nb);
    // Persistent CTAs (one per SM) take work items (row tile, b block, a range) round-robin.  Every item
    // writes one partial result (one X tile's worth of traffic at rank 64) and pays a pipeline fill of a few
    // tiles, so score = how well the items fill whole rounds of 148 CTAs x the useful share of an item.
    int64_t best = 1;
    double best_score = -1.0;
    const double per_item = 2.5 + 2.0 * (double)pl->rank_padded / 64.0;
    for (int64_t s = 1; s <= pl->A; ++s) {
        const int64_t apr = ceil_div(pl->A, s);
        const int64_t k = ceil_div(pl->A, apr);
        if (k != s) continue;                                  // same partition as a smaller s
        const int64_t items = m_tiles * n_bblocks * k;
        if (items > (int64_t)kNumSMs * 8 && s > 1) break;
        const double tiles = (double)apr * (double)(cpa < nb ? cpa : nb);
        const double eff = (double)items / (double)(ceil_div(items, kNumSMs) * kNumSMs);
        const double score = eff * tiles / (tiles + per_item);
        if (score > best_score + 1e-9) { best_score = score; best = k; }
    }
    if (const char* e = getenv("TLB200_TC_SPLITS")) {
        const int64_t v = atoll(e);
        if (v >= 1 && v <= pl->A) best = ceil_div(pl->A, ceil_div(pl->A, v));
    }
    pl->splits = best * n_bblocks;         // number of partial results
}

size_t mttkrp_tc_extra_workspace(const tlb200_mttkrp_plan_t&) { return 0; }

int mttkrp_tc_launch(const float* x, const tlb200_mttkrp_plan_t& pl, int64_t /*rank*/, const float* P, const float* Q,
                     float* partial, void* /*extra_ws*/, cudaStream_t stream, const float* x_absmax) {
    if (reinterpret_cast<uintptr_t>(x) % 16) return TLB200_EUNSUPPORTED;
    TcStreamLaunch l;
    l.hf = x_absmax != nullptr;
    l.rp = (int)pl.rank_padded;
    l.x_layout = layout_for(pl);
    l.b_mode = TC_B_MAT;
    const int ks = tc_chunk_k(l.x_layout);
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    int st;
    if (l.x_layout == TC_X_KMAJOR_1) {
        dims[0] = pl.B; dims[1] = pl.J; dims[2] = pl.A;
        strides[0] = (uint64_t)pl.sj * 4; strides[1] = (uint64_t)pl.sa * 4;
        box[0] = 32; box[1] = 128; box[2] = 1;
        st = tc_encode_map(&l.x_map, x, 3, dims, strides, box, true);
    } else if (l.x_layout == TC_X_KMAJOR_2) {
        dims[0] = 32; dims[1] = pl.B / 32; dims[2] = pl.J; dims[3] = pl.A;
        strides[0] = 128; strides[1] = (uint64_t)pl.sj * 4; strides[2] = (uint64_t)pl.sa * 4;
        box[0] = 32; box[1] = 2; box[2] = 128; box[3] = 1;
        st = tc_encode_map(&l.x_map, x, 4, dims, strides, box, true);
    } else {
        dims[0] = pl.J; dims[1] = pl.B; dims[2] = pl.A;
        strides[0] = (uint64_t)pl.sb * 4; strides[1] = (uint64_t)pl.sa * 4;
        box[0] = 128; box[1] = 64; box[2] = 1;
        st = tc_encode_map(&l.x_map, x, 3, dims, strides, box, false);
    }
    if (st) return st;
    const uint64_t ldq = (uint64_t)ceil_div(pl.B, 64) * 64;
    if (l.hf) {   // fp16 tables: one slot = box of 64 contraction elements (128 bytes) x all rows
        uint64_t qd[2] = {ldq, (uint64_t)pl.rank_padded}, qs[1] = {ldq * 2};
        uint32_t qb[2] = {64, (uint32_t)pl.rank_padded};
        const __half* qh = reinterpret_cast<const __half*>(Q);
        st = tc_encode_map(&l.bhi_map, qh, 2, qd, qs, qb, true, true);
        if (st) return st;
        st = tc_encode_map(&l.blo_map, qh + pl.rank_padded * ldq, 2, qd, qs, qb, true, true);
        if (st) return st;
    } else {   // Q^T hi / lo tables [rank_padded][ldq]: one unit = box of 32 contraction elements x all rows, swizzled
        uint64_t qd[2] = {ldq, (uint64_t)pl.rank_padded}, qs[1] = {ldq * 4};
        uint32_t qb[2] = {32, (uint32_t)pl.rank_padded};
        st = tc_encode_map(&l.bhi_map, Q, 2, qd, qs, qb, true);
        if (st) return st;
        st = tc_encode_map(&l.blo_map, Q + pl.rank_padded * ldq, 2, qd, qs, qb, true);
        if (st) return st;
    }

    TcStreamParams& p = l.p;
    p.M = pl.J; p.A = pl.A; p.B = pl.B;
    p.chunks_per_a = ceil_div(pl.B, ks);
    p.m_tiles = (int)ceil_div(pl.J, 128);
    p.nb = block_chunks(pl, pl.rank_padded);
    p.n_bblocks = (int)ceil_div(p.chunks_per_a, p.nb);
    p.k_ranges = pl.splits / p.n_bblocks;
    if (p.k_ranges < 1 || p.k_ranges * p.n_bblocks != pl.splits) return TLB200_EINVAL;
    p.a_per_range = ceil_div(pl.A, p.k_ranges);
    p.b_resident = 1;
    p.group_units = tc_group_units(l.hf != 0);
    p.x_absmax = x_absmax;
    p.col_inv = l.hf ? Q + pl.rank_padded * ldq : nullptr;
    p.P = P;     // outer Khatri-Rao table: applied per `a` by the epilogue
    p.out = partial;
    p.sOk = pl.J * pl.rank_padded; p.sOm = pl.rank_padded; p.sOn = 1;
    p.n_valid = (int)pl.rank_padded;
    return tc_stream_launch(l, stream);
}

}  // namespace tlb200
