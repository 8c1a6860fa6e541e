// MTTKRP on the tcgen05 stream-GEMM engine (tc_stream.cu): maps the (A, J, B) plan of
// mttkrp.cu onto work items and TMA tensor maps.
//
//   kept mode not last : X[a][j][b], b contiguous  -> rows = j, K-major tiles.  When B is a
//                        multiple of 32 each TMA box fetches two adjacent 128-byte lines per row
//                        (4-D map {32, B/32, J, A}), the DRAM-friendly shape; otherwise one line.
//   kept mode last     : X[a][b][j], j contiguous  -> rows = j, tiles of 64 b-rows x 128 j.
// Split-K: item (j-tile, split) covers a contiguous range of K chunks; partials are summed by
// the deterministic reduction kernel of mttkrp.cu.
#include "mttkrp_tc.cuh"
#include "tc_stream.cuh"

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

void mttkrp_tc_fill_plan(tlb200_mttkrp_plan_t* pl, int64_t rank) {
    pl->rank_padded = rank <= 32 ? 32 : 64;
    const int ks = tc_chunk_k(layout_for(*pl));
    const int64_t m_tiles = ceil_div(pl->J, 128);
    const int64_t total = pl->A * ceil_div(pl->B, ks);
    // Persistent CTAs (one per SM) take work items (row tile, K range) round-robin.  Pick the split-K factor
    // whose item count fills whole rounds of 148 best; among equally good ones the smallest (longer items,
    // fewer partials), but at least 8 tiles per item and at most 8 rounds.
    int64_t best = 1;
    double best_eff = -1.0;
    const int64_t s_max = total / 8 > 0 ? total / 8 : 1;
    for (int64_t s = 1; s <= s_max && m_tiles * s <= (int64_t)kNumSMs * 8; ++s) {
        const int64_t per = ceil_div(total, s);
        const int64_t items = m_tiles * ceil_div(total, per);
        const double eff = (double)items / (double)(ceil_div(items, kNumSMs) * kNumSMs);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    if (const char* e = getenv("TLB200_TC_SPLITS")) { const int64_t v = atoll(e); if (v >= 1 && v <= total) best = v; }
    const int64_t per = ceil_div(total, best);
    pl->splits = ceil_div(total, per);
}

size_t mttkrp_tc_extra_workspace(const tlb200_mttkrp_plan_t&) { return 0; }

int mttkrp_tc_launch(const float* x, const tlb200_mttkrp_plan_t& pl, int64_t /*rank*/, const float* P, const float* Q,
                     float* partial, void* /*extra_ws*/, cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(x) % 16) return TLB200_EUNSUPPORTED;
    TcStreamLaunch l;
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
    {   // Q^T hi / lo tables [rank_padded][ldq]: one unit = box of 32 contraction elements x all rows, swizzled
        const uint64_t ldq = (uint64_t)ceil_div(pl.B, 64) * 64;
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
    p.total_chunks = pl.A * p.chunks_per_a;
    p.m_tiles = (int)ceil_div(pl.J, 128);
    p.k_ranges = pl.splits;
    p.chunks_per_range = ceil_div(p.total_chunks, pl.splits);
    p.group_units = tc_group_units();
    p.P = P;     // outer Khatri-Rao table: applied per `a` by the epilogue
    p.out = partial;
    p.sOk = pl.J * pl.rank_padded; p.sOm = pl.rank_padded; p.sOn = 1;
    p.n_valid = (int)pl.rank_padded;
    return tc_stream_launch(l, stream);
}

}  // namespace tlb200
