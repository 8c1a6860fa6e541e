// TTM (mode_dot) on the tcgen05 stream-GEMM engine (tc_stream.cu).
//
//   out[l, i, t] = sum_j M[i, j] * X[l, j, t]         X viewed in place as [L, J, T]
//
//   T >= 32 : rows of the MMA = t (contiguous), contraction = j: per l, tiles of 64 j-rows x 128 t
//             (TC_X_MMAJOR); one work item per (l, t-tile), written straight into out[l, :, t-tile].
//   T == 1  : rows = l, contraction = j contiguous (TC_X_KMAJOR_*): out[l, i].
// The small matrix is split once per call into tf32 hi / lo parts (zero-padded [RP][Kpad], K-major)
// by a tiny prep kernel; the B-producer warp then streams its K slices with TMA.
#include "ttm_tc.cuh"
#include "tc_stream.cuh"
#include "stream_gemm.cuh"
#include "hf_split.cuh"

namespace tlb200 {
namespace {

__global__ void __launch_bounds__(256)
split_matrix_kernel(const float* __restrict__ m, int64_t I, int64_t J, int64_t mrs, int64_t mcs, int RP, int64_t Kpad,
                    float* __restrict__ hi, float* __restrict__ lo) {
    const int64_t total = (int64_t)RP * Kpad;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / Kpad, j = e - i * Kpad;
        const float v = (i < I && j < J) ? m[i * mrs + j * mcs] : 0.f;
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[e] = h;
        lo[e] = v - h;
    }
}

constexpr int64_t kTtmRowBlock = 64;   // output rows (accumulator columns) per pass

struct TtmGeom {
    int layout;
    int64_t M, A, B;       // engine extents: rows, batch, contraction
    int ks;
    int64_t kpad;
    int rp;
    // T == 1 with few row tiles (e.g. the 512 x 512 Gram matrix of an unfolding: 4 tiles): the contraction is split
    // into `ksplit` blocks of `nb` chunks so that the items fill the machine; partials are summed in block order
    int ksplit;
    int nb;
};

bool geom(int64_t L, int64_t J, int64_t T, int64_t I, TtmGeom* g) {
    if (I > 64 || I < 1) return false;
    g->rp = I <= 32 ? 32 : 64;
    if (T >= 32) {
        if (T % 4) return false;                 // TMA strides: multiples of 16 bytes
        g->layout = TC_X_MMAJOR; g->M = T; g->A = L; g->B = J;
    } else if (T == 1) {
        if (J % 4) return false;
        g->layout = (J % 32 == 0) ? TC_X_KMAJOR_2 : TC_X_KMAJOR_1; g->M = L; g->A = 1; g->B = J;
    } else {
        return false;
    }
    if (g->M >= (1LL << 31) || g->A >= (1LL << 31) || g->B >= (1LL << 31)) return false;
    g->ks = tc_chunk_k(g->layout);
    g->kpad = ceil_div(J, g->ks) * g->ks;
    const int64_t chunks = ceil_div(g->B, g->ks);
    g->ksplit = 1;
    g->nb = (int)chunks;
    if (g->layout != TC_X_MMAJOR) {
        const int64_t m_tiles = ceil_div(g->M, 128);
        if (m_tiles * 2 <= kNumSMs && chunks >= 8) {
            int64_t want = kNumSMs / m_tiles;                       // items per row tile
            int64_t nb = ceil_div(chunks, want);
            if (nb < 4) nb = 4;                                     // keep the pipeline fill amortised
            g->nb = (int)nb;
            g->ksplit = (int)ceil_div(chunks, nb);
        }
    }
    return true;
}

}  // namespace

bool ttm_tc_supported(int64_t L, int64_t J, int64_t T, int64_t I) {
    TtmGeom g;
    // one pass over the tensor per 64 output rows.  The SIMT alternative is FMA-bound at such widths (measured:
    // 20 ms for the 512 x 512 Gram matrix of a 512 x 262144 unfolding, i.e. 6.7 TFLOP/s, against ~0.1 ms per pass
    // here), so up to 16 passes (1024 rows) stay on the tensor cores.
    if (I < 1 || I > 16 * kTtmRowBlock) return false;
    if (!geom(L, J, T, I < kTtmRowBlock ? I : kTtmRowBlock, &g)) return false;
    if (L * J * T < (1 << 18) || g.M < 32 || J < 16) return false;   // launch-bound sizes: SIMT
    return tc_available();
}

size_t ttm_tc_workspace(int64_t L, int64_t J, int64_t T, int64_t I) {
    TtmGeom g;
    if (!geom(L, J, T, I < kTtmRowBlock ? I : kTtmRowBlock, &g)) return 0;
    size_t total = 2 * align_up((size_t)g.rp * g.kpad * 4, 256) + 256 + 256;      // + the fp16 engine's column scales
    if (g.ksplit > 1) total += align_up((size_t)g.ksplit * g.M * g.rp * 4, 256);
    return total;
}

// one pass: rows [0, I) of `m` (I <= 64) into output rows of an array whose mode extent is I_total
static int ttm_tc_launch_block(const float* x, int64_t L, int64_t J, int64_t T, const float* m, int64_t I, int64_t I_total,
                               int64_t mrs, int64_t mcs, float* out, void* workspace, cudaStream_t stream) {
    TtmGeom g;
    if (!geom(L, J, T, I, &g) || !workspace) return TLB200_EUNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(x) % 16) return TLB200_EUNSUPPORTED;
    Carver ws(workspace);
    float* bhi = ws.take<float>((size_t)g.rp * g.kpad);
    float* blo = ws.take<float>((size_t)g.rp * g.kpad);
    float* partial = g.ksplit > 1 ? ws.take<float>((size_t)g.ksplit * g.M * g.rp) : nullptr;
    float* col_inv = ws.take<float>(64);
    // a registered range hint for this tensor selects the fp16-split engine (64-element tiles only)
    const float* x_absmax = g.ks == 64 ? tc_range_hint(x) : nullptr;
    set_last_path(x_absmax ? "tcgen05-f16" : "tcgen05");
    if (x_absmax != nullptr) {
        const int st0 = launch_split_matrix_f16(m, I, J, mrs, mcs, g.rp, g.kpad, reinterpret_cast<__half*>(bhi),
                                                reinterpret_cast<__half*>(blo), col_inv, stream);
        if (st0) return st0;
    } else {
        const int64_t total = (int64_t)g.rp * g.kpad;
        int64_t blocks = ceil_div(total, 256);
        if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
        split_matrix_kernel<<<(unsigned)blocks, 256, 0, stream>>>(m, I, J, mrs, mcs, g.rp, g.kpad, bhi, blo);
        TLB_CHECK_LAUNCH();
    }
    TcStreamLaunch l;
    l.hf = x_absmax != nullptr;
    l.rp = g.rp;
    l.x_layout = g.layout;
    l.b_mode = TC_B_MAT;
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    int st;
    if (g.layout == TC_X_MMAJOR) {           // X[l][j][t]: dims (T, J, L)
        dims[0] = T; dims[1] = J; dims[2] = L;
        strides[0] = (uint64_t)T * 4; strides[1] = (uint64_t)J * T * 4;
        box[0] = 128; box[1] = 64; box[2] = 1;
        st = tc_encode_map(&l.x_map, x, 3, dims, strides, box, false);
    } else if (g.layout == TC_X_KMAJOR_2) {  // X[l][j]: dims (32, J/32, L, 1)
        dims[0] = 32; dims[1] = J / 32; dims[2] = L; dims[3] = 1;
        strides[0] = 128; strides[1] = (uint64_t)J * 4; strides[2] = (uint64_t)J * 4 * L;
        box[0] = 32; box[1] = 2; box[2] = 128; box[3] = 1;
        st = tc_encode_map(&l.x_map, x, 4, dims, strides, box, true);
    } else {                                 // X[l][j]: dims (J, L, 1)
        dims[0] = J; dims[1] = L; dims[2] = 1;
        strides[0] = (uint64_t)J * 4; strides[1] = (uint64_t)J * 4 * L;
        box[0] = 32; box[1] = 128; box[2] = 1;
        st = tc_encode_map(&l.x_map, x, 3, dims, strides, box, true);
    }
    if (st) return st;
    if (l.hf) {
        uint64_t bd[2] = {(uint64_t)g.kpad, (uint64_t)g.rp}, bs[1] = {(uint64_t)g.kpad * 2};
        uint32_t bb[2] = {64, (uint32_t)g.rp};
        st = tc_encode_map(&l.bhi_map, bhi, 2, bd, bs, bb, true, true);
        if (st) return st;
        st = tc_encode_map(&l.blo_map, blo, 2, bd, bs, bb, true, true);
        if (st) return st;
    } else {
        uint64_t bd[2] = {(uint64_t)g.kpad, (uint64_t)g.rp}, bs[1] = {(uint64_t)g.kpad * 4};
        uint32_t bb[2] = {32, (uint32_t)g.rp};
        st = tc_encode_map(&l.bhi_map, bhi, 2, bd, bs, bb, true);
        if (st) return st;
        st = tc_encode_map(&l.blo_map, blo, 2, bd, bs, bb, true);
        if (st) return st;
    }
    TcStreamParams& p = l.p;
    p.M = g.M; p.A = g.A; p.B = g.B;
    p.chunks_per_a = ceil_div(g.B, g.ks);
    p.m_tiles = (int)ceil_div(g.M, 128);
    p.k_ranges = g.A;                         // one item per (row tile, batch): all of K
    p.a_per_range = 1;
    p.nb = g.nb; p.n_bblocks = g.ksplit;
    p.b_resident = 0;                         // the matrix is streamed with the tiles (re-read from L2 per item)
    p.group_units = tc_group_units(l.hf != 0);
    p.P = nullptr;
    p.x_absmax = x_absmax;
    p.col_inv = l.hf ? col_inv : nullptr;
    p.out = out;
    if (g.layout == TC_X_MMAJOR) { p.sOk = I_total * T; p.sOm = 1; p.sOn = T; }
    else                         { p.sOk = 0; p.sOm = I_total; p.sOn = 1; }
    p.n_valid = (int)I;
    if (partial) {                            // split contraction: [ksplit][M][rp] partials, then an ordered sum
        p.out = partial;
        p.sOk = g.M * g.rp; p.sOm = g.rp; p.sOn = 1;
        p.n_valid = g.rp;
    }
    st = tc_stream_launch(l, stream);
    if (st || !partial) return st;
    const int64_t total = g.M * I;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    splitk_reduce_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(partial, g.ksplit, g.M, I, g.rp, out, I_total);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

// More than 64 output rows: one pass over the tensor per block of 64 rows of the matrix (row blocks of `m` and of
// the output are pointer offsets).  2-4 passes at several TB/s still beat the FMA-bound SIMT kernel.
int ttm_tc_launch(const float* x, int64_t L, int64_t J, int64_t T, const float* m, int64_t I, int64_t mrs, int64_t mcs,
                  float* out, void* workspace, cudaStream_t stream) {
    for (int64_t c0 = 0; c0 < I; c0 += kTtmRowBlock) {
        const int64_t rc = I - c0 < kTtmRowBlock ? I - c0 : kTtmRowBlock;
        float* dst = out + (T >= 32 ? c0 * T : c0);
        const int st = ttm_tc_launch_block(x, L, J, T, m + c0 * mrs, rc, I, mrs, mcs, dst, workspace, stream);
        if (st) return st;
    }
    return TLB200_OK;
}

}  // namespace tlb200
