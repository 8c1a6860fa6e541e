// tcgen05 "stream GEMM" — the tensor-core engine behind the fp32 MTTKRP and TTM paths.
//
// Why tensor cores for a "streaming" op: MTTKRP/TTM need 2R flops per tensor element, i.e.
// 16-32 flop/byte at R = 32-64.  Streaming the tensor at HBM speed therefore needs
// 100-200 TFLOP/s of fp32-accurate math — more than the CUDA cores have — so the
// contraction runs as a skinny GEMM  D[128 x R] += X_tile[128 x KS] * B_tile[KS x R]
// on tcgen05 with the accumulator in tensor memory.
//
// Why 3xTF32: kind::tf32 truncates its 32-bit operands to 10 mantissa bits (measured:
// probes/tc_probe.cu) — ~7e-4 relative error, far above the 1e-5 gate.  Each operand is split
// exactly into hi = trunc_tf32(x) and lo = x - hi and  x*y ~= hi*hi + lo*hi + hi*lo  (dropped
// term ~2^-22): three MMAs per K step.
//
// Why short accumulation groups: the tensor core accumulates with round-toward-zero
// (measured), a systematic ~2^-24 relative loss per MMA.  The TMEM accumulators are therefore
// drained every `group_units` 32-element K units into fp32 registers (round-to-nearest adds)
// by dedicated epilogue warps while the MMAs continue into the second accumulator set.
//
// Why two MMAs of different width per K step: a chain of small dependent MMAs (N = 32) runs
// at the tensor pipe's latency (~100 clk each, measured), not its throughput.  The three
// products are therefore issued as  D[128 x 2R] += A_hi * [B_hi | B_lo]  followed by
// D[:, R:2R] += A_lo * B_hi : two instructions per K step, the wide one covering two
// products (measured cost per MMA: max(48, N/2) clk for M=128, K=8 — probes/mma_rate.cu).
// The hi*hi columns carry the only significant truncation error; the two small cross terms
// share the other R columns.
//
// Why the A operand goes through TMEM: with A read from shared memory the tile would cross
// the 128 B/clk shared-memory port six times (TMA write, split read+write, three MMA reads);
// converted in registers and stored to TMEM (tcgen05.st) it crosses it twice.  The register
// hop also performs the transposition for the m-contiguous layout and lets the X tile use
// the DRAM-friendliest shape: two adjacent 128-byte lines per row per TMA box (measured
// 6.8 TB/s vs 4.5 TB/s for one line per row on rows that are megabytes apart).
//
// Warp roles (14 warps, 1 CTA per SM, persistent over work items):
//   warp 0       TMA producer of X tiles (ring of XS stages)
//   warp 1       MMA issuer (single thread)
//   warps 2-5    convert: smem X tile -> registers -> hi/lo -> TMEM A ring (AS stages)
//   warps 6-9    B producer: Khatri-Rao rows P[a,:]*Q[b,:] split hi/lo into a K-major
//                SWIZZLE_128B smem ring (or, for TTM, TMA loads of the pre-split matrix)
//   warps 10-13  epilogue: drain TMEM accumulation groups, write C
#include "tc_stream.cuh"

#include <cstdlib>

namespace tlb200 {
namespace {

constexpr int TM = 128;
constexpr int NUM_THREADS = 448;
constexpr uint32_t SPIN_LIMIT = 1u << 22;   // x ~1 us per try_wait: seconds, then trap (never hang the GPU)

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    const uint32_t addr = smem_u32(bar);
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (++spins > SPIN_LIMIT) asm volatile("trap;");
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row atoms of 1 KB)
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version for sm_100
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TLB_TMEM_LD32(taddr, r)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),         \
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),        \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                      \
                 : "r"(taddr))
#define TLB_TMEM_ST32(taddr, r)                                                                                          \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                        \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),  \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),       \
                 "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),      \
                 "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                       \
                 : "memory")

// ring position: stage index + phase parity, advanced without divisions
struct Ring {
    int idx = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1u; } }
};

// ---- compile-time configuration per (RP, X layout) -------------------------------------
template <int RP, int XL>
struct Cfg {
    static constexpr int KS = XL == TC_X_KMAJOR_1 ? 32 : 64;        // K extent of one X stage
    static constexpr int KO = KS / 32;                              // 32-element units per X stage
    static constexpr int X_STAGE = TM * KS * 4;
    static constexpr int XS = KS == 32 ? 6 : 4;
    static constexpr int D_COLS = 2 * RP;                           // per accumulator set: [hi*hi (RP) | hi*lo + lo*hi (RP)]
    static constexpr int A_COLS = 64;                               // TMEM columns per A unit: [hi 32 | lo 32]
    static constexpr int AS = 4;
    static constexpr int B_UNIT = 2 * RP * 128;                     // [hi RP rows | lo RP rows] x 128 B, K-major SW128
    static constexpr int BS = RP == 64 ? 4 : 8;                     // power of two (slot = unit % BS)
    static constexpr int STAGE_F = 32 * RP + RP;                    // per KR warp: rotated Q staging [32][RP] + P row [RP]
    static constexpr int OFF_X = 0;
    static constexpr int OFF_B = OFF_X + XS * X_STAGE;
    static constexpr int OFF_STAGE = OFF_B + BS * B_UNIT;
    static constexpr int OFF_BAR = OFF_STAGE + 4 * STAGE_F * 4;
    static constexpr int NUM_BARS = 2 * XS + 2 * AS + 2 * BS + 4;
    static constexpr int SMEM = OFF_BAR + NUM_BARS * 8 + 16;
    static constexpr int TMEM_COLS = 2 * D_COLS + AS * A_COLS;
    static_assert(AS >= 2, "need at least two A units");
    static_assert(TMEM_COLS <= 512, "TMEM budget");
    static_assert(SMEM + 1024 <= 227 * 1024, "smem budget");
};

template <int RP, int XL, int BM>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_stream_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap bhi_map,
                 const __grid_constant__ CUtensorMap blo_map, const TcStreamParams p) {
    using C = Cfg<RP, XL>;
    constexpr int KS = C::KS, KO = C::KO, XS = C::XS, AS = C::AS, BS = C::BS;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* x_smem = smem + C::OFF_X;
    unsigned char* b_smem = smem + C::OFF_B;
    float* stage = reinterpret_cast<float*>(smem + C::OFF_STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + XS;
    uint64_t* a_full = x_empty + XS;
    uint64_t* a_empty = a_full + AS;
    uint64_t* b_full = a_empty + AS;
    uint64_t* b_empty = b_full + BS;
    uint64_t* d_full = b_empty + BS;
    uint64_t* d_empty = d_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_items = (int64_t)p.m_tiles * p.k_ranges;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 128); }
        for (int i = 0; i < AS; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < BS; ++i) { mbar_init(&b_full[i], BM == TC_B_KR ? 32 : 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
            if (BM == TC_B_MAT) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&bhi_map)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&blo_map)) : "memory");
            }
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: accumulator set 0 [0, 2RP), set 1 [2RP, 4RP), then AS A-operand units of [hi 32 | lo 32]
    const uint32_t a_col0 = 2 * C::D_COLS;
    const int GU = p.group_units;

    if (warp == 0) {
        // ================= TMA producer of X tiles =================
        if (lane == 0) {
            Ring xr;
            for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int mt = (int)(it % p.m_tiles);
                const int64_t kr = it / p.m_tiles;
                const int64_t c_begin = kr * p.chunks_per_range;
                const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_range);
                int a = (int)(c_begin / p.chunks_per_a);
                int bc = (int)(c_begin - (int64_t)a * p.chunks_per_a);
                const int m0 = mt * TM;
                for (int64_t c = c_begin; c < c_end; ++c) {
                    mbar_wait(&x_empty[xr.idx], xr.phase ^ 1u);
                    mbar_expect_tx(&x_full[xr.idx], C::X_STAGE);
                    unsigned char* dst = x_smem + xr.idx * C::X_STAGE;
                    if constexpr (XL == TC_X_KMAJOR_1)      tma_load_3d(dst, &xmap, &x_full[xr.idx], bc * 32, m0, a);
                    else if constexpr (XL == TC_X_KMAJOR_2) tma_load_4d(dst, &xmap, &x_full[xr.idx], 0, bc * 2, m0, a);
                    else                                    tma_load_3d(dst, &xmap, &x_full[xr.idx], m0, bc * 64, a);
                    xr.advance(XS);
                    if (++bc == (int)p.chunks_per_a) { bc = 0; ++a; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc1 = idesc_tf32(TM, 2 * RP);   // A_hi x [B_hi | B_lo] -> columns [0, 2RP)
            constexpr uint32_t idesc2 = idesc_tf32(TM, RP);       // A_lo x B_hi          -> columns [RP, 2RP)
            Ring ar, br;
            uint32_t G = 0;          // global accumulation-group counter
            for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int64_t kr = it / p.m_tiles;
                const int64_t c_begin = kr * p.chunks_per_range;
                const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_range);
                const int n = (int)(c_end - c_begin) * KO;       // 32-element units of this item
                int ug = 0;
                for (int i = 0; i < n; ++i) {
                    const uint32_t buf = G & 1u;
                    if (ug == 0 && G >= 2) mbar_wait(&d_empty[buf], ((G >> 1) - 1) & 1u);
                    mbar_wait(&a_full[ar.idx], ar.phase);
                    mbar_wait(&b_full[br.idx], br.phase);
                    tc_fence_after();
                    const uint32_t d1 = tmem_base + buf * C::D_COLS;
                    const uint32_t d2 = d1 + RP;
                    const uint32_t a_hi = tmem_base + a_col0 + ar.idx * C::A_COLS;
                    const uint32_t a_lo = a_hi + 32;
                    const uint32_t bbase = smem_u32(b_smem + br.idx * C::B_UNIT);
                    if (!(p.debug & 2))
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t db = desc_kmajor_sw128(bbase + ks * 32);
                        const uint32_t accf = (ug == 0 && ks == 0) ? 0u : 1u;
                        mma_ts_tf32(d1, a_hi + ks * 8, db, idesc1, accf);
                        mma_ts_tf32(d2, a_lo + ks * 8, db, idesc2, 1u);
                    }
                    tc_commit(&a_empty[ar.idx]);
                    tc_commit(&b_empty[br.idx]);
                    ar.advance(AS);
                    br.advance(BS);
                    if (++ug == GU || i == n - 1) { tc_commit(&d_full[buf]); ++G; ug = 0; }
                }
            }
        }
    } else if (warp < 6) {
        // ================= convert: smem X tile -> hi/lo -> TMEM A ring =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        Ring xr, ar, ar_pub;       // ar: next A unit to fill; ar_pub: next A unit to publish (a_full)
        int unpublished = 0;       // A units whose tcgen05.st have been issued but not yet waited for
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int64_t kr = it / p.m_tiles;
            const int64_t c_begin = kr * p.chunks_per_range;
            const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_range);
            const int n = (int)(c_end - c_begin);
            for (int i = 0; i < n; ++i) {
                mbar_wait(&x_full[xr.idx], xr.phase);
                const unsigned char* xt = x_smem + xr.idx * C::X_STAGE;
                // ta = first 32 contraction elements of this thread's row, tb = the next 32 (KS == 64 only)
                uint32_t ta[32];
                uint32_t tb[KS == 64 ? 32 : 1];
                int flip = 0;
                if constexpr (XL == TC_X_KMAJOR_1) {
                    const unsigned char* xrow = xt + row * 128;                 // [128 rows][128 B], swizzled by row
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 w = *reinterpret_cast<const uint4*>(xrow + ((c ^ (row & 7)) << 4));
                        ta[4 * c + 0] = w.x; ta[4 * c + 1] = w.y; ta[4 * c + 2] = w.z; ta[4 * c + 3] = w.w;
                    }
                } else if constexpr (XL == TC_X_KMAJOR_2) {
                    // [128 rows][2 lines][128 B]; line L = 2*row + ko, 16-byte chunk c stored at c ^ (L & 7).
                    // Lanes 4-7 of every 8 read the two lines in the opposite order so that a quarter-warp
                    // touches 8 distinct swizzle phases (conflict-free); the selects below undo the swap.
                    flip = (lane >> 2) & 1;
                    {
                        const int L = 2 * row + flip;
                        const unsigned char* xl = xt + L * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 w = *reinterpret_cast<const uint4*>(xl + ((c ^ (L & 7)) << 4));
                            ta[4 * c + 0] = w.x; ta[4 * c + 1] = w.y; ta[4 * c + 2] = w.z; ta[4 * c + 3] = w.w;
                        }
                    }
                    {
                        const int L = 2 * row + (flip ^ 1);
                        const unsigned char* xl = xt + L * 128;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 w = *reinterpret_cast<const uint4*>(xl + ((c ^ (L & 7)) << 4));
                            tb[4 * c + 0] = w.x; tb[4 * c + 1] = w.y; tb[4 * c + 2] = w.z; tb[4 * c + 3] = w.w;
                        }
                    }
                } else {
                    const float* xc = reinterpret_cast<const float*>(xt) + row;      // [64 k][128 m]
#pragma unroll
                    for (int k = 0; k < 32; ++k) ta[k] = __float_as_uint(xc[k * TM]);
#pragma unroll
                    for (int k = 0; k < 32; ++k) tb[k] = __float_as_uint(xc[(32 + k) * TM]);
                }
                // While those shared-memory loads are in flight, publish the A units stored in the
                // previous iteration (their tcgen05.st have had a whole iteration to complete).
                if (unpublished) {
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    for (; unpublished > 0; --unpublished) { mbar_arrive(&a_full[ar_pub.idx]); ar_pub.advance(AS); }
                }
                uint32_t h[32];
#pragma unroll
                for (int u = 0; u < KO; ++u) {
                    mbar_wait(&a_empty[ar.idx], ar.phase ^ 1u);
                    tc_fence_after();
                    const uint32_t abase = lane_addr + a_col0 + ar.idx * C::A_COLS;
                    if (!(p.debug & 4)) {
                        if (KS == 32 || u == 0) {
#pragma unroll
                            for (int k = 0; k < 32; ++k) h[k] = ((KS == 64 && flip) ? tb[k & (KS == 64 ? 31 : 0)] : ta[k]) & 0xFFFFE000u;
                            TLB_TMEM_ST32(abase, h);                                         // exact tf32 part
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                h[k] = __float_as_uint(__uint_as_float((KS == 64 && flip) ? tb[k & (KS == 64 ? 31 : 0)] : ta[k]) -
                                                       __uint_as_float(h[k]));
                            TLB_TMEM_ST32(abase + 32, h);                                    // exact remainder
                        } else {
#pragma unroll
                            for (int k = 0; k < 32; ++k) h[k] = (flip ? ta[k] : tb[k & (KS == 64 ? 31 : 0)]) & 0xFFFFE000u;
                            TLB_TMEM_ST32(abase, h);
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                h[k] = __float_as_uint(__uint_as_float(flip ? ta[k] : tb[k & (KS == 64 ? 31 : 0)]) - __uint_as_float(h[k]));
                            TLB_TMEM_ST32(abase + 32, h);
                        }
                    }
                    ar.advance(AS);
                    ++unpublished;
                }
                // every loaded value has been consumed: only now hand the X stage back to the TMA producer
                // (generic-proxy reads must be ordered before the async-proxy overwrite)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&x_empty[xr.idx]);
                xr.advance(XS);
            }
        }
        if (unpublished) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            for (; unpublished > 0; --unpublished) { mbar_arrive(&a_full[ar_pub.idx]); ar_pub.advance(AS); }
        }
    } else if (warp < 10) {
        // ================= B producer (one 32-element K unit at a time) =================
        if constexpr (BM == TC_B_MAT) {
            if (warp == 6 && lane == 0) {
                Ring br;
                for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                    const int64_t kr = it / p.m_tiles;
                    const int64_t c_begin = kr * p.chunks_per_range;
                    const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_range);
                    int bc = (int)(c_begin % p.chunks_per_a);
                    for (int64_t c = c_begin; c < c_end; ++c) {
#pragma unroll
                        for (int u = 0; u < KO; ++u) {
                            mbar_wait(&b_empty[br.idx], br.phase ^ 1u);
                            mbar_expect_tx(&b_full[br.idx], C::B_UNIT);
                            unsigned char* dst = b_smem + br.idx * C::B_UNIT;
                            tma_load_2d(dst, &bhi_map, &b_full[br.idx], bc * KS + u * 32, 0);
                            tma_load_2d(dst + RP * 128, &blo_map, &b_full[br.idx], bc * KS + u * 32, 0);
                            br.advance(BS);
                        }
                        if (++bc == (int)p.chunks_per_a) bc = 0;
                    }
                }
            }
        } else {
            // Each of the 4 KR warps synthesises whole 32-element units on its own (units g = kw, kw+4, ...
            // of this CTA's unit sequence), so four units are in flight and no cross-warp barrier is needed.
            const int kw = warp - 6;
            constexpr int F4 = RP / 4;                    // float4 per lane per unit (a unit is 32 x RP floats)
            float* st = stage + kw * C::STAGE_F;         // rotated staging: (k, r) at st[k*RP + (r + k) % RP]
            float* prow = st + 32 * RP;                  // P[a, :] of the current `a`
            float4 qreg[F4];
            auto load_q = [&](int64_t b0) {
#pragma unroll
                for (int f = 0; f < F4; ++f) {
                    const int idx = lane + f * 32;        // float4 index within the unit, row-major [32][RP/4]
                    const int k = idx / F4;
                    qreg[f] = (b0 + k < p.B) ? __ldg(reinterpret_cast<const float4*>(p.Q + (b0 + k) * RP) + (idx % F4))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            uint32_t g_base = 0;                          // global index of the current item's first unit
            int a_loaded = -1;
            for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int64_t kr = it / p.m_tiles;
                const int64_t c_begin = kr * p.chunks_per_range;
                const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_range);
                const int n = (int)(c_end - c_begin) * KO;      // units of this item
                const int units_per_a = (int)p.chunks_per_a * KO;
                const int a0 = (int)(c_begin / p.chunks_per_a);
                const int bu0 = (int)(c_begin - (int64_t)a0 * p.chunks_per_a) * KO;
                int i = (int)((kw - (int)(g_base & 3u)) & 3);     // first local unit handled by this warp
                int a = a0, bu = bu0 + i;
                while (bu >= units_per_a) { bu -= units_per_a; ++a; }
                if (i < n) load_q((int64_t)bu * 32);
                for (; i < n; i += 4) {
                    const uint32_t g = g_base + (uint32_t)i;
                    const int slot = (int)(g & (BS - 1));
                    const uint32_t use_parity = (g / BS) & 1u;
                    __syncwarp();                          // previous unit's staging reads are done
#pragma unroll
                    for (int f = 0; f < F4; ++f) {
                        const int idx = lane + f * 32;
                        const int k = idx / F4, r = (idx % F4) * 4;
                        float* rowp = st + k * RP;
                        rowp[(r + 0 + k) & (RP - 1)] = qreg[f].x;
                        rowp[(r + 1 + k) & (RP - 1)] = qreg[f].y;
                        rowp[(r + 2 + k) & (RP - 1)] = qreg[f].z;
                        rowp[(r + 3 + k) & (RP - 1)] = qreg[f].w;
                    }
                    if (a != a_loaded) {
                        for (int r = lane; r < RP; r += 32) prow[r] = p.P ? __ldg(p.P + (int64_t)a * RP + r) : 1.0f;
                        a_loaded = a;
                    }
                    // coordinates of this warp's next unit; prefetch its Q rows from L2
                    int a_n = a, bu_n = bu + 4;
                    while (bu_n >= units_per_a) { bu_n -= units_per_a; ++a_n; }
                    if (i + 4 < n) load_q((int64_t)bu_n * 32);
                    __syncwarp();
                    mbar_wait(&b_empty[slot], use_parity ^ 1u);
                    unsigned char* bhi = b_smem + slot * C::B_UNIT;
                    unsigned char* blo = bhi + RP * 128;
                    const float* srow = st + lane * RP;                   // lane = k within the unit
                    if (!(p.debug & 1))
#pragma unroll
                    for (int r0 = 0; r0 < RP; r0 += 8) {
                        float qv[8], pv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { qv[j] = srow[(r0 + j + lane) & (RP - 1)]; pv[j] = prow[r0 + j]; }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int r = r0 + j;
                            const float krv = __fmul_rn(pv[j], qv[j]);
                            const uint32_t hbits = __float_as_uint(krv) & 0xFFFFE000u;
                            const float lo = krv - __uint_as_float(hbits);
                            const uint32_t off = r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4;
                            *reinterpret_cast<uint32_t*>(bhi + off) = hbits;
                            *reinterpret_cast<float*>(blo + off) = lo;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core reads
                    mbar_arrive(&b_full[slot]);
                    a = a_n; bu = bu_n;
                }
                g_base += (uint32_t)n;
            }
        }
    } else {
        // ================= epilogue: drain accumulation groups, write C =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t G = 0;
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int mt = (int)(it % p.m_tiles);
            const int64_t kr = it / p.m_tiles;
            const int64_t c_begin = kr * p.chunks_per_range;
            const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_range);
            const int n = (int)(c_end - c_begin) * KO;
            const int ngroups = (n + GU - 1) / GU;
            float acc[RP];
#pragma unroll
            for (int c = 0; c < RP; ++c) acc[c] = 0.f;
            for (int g = 0; g < ngroups; ++g) {
                const uint32_t buf = G & 1u;
                mbar_wait(&d_full[buf], (G >> 1) & 1u);
                tc_fence_after();
                if (!(p.debug & 8))
#pragma unroll
                for (int part = 0; part < 2; ++part) {        // hi*hi block, cross-term block of this set
#pragma unroll
                    for (int c0 = 0; c0 < RP; c0 += 32) {
                        uint32_t r[32];
                        TLB_TMEM_LD32(lane_addr + buf * C::D_COLS + part * RP + c0, r);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[c0 + c] += __uint_as_float(r[c]);
                    }
                }
                tc_fence_before();
                mbar_arrive(&d_empty[buf]);
                ++G;
            }
            const int64_t gm = (int64_t)mt * TM + row;
            if (gm < p.M) {
                float* dst = p.out + kr * p.sOk + gm * p.sOm;
                if (p.sOn == 1 && (p.n_valid & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                    for (int c = 0; c < RP / 4; ++c)
                        if (4 * c < p.n_valid)
                            reinterpret_cast<float4*>(dst)[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                } else {
#pragma unroll
                    for (int c = 0; c < RP; ++c)
                        if (c < p.n_valid) dst[c * p.sOn] = acc[c];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

template <int RP, int XL, int BM>
int launch_cfg(const TcStreamLaunch& l, cudaStream_t stream) {
    using C = Cfg<RP, XL>;
    const int smem = C::SMEM + 1024;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_stream_kernel<RP, XL, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return TLB200_ECUDA;
        attr = true;
    }
    int64_t n_items = (int64_t)l.p.m_tiles * l.p.k_ranges;
    if (n_items <= 0) return TLB200_OK;
    static int grid_cap = -1;
    if (grid_cap < 0) { const char* e = getenv("TLB200_TC_GRID"); grid_cap = e ? atoi(e) : kNumSMs; if (grid_cap < 1) grid_cap = kNumSMs; }
    const unsigned grid = (unsigned)(n_items < grid_cap ? n_items : grid_cap);
    tc_stream_kernel<RP, XL, BM><<<grid, NUM_THREADS, smem, stream>>>(l.x_map, l.bhi_map, l.blo_map, l.p);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <int RP, int XL>
int launch_bm(const TcStreamLaunch& l, cudaStream_t s) {
    return l.b_mode == TC_B_KR ? launch_cfg<RP, XL, TC_B_KR>(l, s) : launch_cfg<RP, XL, TC_B_MAT>(l, s);
}
template <int RP>
int launch_xl(const TcStreamLaunch& l, cudaStream_t s) {
    switch (l.x_layout) {
        case TC_X_KMAJOR_1: return launch_bm<RP, TC_X_KMAJOR_1>(l, s);
        case TC_X_KMAJOR_2: return launch_bm<RP, TC_X_KMAJOR_2>(l, s);
        case TC_X_MMAJOR: return launch_bm<RP, TC_X_MMAJOR>(l, s);
    }
    return TLB200_EINVAL;
}

}  // namespace

bool tc_available() { return get_encode_fn() != nullptr && !getenv("TLB200_DISABLE_TC"); }

int tc_encode_map(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return TLB200_EUNSUPPORTED;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, s, b, e,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TLB200_OK : TLB200_ECUDA;
}

int tc_group_units() {
    // accumulation-group length in 32-element K units.  Only the hi*hi products carry a significant
    // round-toward-zero loss (4 MMAs per unit, ~6e-8 each, measured): 8 units keep MTTKRP/TTM near 2e-6.
    static int units = -1;
    if (units < 0) {
        const char* e = getenv("TLB200_TC_FLUSH");
        units = e ? atoi(e) : 8;
        if (units < 1) units = 1;
    }
    return units;
}

int tc_stream_launch(const TcStreamLaunch& l_in, cudaStream_t stream) {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TLB200_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    TcStreamLaunch l = l_in;
    l.p.debug = dbg;
    if (l.rp == 32) return launch_xl<32>(l, stream);
    if (l.rp == 64) return launch_xl<64>(l, stream);
    return TLB200_EUNSUPPORTED;
}

}  // namespace tlb200
