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
// Warp roles (16 warps, 1 CTA per SM, persistent over work items):
//   warp 0       TMA producer of X tiles (ring of XS stages, XS even)
//   warps 1, 11  MMA issuers (one elected lane each): they take alternate tiles and pass a turn token, so one
//                polls the barriers of the next tile while the other's MMAs execute; their per-tile bookkeeping
//                is kept minimal (they share a scheduler with two convert warps and an epilogue warp)
//   warps 2-9    convert: smem X tile -> registers -> hi/lo -> TMEM A ring; two sets of four
//                warps (one warp per TMEM lane quarter) take alternate tiles, which hides the
//                barrier/LDS/tcgen05.st latencies of one tile behind the other
//   warp 10      B producer: TMA loads of the pre-split (tf32 hi / lo) small operand, K-major SWIZZLE_128B,
//                one 32-element unit at a time.  MTTKRP (b_resident): the work item's block of the inner
//                Khatri-Rao table Q is loaded ONCE and stays in shared memory while the item streams its `a`
//                range; TTM: the factor matrix is streamed with the tiles through the same slots as a ring.
//   warps 12-15  epilogue: drain TMEM accumulation groups, write C.  For MTTKRP the Khatri-Rao
//                row is P[a,:] * Q[b,:]: the tensor core contracts with Q only and the epilogue
//                scales each drained group by P[a,:] (groups never straddle an `a` boundary), so
//                no Khatri-Rao tile is ever formed — not even in shared memory.
#include "tc_stream.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

namespace tlb200 {
namespace {

constexpr int TM = 128;
constexpr int NUM_THREADS = 512;
// A wait gives up (trap) after ~20 s of wall time — never hang the GPU.  The clock is only consulted every 1024 polls:
// a poll may return at once or, with the suspend-time hint (100 us) honoured, sleep that long.  (The hint takes the
// spin loops from ~38 % of all issued warp instructions to a few percent — same-box A/B: SM clock up ~100 MHz under the
// power cap at the same 4.85-4.9 TB/s: the sustained rate is set by the energy of the data path, not by issue slots.)
constexpr uint64_t WAIT_LIMIT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void wait_watchdog(uint32_t& spins, uint64_t& t0) {
    if ((++spins & 0x3FFu) != 0) return;
    uint64_t now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    else if (now - t0 > WAIT_LIMIT_NS) asm volatile("trap;");
}

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
    uint64_t t0 = 0;
    const uint32_t addr = smem_u32(bar);
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x186A0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        wait_watchdog(spins, t0);
    }
}
// One elected lane of a fully converged warp (the operands of tcgen05/TMA instructions must live in uniform
// registers: computing them warp-uniformly and predicating only the issue avoids a per-instruction
// R2UR "waterfall" loop on the issuing thread — measured ~95 clk per MMA with `if (lane == 0)`).
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
    return pred;
}
// wait for two barriers at once: the two try_wait round trips overlap instead of adding up
__device__ __forceinline__ void mbar_wait2(uint64_t* bar_a, uint32_t parity_a, uint64_t* bar_b, uint32_t parity_b) {
    uint32_t da = 0, db = 0, spins = 0;
    uint64_t t0 = 0;
    const uint32_t aa = smem_u32(bar_a), ab = smem_u32(bar_b);
    while (true) {
        if (!da) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x186A0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                              : "=r"(da) : "r"(aa), "r"(parity_a) : "memory");
        if (!db) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x186A0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                              : "=r"(db) : "r"(ab), "r"(parity_b) : "memory");
        if (da && db) break;
        wait_watchdog(spins, t0);
    }
}
// three barriers at once (A tile + both B units of a tile)
__device__ __forceinline__ void mbar_wait3(uint64_t* bar_a, uint32_t pa, uint64_t* bar_b, uint32_t pb, uint64_t* bar_c, uint32_t pc) {
    uint32_t da = 0, db = 0, dc = 0, spins = 0;
    uint64_t t0 = 0;
    const uint32_t aa = smem_u32(bar_a), ab = smem_u32(bar_b), ac = smem_u32(bar_c);
    while (true) {
        if (!da) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x186A0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                              : "=r"(da) : "r"(aa), "r"(pa) : "memory");
        if (!db) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x186A0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                              : "=r"(db) : "r"(ab), "r"(pb) : "memory");
        if (!dc) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x186A0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                              : "=r"(dc) : "r"(ac), "r"(pc) : "memory");
        if (da && db && dc) break;
        wait_watchdog(spins, t0);
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
// same, accumulate flag hard-wired to true (no predicate register set-up on the issuing thread)
__device__ __forceinline__ void mma_ts_tf32_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
}
// kind::f16 (fp16 operands, fp32 accumulate), A from TMEM: K = 16 per instruction at the cost of a K = 8 tf32 one
__device__ __forceinline__ void mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_f16_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
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

// kind::f16 with fp16 A and B, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the power-of-two scale that maps |x| <= absmax into [2^14, 2^15) (fp16 range), and its inverse
__device__ __forceinline__ float hf_scale(const float* __restrict__ absmax, float* inv) {
    const int e = (int)((__float_as_uint(__ldg(absmax)) >> 23) & 0xFFu);
    int se = 268 - e;                       // biased exponent of 2^(14 - (e - 127))
    se = se < 1 ? 1 : (se > 253 ? 253 : se);
    *inv = __uint_as_float((uint32_t)(254 - se) << 23);
    return __uint_as_float((uint32_t)se << 23);
}
// fp16 split of two scaled values: hi = RN_fp16(x), lo = RN_fp16((x - hi) * 2^11); x0 goes to the low half (even k)
__device__ __forceinline__ void hf_split2(float x0, float x1, uint32_t& hi2, uint32_t& lo2) {
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(x1), "f"(x0));
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
    const float l0 = (x0 - h.x) * 2048.f, l1 = (x1 - h.y) * 2048.f;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(l1), "f"(l0));
}

#define TLB_TMEM_ST16(taddr, r)                                                                                          \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                        \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                             \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),  \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])                                \
                 : "memory")
#define TLB_TMEM_LD32(taddr, r)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),         \
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),        \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                      \
                 : "r"(taddr))
#define TLB_TMEM_LD16(taddr, r)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                             \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                      \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                       \
                 : "r"(taddr))
#define TLB_TMEM_ST32(taddr, r)                                                                                          \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                        \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),  \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),       \
                 "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),      \
                 "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                       \
                 : "memory")

// perf triage: timestamp event `ev` of iteration `i` of role `role` (CTA 0 only, first 256 iterations)
#define TLB_TRACE(role, i, ev)                                                                   \
    do { if (tr && (i) < 256) p.trace[(((role) * 256) + (i)) * 8 + (ev)] = clock64(); } while (0)

// ring position: stage index + phase parity, advanced without divisions
struct Ring {
    int idx = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1u; } }
};

// work item = (row tile, b block, a range); consecutive items differ in the row tile first
struct TcItem {
    int mt;          // row tile
    int a0, a1;      // range of the outer K index
    int bc0, bc1;    // range of K chunks (KS elements each) inside every `a`
    int64_t part;    // index of the partial result this item writes
};
__device__ __forceinline__ TcItem tc_item(const TcStreamParams& p, int64_t it) {
    TcItem t;
    t.mt = (int)(it % p.m_tiles);
    const int64_t r = it / p.m_tiles;
    const int bb = (int)(r % p.n_bblocks);
    const int64_t kr = r / p.n_bblocks;
    t.part = r;
    t.a0 = (int)min(p.A, kr * p.a_per_range);
    t.a1 = (int)min(p.A, (int64_t)t.a0 + p.a_per_range);
    t.bc0 = bb * p.nb;
    t.bc1 = (int)min(p.chunks_per_a, (int64_t)t.bc0 + p.nb);
    return t;
}

// ---- compile-time configuration per (RP, X layout) -------------------------------------
// HF: the fp16-split engine (range hint present) — half the MMA instructions and TMEM/shared-memory operand traffic
// of the tf32 one.  A units are [hi 16 | lo 16] columns of packed halves; a B slot holds 64 contraction elements
// (one whole X tile) per 128-byte row, so the K-step arithmetic on the B descriptor is the tf32 engine's.
template <int RP, int XL, bool HF>
struct Cfg {
    static constexpr int KS = XL == TC_X_KMAJOR_1 ? 32 : 64;        // K extent of one X stage
    static constexpr int KO = KS / 32;                              // 32-element units per X stage
    static constexpr int X_STAGE = TM * KS * 4;
    static constexpr int XS = KS == 32 ? 8 : 4;                     // bytes in flight per SM hide HBM latency
    static constexpr int D_COLS = 2 * RP;                           // per accumulator set: [hi*hi (RP) | hi*lo + lo*hi (RP)]
    static constexpr int BKO = HF ? 1 : KO;                         // B slots per X tile
    static constexpr int A_COLS = HF ? 32 : 64;                     // TMEM columns per A unit: [hi 32 | lo 32] (HF: packed halves)
    static constexpr int AS = HF ? 8 : (RP == 32 ? 6 : 4);
    static constexpr int B_UNIT = 2 * RP * 128;                     // [hi RP rows | lo RP rows] x 128 B, K-major SW128
    static constexpr int BS = RP == 32 ? 8 : 6;                     // B-operand slots (32-element units; HF: 64): 64 / 96 KB, a ring
                                                                    // (streamed B) or one resident b block per item
    static constexpr int OFF_X = 0;
    static constexpr int OFF_B = OFF_X + XS * X_STAGE;
    static constexpr int OFF_BAR = OFF_B + BS * B_UNIT;
    static constexpr int NUM_BARS = 2 * XS + 2 * AS + 2 * BS + 4 + 2;   // + the two MMA-issuer turn tokens
    static constexpr int SMEM = OFF_BAR + NUM_BARS * 8 + 16;
    static constexpr int TMEM_COLS = 2 * D_COLS + AS * A_COLS;
    static_assert(AS >= 2 && AS % KO == 0, "A ring must hold whole tiles");
    static_assert(TMEM_COLS <= 512, "TMEM budget");
    // the two convert sets take alternate tiles: with an even ring every stage always belongs to the same set,
    // so a set observes every phase of the barriers it waits on (an odd ring would alias phase parities)
    static_assert(XS % 2 == 0, "X ring must be even");
    static_assert(BS % BKO == 0, "B ring must hold whole tiles");
    static_assert(!HF || KS == 64, "the fp16 engine takes 64-element tiles only");
    static_assert(SMEM + 1024 <= 227 * 1024, "smem budget");
};

template <int RP, int XL, bool HF>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_stream_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap bhi_map,
                 const __grid_constant__ CUtensorMap blo_map, const TcStreamParams p) {
    using C = Cfg<RP, XL, HF>;
    constexpr int KS = C::KS, KO = C::KO, XS = C::XS, AS = C::AS, BS = C::BS, BKO = C::BKO;
    constexpr bool kUnitRelease = RP == 64 && KO == 2 && !HF;     // A-ring slots are handed back unit by unit
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* x_smem = smem + C::OFF_X;
    unsigned char* b_smem = smem + C::OFF_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + XS;
    uint64_t* a_full = x_empty + XS;
    uint64_t* a_empty = a_full + AS;
    uint64_t* b_full = a_empty + AS;
    uint64_t* b_empty = b_full + BS;
    uint64_t* d_full = b_empty + BS;
    uint64_t* d_empty = d_full + 2;
    uint64_t* tok = d_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_items = (int64_t)p.m_tiles * p.n_bblocks * p.k_ranges;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
    int tri = 0;   // trace iteration counter of this role

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 128); }
        for (int i = 0; i < AS; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < BS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 128); mbar_init(&tok[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&bhi_map)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&blo_map)) : "memory");
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
        // ================= TMA producer of X tiles (whole warp loops, one elected lane issues) =================
        {
            Ring xr;
            for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                const TcItem t = tc_item(p, it);
                const int m0 = t.mt * TM;
                for (int a = t.a0; a < t.a1; ++a) {
                    for (int bc = t.bc0; bc < t.bc1; ++bc) {
                        mbar_wait(&x_empty[xr.idx], xr.phase ^ 1u);
                        TLB_TRACE(0, tri, 0);
                        if (elect_one_sync()) {
                            mbar_expect_tx(&x_full[xr.idx], C::X_STAGE);
                            unsigned char* dst = x_smem + xr.idx * C::X_STAGE;
                            if constexpr (XL == TC_X_KMAJOR_1)      tma_load_3d(dst, &xmap, &x_full[xr.idx], bc * 32, m0, a);
                            else if constexpr (XL == TC_X_KMAJOR_2) tma_load_4d(dst, &xmap, &x_full[xr.idx], 0, bc * 2, m0, a);
                            else                                    tma_load_3d(dst, &xmap, &x_full[xr.idx], m0, bc * 64, a);
                        }
                        __syncwarp();
                        TLB_TRACE(0, tri, 1); ++tri;
                        xr.advance(XS);
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 11) {
        // ================= MMA issuers (two warps take alternate tiles; whole warp loops, one elected lane issues) ====
        // While one warp's MMAs execute, the other has already polled the barriers of the next tile and only waits
        // for its turn, so the tensor pipe no longer idles through the ~250 clk barrier round trips.
        {
            const int mw = warp == 1 ? 0 : 1;       // this issuer takes tiles with (global tile index & 1) == mw
            uint32_t gt = 0;                        // global tile index (per CTA)
            uint32_t own = 0;                       // tiles issued by this warp so far
            constexpr uint32_t idesc1 = HF ? idesc_f16(TM, 2 * RP) : idesc_tf32(TM, 2 * RP);   // A_hi x [B_hi | B_lo] -> columns [0, 2RP)
            constexpr uint32_t idesc2 = HF ? idesc_f16(TM, RP) : idesc_tf32(TM, RP);           // A_lo x B_hi          -> columns [RP, 2RP)
            Ring ar, br;
            uint32_t G = 0;          // global accumulation-group counter
            // loop-invariant operand pieces: the issuing thread is a single latency-bound instruction
            // stream, so every add it does not execute shortens the MMA cadence
            const uint32_t d_base = tmem_base;
            const uint32_t a_base = tmem_base + a_col0;
            const uint64_t d0 = desc_kmajor_sw128(smem_u32(b_smem));
            const uint32_t bdesc_lo0 = (uint32_t)d0, bdesc_hi = (uint32_t)(d0 >> 32);
            uint32_t bphase = 0;     // resident B: phase bit per slot
            for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                const TcItem t = tc_item(p, it);
                const int nbc = t.bc1 - t.bc0;                  // tiles per `a` row of this item, KO units each
                if (t.a1 <= t.a0 || nbc <= 0) continue;
                int ug = 0;              // units already in the open accumulation group (a multiple of KO: GU % KO == 0)
                const bool scaled = p.P != nullptr;
                for (int a = t.a0; a < t.a1; ++a)
                for (int j = 0; j < nbc; ++j, ++gt) {
                    // Both issuers walk every tile; this part runs twice per issued tile on a single latency-bound
                    // instruction stream (it was 1100 clk per own tile before it was pared down), so only the state
                    // that must advance for the other issuer's tiles is touched before the `continue`.
                    const bool last_a = a + 1 == t.a1;
                    const bool cut = j + 1 == nbc && (last_a || scaled);     // the group may not continue past this tile
                    const bool first = ug == 0;                              // this tile opens a group
                    const bool gend = ug + KO == GU || cut;                  // ... closes it
                    const uint32_t Gc = G;
                    if (gend) { ++G; ug = 0; } else { ug += KO; }
                    const int as0 = ar.idx;
                    const uint32_t aph = ar.phase;
                    ar.idx += KO;
                    if (ar.idx == AS) { ar.idx = 0; ar.phase ^= 1u; }
                    const Ring bs0 = br;                                      // streamed B: ring position of unit 0
                    br.idx += BKO;
                    if (br.idx == BS) { br.idx = 0; br.phase ^= 1u; }
                    if (((gt ^ (uint32_t)mw) & 1u) != 0) continue;            // the other issuer's tile
                    TLB_TRACE(1 + 2 * mw, tri, 4);
                    const uint32_t dbuf = Gc & 1u;
                    if (first && Gc >= 2) mbar_wait(&d_empty[dbuf], ((Gc >> 1) - 1) & 1u);
                    Ring b0 = bs0, b1 = bs0;
                    if (p.b_resident) {      // slot = position inside the item's b block, one phase per item
                        b0.idx = j * BKO; b0.phase = (bphase >> b0.idx) & 1u;
                        b1.idx = j * BKO + (BKO - 1); b1.phase = (bphase >> b1.idx) & 1u;
                    } else if constexpr (BKO == 2) {
                        b1.idx = bs0.idx + 1;                                 // BS is even: a tile never wraps inside
                    }
                    TLB_TRACE(1 + 2 * mw, tri, 0);
                    // one overlapped wait per tile: the A tile and its B unit(s)
                    if constexpr (BKO == 2) mbar_wait3(&a_full[as0], aph, &b_full[b0.idx], b0.phase, &b_full[b1.idx], b1.phase);
                    else mbar_wait2(&a_full[as0], aph, &b_full[b0.idx], b0.phase);
                    // my turn: the other issuer has put its tile into the pipe (accumulate order within a group)
                    if (mw == 1) mbar_wait(&tok[1], own & 1u);
                    else if (own > 0) mbar_wait(&tok[0], (own - 1) & 1u);
                    TLB_TRACE(1 + 2 * mw, tri, 2);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        if constexpr (HF) {
                            if (!(p.debug & 2)) {
                                const uint32_t d1 = d_base + dbuf * C::D_COLS;
                                const uint32_t d2 = d1 + RP;
                                const uint32_t blo = bdesc_lo0 + b0.idx * (C::B_UNIT >> 4);
                                // 4 K steps of 16: step ks reads A unit ks / 2, halves [8 (ks & 1), +8) of its hi / lo
                                // columns, and bytes [32 ks, +32) of the B rows
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint32_t a_hi = a_base + (as0 + (ks >> 1)) * C::A_COLS + (ks & 1) * 8;
                                    const uint64_t db = ((uint64_t)bdesc_hi << 32) | (blo + ks * 2);
                                    if (ks == 0) mma_ts_f16(d1, a_hi, db, idesc1, first ? 0u : 1u);
                                    else mma_ts_f16_acc(d1, a_hi, db, idesc1);
                                    mma_ts_f16_acc(d2, a_hi + 16, db, idesc2);
                                }
                            }
                        } else if (!(p.debug & 2)) {
#pragma unroll
                            for (int u = 0; u < KO; ++u) {
                                const uint32_t d1 = d_base + dbuf * C::D_COLS;
                                const uint32_t d2 = d1 + RP;
                                const uint32_t a_hi = a_base + (as0 + u) * C::A_COLS;       // a_lo = a_hi + 32
                                const uint32_t blo = bdesc_lo0 + (u == 0 ? b0.idx : b1.idx) * (C::B_UNIT >> 4); // low descriptor word
                                // first K step of a group overwrites the accumulators, everything else accumulates
                                mma_ts_tf32(d1, a_hi, ((uint64_t)bdesc_hi << 32) | blo, idesc1, (first && u == 0) ? 0u : 1u);
                                mma_ts_tf32_acc(d2, a_hi + 32, ((uint64_t)bdesc_hi << 32) | blo, idesc2);
#pragma unroll
                                for (int ks = 1; ks < 4; ++ks) {
                                    const uint64_t db = ((uint64_t)bdesc_hi << 32) | (blo + ks * 2);   // +32 bytes per K step
                                    mma_ts_tf32_acc(d1, a_hi + ks * 8, db, idesc1);
                                    mma_ts_tf32_acc(d2, a_hi + 32 + ks * 8, db, idesc2);
                                }
                                // rank 64: the A ring holds only two tiles (TMEM budget), so convert and MMA take turns on
                                // a slot; releasing the first unit before the second one is issued lets the converters
                                // refill it while the second unit's MMAs execute (measured cycle per slot 2200 -> 1800 clk)
                                if constexpr (kUnitRelease) {
                                    if (u == 0) tc_commit(&a_empty[as0]);
                                }
                            }
                        }
                        // hand the turn over before the (slower) commits
                        tc_fence_before();
                        mbar_arrive(&tok[mw ^ 1]);
                        if (!p.b_resident || last_a) {     // a resident B slot is released by its last reader only
                            tc_commit(&b_empty[b0.idx]);
                            if constexpr (BKO == 2) tc_commit(&b_empty[b1.idx]);
                        }
                        if (kUnitRelease && (p.debug & 2)) tc_commit(&a_empty[as0]);      // perf triage: MMAs skipped
                        tc_commit(&a_empty[kUnitRelease ? as0 + KO - 1 : as0]);
                        if (gend) tc_commit(&d_full[dbuf]);
                    }
                    __syncwarp();
                    TLB_TRACE(1 + 2 * mw, tri, 3); ++tri;
                    ++own;
                }
                bphase ^= (1u << (nbc * BKO)) - 1u;
            }
        }
    } else if (warp < 10) {
        // ================= convert: smem X tile -> hi/lo -> TMEM A ring =================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int cset = (warp - 2) >> 2;       // convert set 0 / 1: takes the tiles with (global index & 1) == cset
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t cbase = 0;        // global index (per CTA) of the current item's first tile
        int pub_slot[2] = {0, 0};  // A slots stored but not yet published
        int unpublished = 0;       // A units whose tcgen05.st have been issued but not yet waited for
        float hf_inv = 1.f;
        const float hf_s = HF ? hf_scale(p.x_absmax, &hf_inv) : 1.f;
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
            const TcItem t = tc_item(p, it);
            const int n = max(0, t.a1 - t.a0) * max(0, t.bc1 - t.bc0);    // tiles of this item
            for (int i = (int)((cset - (int)(cbase & 1u)) & 1); i < n; i += 2) {
                const uint32_t gc = cbase + (uint32_t)i;                 // global tile index
                const int xs = (int)(gc % XS);
                const uint32_t xphase = (gc / XS) & 1u;
                if (warp == 2) TLB_TRACE(2, tri, 0);
                mbar_wait(&x_full[xs], xphase);
                if (warp == 2) TLB_TRACE(2, tri, 1);
                const unsigned char* xt = x_smem + xs * C::X_STAGE;
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
                // Hand the X stage back to the TMA producer as early as possible (bytes in flight hide the
                // HBM latency) — but only once every load has really landed in registers: touching one
                // register of each 128-bit load makes the scoreboard wait for it, and the proxy fence orders
                // these generic-proxy reads before the async-proxy overwrite.  (Releasing right after *issuing*
                // the loads let TMA overwrite rows that slower warps had not read yet.)
                {
                    uint32_t touch = 0;
                    if constexpr (XL == TC_X_MMAJOR) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) touch ^= ta[k] ^ tb[k & (KS == 64 ? 31 : 0)];
                    } else {
#pragma unroll
                        for (int c = 0; c < 8; ++c) touch ^= ta[4 * c] ^ tb[(4 * c) & (KS == 64 ? 31 : 0)];
                    }
                    asm volatile("" ::"r"(touch));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&x_empty[xs]);
                }
                if (warp == 2) TLB_TRACE(2, tri, 2);
                // Publish the A units stored in the previous iteration (their tcgen05.st have had a whole
                // iteration to complete).
                if (unpublished) {
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    for (int j = 0; j < unpublished; ++j) mbar_arrive(&a_full[pub_slot[j]]);
                    unpublished = 0;
                }
                if (warp == 2) TLB_TRACE(2, tri, 3);
                uint32_t h[32];
#pragma unroll
                for (int u = 0; u < KO; ++u) {
                    const uint32_t gu = gc * KO + u;                      // global unit index
                    const int as = (int)(gu % AS);
                    if (u == 0 || kUnitRelease) {                         // one release per tile (slot of unit 0) or per unit
                        mbar_wait(&a_empty[as], ((gu / AS) & 1u) ^ 1u);
                        tc_fence_after();
                    }
                    if (warp == 2) TLB_TRACE(2, tri, 4 + u);
                    const uint32_t abase = lane_addr + a_col0 + as * C::A_COLS;
                    if constexpr (HF) {
                        if (!(p.debug & 4)) {
                            // unit u of this row: ta for (u == 0) != flip, else tb; scaled into fp16 range, split, packed
                            const bool first_line = (u == 0) != (flip != 0);
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float x0 = __uint_as_float(first_line ? ta[2 * j] : tb[(2 * j) & (KS == 64 ? 31 : 0)]) * hf_s;
                                const float x1 = __uint_as_float(first_line ? ta[2 * j + 1] : tb[(2 * j + 1) & (KS == 64 ? 31 : 0)]) * hf_s;
                                hf_split2(x0, x1, h[j], h[16 + j]);
                            }
                            TLB_TMEM_ST16(abase, h);
                            TLB_TMEM_ST16(abase + 16, (h + 16));
                        }
                    } else if (!(p.debug & 4)) {
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
                    if (u == 0) { pub_slot[0] = as; unpublished = 1; }
                }
                // publish right away: the other convert set hides this wait
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                for (int j = 0; j < unpublished; ++j) mbar_arrive(&a_full[pub_slot[j]]);
                unpublished = 0;
                if (warp == 2) { TLB_TRACE(2, tri, 6); ++tri; }
            }
            cbase += (uint32_t)n;
        }
        if (unpublished) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            for (int j = 0; j < unpublished; ++j) mbar_arrive(&a_full[pub_slot[j]]);
        }
    } else if (warp < 12) {
        // ================= B producer: TMA loads of the pre-split small operand, one 32-element unit at a time =========
        if (warp == 10) {
            Ring br;
            uint32_t bphase = 0;     // resident B: phase bit per slot
            int nb_loaded = 0;
            for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
                const TcItem t = tc_item(p, it);
                const int nbc = t.bc1 - t.bc0;
                if (t.a1 <= t.a0 || nbc <= 0) continue;
                if (p.b_resident) {
                    // the item's b block is loaded once and stays in shared memory for all of its `a` rows
                    for (int sl = 0; sl < nbc * BKO; ++sl) {
                        mbar_wait(&b_empty[sl], ((bphase >> sl) & 1u) ^ 1u);
                        if (elect_one_sync()) {
                            mbar_expect_tx(&b_full[sl], C::B_UNIT);
                            unsigned char* dst = b_smem + sl * C::B_UNIT;
                            tma_load_2d(dst, &bhi_map, &b_full[sl], t.bc0 * KS + sl * (KS / BKO), 0);
                            tma_load_2d(dst + RP * 128, &blo_map, &b_full[sl], t.bc0 * KS + sl * (KS / BKO), 0);
                        }
                        __syncwarp();
                    }
                    bphase ^= (1u << (nbc * BKO)) - 1u;
                    continue;
                }
                for (int a = t.a0; a < t.a1; ++a)
                for (int bc = t.bc0; bc < t.bc1; ++bc) {
#pragma unroll
                    for (int u = 0; u < BKO; ++u) {
                        mbar_wait(&b_empty[br.idx], br.phase ^ 1u);
                        if (elect_one_sync()) {
                            if ((p.debug & 8) && nb_loaded >= BS) {
                                mbar_arrive(&b_full[br.idx]);      // perf triage: reuse the stale tile, no L2 traffic
                            } else {
                                mbar_expect_tx(&b_full[br.idx], C::B_UNIT);
                                unsigned char* dst = b_smem + br.idx * C::B_UNIT;
                                tma_load_2d(dst, &bhi_map, &b_full[br.idx], bc * KS + u * (KS / BKO), 0);
                                tma_load_2d(dst + RP * 128, &blo_map, &b_full[br.idx], bc * KS + u * (KS / BKO), 0);
                            }
                        }
                        ++nb_loaded;
                        __syncwarp();
                        br.advance(BS);
                    }
                }
            }
        }
    } else {
        // ================= epilogue: drain accumulation groups, write C =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t G = 0;
        for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
            const TcItem t = tc_item(p, it);
            const int mt = t.mt;
            const int units_per_a = max(0, t.bc1 - t.bc0) * KO;           // 32-element units per `a` row of this item
            const int n = max(0, t.a1 - t.a0) * units_per_a;              // units of this item (0: the partial is zero)
            int a = t.a0;
            int bu = 0;                                                   // unit index within the current `a` row
            float acc[RP];
#pragma unroll
            for (int c = 0; c < RP; ++c) acc[c] = 0.f;
            int done = 0;
            while (done < n) {
                // this group: up to GU units, never across an `a` boundary when the result is scaled by P[a, :]
                int len = min(GU, n - done);
                if (p.P != nullptr) len = min(len, units_per_a - bu);
                const uint32_t buf = G & 1u;
                if (warp == 12) TLB_TRACE(4, tri, 0);
                mbar_wait(&d_full[buf], (G >> 1) & 1u);
                if (warp == 12) TLB_TRACE(4, tri, 1);
                tc_fence_after();
                // 16-column chunks at rank 64 keep accumulators + staging inside the register budget (no spills);
                // the P row of this `a` is fetched while the TMEM load is in flight
                constexpr int CW = 16;
                const float4* prow4 = p.P != nullptr ? reinterpret_cast<const float4*>(p.P + (int64_t)a * RP) : nullptr;
#pragma unroll
                for (int c0 = 0; c0 < RP; c0 += CW) {
                    uint32_t r0[CW], r1[CW];
                    if constexpr (CW == 32) {
                        TLB_TMEM_LD32(lane_addr + buf * C::D_COLS + c0, r0);            // hi*hi block
                        TLB_TMEM_LD32(lane_addr + buf * C::D_COLS + RP + c0, r1);       // cross-term block
                    } else {
                        TLB_TMEM_LD16(lane_addr + buf * C::D_COLS + c0, r0);
                        TLB_TMEM_LD16(lane_addr + buf * C::D_COLS + RP + c0, r1);
                    }
                    float4 pv[CW / 4];
                    if (prow4 != nullptr) {
#pragma unroll
                        for (int k = 0; k < CW / 4; ++k) pv[k] = __ldg(prow4 + c0 / 4 + k);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c0 + CW == RP) {            // every column of the buffer has been read
                        tc_fence_before();
                        mbar_arrive(&d_empty[buf]);
                    }
                    if constexpr (HF) {
                        // the cross-term block was formed from lo parts scaled by 2^11
#pragma unroll
                        for (int c = 0; c < CW; ++c) r1[c] = __float_as_uint(__uint_as_float(r1[c]) * (1.f / 2048.f));
                    }
                    if (prow4 != nullptr) {
#pragma unroll
                        for (int k = 0; k < CW / 4; ++k) {
                            const int c = c0 + 4 * k;
                            acc[c + 0] = fmaf(pv[k].x, __uint_as_float(r0[4 * k + 0]) + __uint_as_float(r1[4 * k + 0]), acc[c + 0]);
                            acc[c + 1] = fmaf(pv[k].y, __uint_as_float(r0[4 * k + 1]) + __uint_as_float(r1[4 * k + 1]), acc[c + 1]);
                            acc[c + 2] = fmaf(pv[k].z, __uint_as_float(r0[4 * k + 2]) + __uint_as_float(r1[4 * k + 2]), acc[c + 2]);
                            acc[c + 3] = fmaf(pv[k].w, __uint_as_float(r0[4 * k + 3]) + __uint_as_float(r1[4 * k + 3]), acc[c + 3]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < CW; ++c) acc[c0 + c] += __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
                    }
                }
                if (warp == 12) { TLB_TRACE(4, tri, 2); ++tri; }
                ++G;
                done += len;
                bu += len;
                if (bu >= units_per_a) { bu = 0; ++a; }
            }
            if constexpr (HF) {       // undo the power-of-two scales of the tensor and of the small operand's columns
                float xinv;
                hf_scale(p.x_absmax, &xinv);
                const float4* ci = reinterpret_cast<const float4*>(p.col_inv);
#pragma unroll
                for (int c = 0; c < RP / 4; ++c) {
                    const float4 v = __ldg(ci + c);
                    acc[4 * c] *= v.x * xinv; acc[4 * c + 1] *= v.y * xinv; acc[4 * c + 2] *= v.z * xinv; acc[4 * c + 3] *= v.w * xinv;
                }
            }
            const int64_t gm = (int64_t)mt * TM + row;
            if (gm < p.M) {
                float* dst = p.out + t.part * p.sOk + gm * p.sOm;
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

template <int RP, int XL, bool HF>
int launch_cfg(const TcStreamLaunch& l, cudaStream_t stream) {
    using C = Cfg<RP, XL, HF>;
    const int smem = C::SMEM + 1024;
    static std::atomic<uint64_t> attr_done{0};        // per instantiation, one bit per device
    if (ensure_dynamic_smem(tc_stream_kernel<RP, XL, HF>, smem, attr_done)) return TLB200_ECUDA;
    int64_t n_items = (int64_t)l.p.m_tiles * l.p.n_bblocks * l.p.k_ranges;
    if (n_items <= 0) return TLB200_OK;
    static int grid_cap = -1;
    if (grid_cap < 0) { const char* e = getenv("TLB200_TC_GRID"); grid_cap = e ? atoi(e) : kNumSMs; if (grid_cap < 1) grid_cap = kNumSMs; }
    const unsigned grid = (unsigned)(n_items < grid_cap ? n_items : grid_cap);
    tc_stream_kernel<RP, XL, HF><<<grid, NUM_THREADS, smem, stream>>>(l.x_map, l.bhi_map, l.blo_map, l.p);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <int RP>
int launch_xl(const TcStreamLaunch& l, cudaStream_t s) {
    if (l.hf) {
        if (!l.p.x_absmax || !l.p.col_inv) return TLB200_EINVAL;
        switch (l.x_layout) {
            case TC_X_KMAJOR_2: return launch_cfg<RP, TC_X_KMAJOR_2, true>(l, s);
            case TC_X_MMAJOR: return launch_cfg<RP, TC_X_MMAJOR, true>(l, s);
        }
        return TLB200_EUNSUPPORTED;
    }
    switch (l.x_layout) {
        case TC_X_KMAJOR_1: return launch_cfg<RP, TC_X_KMAJOR_1, false>(l, s);
        case TC_X_KMAJOR_2: return launch_cfg<RP, TC_X_KMAJOR_2, false>(l, s);
        case TC_X_MMAJOR: return launch_cfg<RP, TC_X_MMAJOR, false>(l, s);
    }
    return TLB200_EINVAL;
}

}  // namespace

bool tc_available() { return get_encode_fn() != nullptr && !getenv("TLB200_DISABLE_TC"); }

int tc_encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128, bool half) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return TLB200_EUNSUPPORTED;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                     const_cast<void*>(base), d, s, b, e,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? TLB200_OK : TLB200_ECUDA;
}

int tc_group_units(bool hf) {
    // accumulation-group length in 32-element K units.  Only the hi*hi products carry a significant
    // round-toward-zero loss (4 MMAs per unit, ~6e-8 each, measured): 8 units keep MTTKRP/TTM near 2e-6.
    // The fp16 engine issues 2 hi*hi MMAs per unit: 16 units are the same number of roundings.
    static int units = -1, units_hf = -1;
    if (units < 0) {
        const char* e = getenv("TLB200_TC_FLUSH");
        units = e ? atoi(e) : 8;
        if (units < 2) units = 2;
        units &= ~1;          // whole tiles (2 units) per group
        const char* h = getenv("TLB200_TC_FLUSH_F16");
        units_hf = h ? atoi(h) : 16;
        if (units_hf < 2) units_hf = 2;
        units_hf &= ~1;
    }
    return hf ? units_hf : units;
}

static long long* g_trace = nullptr;

int tc_stream_launch(const TcStreamLaunch& l_in, cudaStream_t stream) {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TLB200_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    TcStreamLaunch l = l_in;
    l.p.debug = dbg;
    l.p.trace = g_trace;
    if (l.rp == 32) return launch_xl<32>(l, stream);
    if (l.rp == 64) return launch_xl<64>(l, stream);
    return TLB200_EUNSUPPORTED;
}

}  // namespace tlb200

// perf triage hook (not part of the public ABI): device buffer of 5*256*8 int64 receiving CTA 0's timestamps
extern "C" void tlb200_debug_set_trace(void* device_buffer) { tlb200::g_trace = static_cast<long long*>(device_buffer); }
