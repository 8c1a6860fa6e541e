// CP reconstruction / masked-ALS imputation on the tensor cores (fp32, rank <= 64).
//
//   out[i, c] = sum_r (w_r F_0[i, r]) * K[c, r],   K[c, :] = prod_{n >= 1} F_n[i_n(c), :]   (c = the C-contiguous
//   column index over modes 1..N-1) — tensorly/cp_tensor.py:433-485, and the mask branch of error_calc,
//   tensorly/decomposition/_cp.py:195-207 (out = x*mask + rec*(1-mask) with both norms) as an epilogue variant.
//
// Why: the SIMT kernel (recon.cu) is bound by fp32 FMA issue from rank 16 up — 1.2 TB/s of output at rank 32, and the
// imputation pass it feeds was 80 % of a masked ALS sweep.  The product is a GEMM with a tiny contraction (K = R) and
// a streaming output, so here the OUTPUT streams: a persistent CTA per SM walks 128 x 128 output tiles,
//   warps 1-8  form the operand tiles in shared memory — the mode-0 factor tile (once per row block) and the
//              Khatri-Rao tile of the tile's 128 columns (never materialised in memory), each split exactly into
//              tf32 hi + lo and stored in the canonical K-major SWIZZLE_128B layout,
//   warp 0     issues D[:, 0:256] = A_hi [K_hi | K_lo] and D[:, 128:256] += A_lo K_hi per 8-wide K step (3xTF32 with
//              two instructions per step, like tc_stream.cu) into one of two TMEM accumulator sets,
//   warps 9-12 drain the other set 32 columns at a time: hi*hi + cross columns, the epilogue variant, and the result
//              into a swizzled shared-memory slot that one thread hands to TMA (cp.async.bulk.tensor store): no
//              global access is issued by a lane, so nothing waits on HBM latency and edges are clipped by TMA,
//   warp 13    (imputation / masked variants) TMA-loads the x and mask boxes of the slots ahead of the epilogue.
// Bound: HBM (4 bytes written per element; 12 bytes moved for the imputation pass).
#include "recon_tc.cuh"
#include "tc_stream.cuh"

#include <cstdlib>
#include <cstring>

namespace tlb200 {
namespace {

constexpr int RT_M = 128, RT_N = 128;
constexpr int RT_THREADS = 14 * 32;     // warp 0 MMA, 1-8 producers, 9-12 epilogue (one per TMEM lane quarter), 13 TMA loader
constexpr int RT_SC = 32;             // columns staged per step
constexpr int RT_SLD = RT_SC + 4;     // staging row pitch in floats (conflict-free 128-bit row writes)
constexpr uint32_t RT_SPIN_LIMIT = 1u << 22;

__device__ __forceinline__ uint32_t rt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rt_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rt_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rt_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    const uint32_t addr = rt_smem_u32(bar);
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (++spins > RT_SPIN_LIMIT) asm volatile("trap;");      // never hang the GPU
    }
}
__device__ __forceinline__ uint32_t rt_elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
    return pred;
}
__device__ __forceinline__ void rt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void rt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void rt_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rt_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void rt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rt_tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(rt_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(rt_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void rt_tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(rt_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(rt_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void rt_tma_store_3d(const void* src, const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(rt_smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void rt_tma_store_2d(const void* src, const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(rt_smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row atoms of 1 KB)
__device__ __forceinline__ uint64_t rt_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t rt_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define RT_TMEM_LD16(taddr, r)                                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                             \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                      \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                       \
                 : "r"(taddr))

struct ReconTcParams {
    int ndim;
    int64_t shape[TLB200_MAX_NDIM];
    int64_t rs[TLB200_MAX_NDIM], cs[TLB200_MAX_NDIM];
    const float* f[TLB200_MAX_NDIM];
    int64_t I, C;
    int R;
    const float* w;
    const float* x;
    const float* mask;
    float* out;
    double* partial;           // MODE 1: [gridDim.x][2]
    int64_t row_tiles, col_tiles, tiles_per_cta;
    int use_tma;               // 1: epilogue through TMA (needs C % 4 == 0 and 16-byte aligned out / x / mask)
};

// MODE 0: out = rec.   MODE 1: out = x*mask + rec*(1-mask) + the two norms.   MODE 2: out = rec * mask.
// shared-memory plan per (CHUNKS, MODE, W): see the notes on NB below
// W = 128-byte lines per row and TMA box (1 or 2): rows of the tensor lie megabytes apart, and DRAM serves two adjacent
// lines per row markedly better than one (measured for the tensor stream of tc_stream.cu: 6.8 against 4.5 TB/s)
template <int CHUNKS, int MODE, int W>
struct RtCfg {
    static constexpr int A_BYTES = CHUNKS * 16384;          // per part (hi or lo): [chunk][128 rows][128 B]
    static constexpr int B_STAGE = 32768;                   // ONE 32-wide chunk of the tile: [hi 128 rows | lo 128 rows][128 B]
    // Khatri-Rao stages (one chunk each; a rank 33..64 tile takes two steps through them).  Rank <= 32: two.  Rank
    // 33..64: the plain reconstruction (4 bytes per element, the shortest tile time) keeps two whole tiles = four
    // stages and one-line boxes (3.5 TB/s; two stages + two-line boxes: 2.9); the imputation variants (12 bytes per
    // element: producer + MMA of both chunks still take less than the tile's HBM time) keep ONE so that their slots
    // can hold two-line boxes (4.8 TB/s; one-line boxes: 4.2).
    static constexpr int NB = CHUNKS == 1 ? 2 : (MODE == 0 ? 4 : 1);
    static constexpr int CPS = (CHUNKS == 2 && MODE == 0) ? 2 : 1;      // chunks handed over per barrier step
    static constexpr int NST = NB / CPS;                                // step-stages
    static constexpr int BOX = 16384 * W;                   // one box: [128 rows][W lines][128 B]
    static constexpr int SLOT = MODE == 0 ? BOX : 2 * BOX;  // [out / x box][mask box]
    static constexpr int SLOT_BYTES = CHUNKS == 1 ? 128 * 1024 : (MODE == 0 ? 32 * 1024 : 128 * 1024);
    static constexpr int NS = SLOT_BYTES / SLOT > 4 ? 4 : SLOT_BYTES / SLOT;
    static_assert(NS >= 2, "at least two slots");
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = 2 * A_BYTES;
    static constexpr int OFF_SLOT = OFF_B + NB * B_STAGE;
    static constexpr int OFF_BAR = OFF_SLOT + NS * SLOT;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static_assert(NS * SLOT >= 4 * 32 * RT_SLD * 4, "the non-TMA epilogue stages through the slots");
    static_assert(SMEM <= 227 * 1024, "smem budget");
};

// ND4: the tensor is 4-way (its producers carry a two-level column odometer; a separate instance because the extra
// registers cost the 2- / 3-way path 5-25 % when both live in one kernel)
template <int CHUNKS, int MODE, int W, bool ND4>
__global__ void __launch_bounds__(RT_THREADS, 1)
recon_tc_kernel(const __grid_constant__ CUtensorMap out_map, const __grid_constant__ CUtensorMap x_map,
                const __grid_constant__ CUtensorMap mask_map, const ReconTcParams p) {
    using Cfg = RtCfg<CHUNKS, MODE, W>;
    constexpr int A_BYTES = Cfg::A_BYTES, B_STAGE = Cfg::B_STAGE, NS = Cfg::NS, SLOT = Cfg::SLOT, BOX = Cfg::BOX;
    constexpr int CPS = Cfg::CPS, NST = Cfg::NST, STEPS = CHUNKS / Cfg::CPS;
    constexpr int PC = 32 * W;                  // columns per epilogue part (TMA path)
    constexpr int OFF_A = Cfg::OFF_A, OFF_B = Cfg::OFF_B, OFF_STG = Cfg::OFF_SLOT, OFF_BAR = Cfg::OFF_BAR;
    extern __shared__ unsigned char rt_smem_raw[];
    unsigned char* smem = rt_smem_raw + ((1024u - (rt_smem_u32(rt_smem_raw) & 1023u)) & 1023u);
    unsigned char* a_hi = smem + OFF_A;
    unsigned char* a_lo = a_hi + A_BYTES;
    unsigned char* b_smem = smem + OFF_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;            // 1
    uint64_t* a_empty = bars + 1;       // 1
    uint64_t* b_full = bars + 2;        // NB (<= 4)
    uint64_t* b_empty = bars + 6;       // NB
    uint64_t* d_full = bars + 10;       // 2
    uint64_t* d_empty = bars + 12;      // 2
    uint64_t* slot_full = bars + 14;    // NS (<= 4)
    uint64_t* slot_free = bars + 18;    // NS
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
    unsigned char* slots = smem + OFF_STG;
    __shared__ double red[2][4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0 && lane == 0) {
        rt_mbar_init(a_full, 256); rt_mbar_init(a_empty, 1);
        for (int i = 0; i < 4; ++i) { rt_mbar_init(&b_full[i], 256); rt_mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { rt_mbar_init(&d_full[i], 1); rt_mbar_init(&d_empty[i], 128); }
        for (int i = 0; i < 4; ++i) { rt_mbar_init(&slot_full[i], 1); rt_mbar_init(&slot_free[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&out_map)) : "memory");
        if (MODE == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&x_map)) : "memory");
        if (MODE != 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mask_map)) : "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rt_smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    rt_fence_before();
    __syncthreads();
    rt_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t total = p.row_tiles * p.col_tiles;
    const int64_t t_begin = (int64_t)blockIdx.x * p.tiles_per_cta;
    const int64_t t_end = min(total, t_begin + p.tiles_per_cta);

    if (warp == 0) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc_wide = rt_idesc(RT_M, 2 * RT_N);      // A_hi x [K_hi | K_lo] -> columns [0, 256)
        constexpr uint32_t idesc_half = rt_idesc(RT_M, RT_N);          // A_lo x K_hi          -> columns [128, 256)
        const uint32_t a_hi_addr = rt_smem_u32(a_hi), a_lo_addr = rt_smem_u32(a_lo), b_addr = rt_smem_u32(b_smem);
        int64_t cur_rt = -1;
        uint32_t a_gen = 0, n = 0;
        for (int64_t t = t_begin; t < t_end; ++t, ++n) {
            const int64_t rt = t / p.col_tiles;
            const uint32_t s = n & 1u;                       // accumulator set
            if (rt != cur_rt) { rt_mbar_wait(a_full, a_gen & 1u); cur_rt = rt; }
            if (n >= 2) rt_mbar_wait(&d_empty[s], ((n >> 1) - 1) & 1u);
            const bool last_of_row = t + 1 == t_end || (t + 1) / p.col_tiles != rt;
            const uint32_t d0 = tmem_base + s * 256;
#pragma unroll
            for (int st = 0; st < STEPS; ++st) {
                const uint32_t step = n * STEPS + st;        // Khatri-Rao stages turn over once per step (CPS chunks)
                const uint32_t sb = step % NST;
                rt_mbar_wait(&b_full[sb], (step / NST) & 1u);
                rt_fence_after();
                if (rt_elect_one()) {
#pragma unroll
                    for (int c2 = 0; c2 < CPS; ++c2) {
                        const int ch = st * CPS + c2;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t bd = rt_desc(b_addr + (sb * CPS + c2) * B_STAGE + ks * 32);
                            rt_mma_ss(d0, rt_desc(a_hi_addr + ch * 16384 + ks * 32), bd, idesc_wide, (ch | ks) ? 1u : 0u);
                            rt_mma_ss(d0 + RT_N, rt_desc(a_lo_addr + ch * 16384 + ks * 32), bd, idesc_half, 1u);
                        }
                    }
                    rt_commit(&b_empty[sb]);
                    if (st == STEPS - 1) {
                        rt_commit(&d_full[s]);
                        if (last_of_row) rt_commit(a_empty);
                    }
                }
                __syncwarp();
            }
            if (last_of_row) ++a_gen;
        }
    } else if (warp <= 8) {
        // ================= operand producers (256 threads: two warps per scheduler hide each other's latencies) =========
        // thread (rr, c): rows rr + 32 i of a tile, 16-byte chunk c (4 rank entries) — 8 threads read one 128-byte
        // factor row together (coalesced), and the chunk goes to its swizzled place with one store (one thread per
        // row of 32 entries made every warp-level load touch 32 cache lines: the kernel was LSU-bound at 0.3 TB/s).
        const int pt = tid - 32, rr = pt >> 3, c = pt & 7;
        int64_t cur_rt = -1;
        uint32_t a_gen = 0, n = 0;
        bool unit_cs = true;
        for (int m = 0; m < p.ndim; ++m) unit_cs = unit_cs && p.cs[m] == 1 && (p.rs[m] % 4) == 0 &&
                                                   (reinterpret_cast<uintptr_t>(p.f[m]) % 16) == 0;
        auto load4 = [&](const float* rowp, int64_t cs, int r0, bool vec) {      // entries r0 .. r0 + 3 of a factor row
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vec && r0 + 4 <= p.R) return __ldg(reinterpret_cast<const float4*>(rowp + r0));
            if (r0 < p.R) v.x = __ldg(rowp + (int64_t)r0 * cs);
            if (r0 + 1 < p.R) v.y = __ldg(rowp + (int64_t)(r0 + 1) * cs);
            if (r0 + 2 < p.R) v.z = __ldg(rowp + (int64_t)(r0 + 2) * cs);
            if (r0 + 3 < p.R) v.w = __ldg(rowp + (int64_t)(r0 + 3) * cs);
            return v;
        };
        auto store_chunk = [&](unsigned char* hi_tile, unsigned char* lo_tile, int row, float4 v) {
            uint4 h, l;
            h.x = __float_as_uint(v.x) & 0xFFFFE000u; h.y = __float_as_uint(v.y) & 0xFFFFE000u;
            h.z = __float_as_uint(v.z) & 0xFFFFE000u; h.w = __float_as_uint(v.w) & 0xFFFFE000u;
            l.x = __float_as_uint(v.x - __uint_as_float(h.x)); l.y = __float_as_uint(v.y - __uint_as_float(h.y));
            l.z = __float_as_uint(v.z - __uint_as_float(h.z)); l.w = __float_as_uint(v.w - __uint_as_float(h.w));
            const uint32_t off = (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(hi_tile + off) = h;
            *reinterpret_cast<uint4*>(lo_tile + off) = l;
        };
        for (int64_t t = t_begin; t < t_end; ++t, ++n) {
            const int64_t rt = t / p.col_tiles, ct = t - rt * p.col_tiles;
            if (rt != cur_rt) {
                // the previous row block's MMAs are done with the A tile
                if (a_gen > 0) rt_mbar_wait(a_empty, (a_gen - 1) & 1u);
#pragma unroll
                for (int ch = 0; ch < CHUNKS; ++ch) {
                    const int r0 = ch * 32 + 4 * c;
                    float4 wv = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (p.w) {
                        wv.x = r0 < p.R ? __ldg(p.w + r0) : 0.f; wv.y = r0 + 1 < p.R ? __ldg(p.w + r0 + 1) : 0.f;
                        wv.z = r0 + 2 < p.R ? __ldg(p.w + r0 + 2) : 0.f; wv.w = r0 + 3 < p.R ? __ldg(p.w + r0 + 3) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = rr + 32 * i;
                        const int64_t gi = rt * RT_M + row;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (gi < p.I) {
                            v = load4(p.f[0] + gi * p.rs[0], p.cs[0], r0, unit_cs);
                            v.x *= wv.x; v.y *= wv.y; v.z *= wv.z; v.w *= wv.w;
                        }
                        store_chunk(a_hi + ch * 16384, a_lo + ch * 16384, row, v);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rt_mbar_arrive(a_full);
                cur_rt = rt;
                ++a_gen;
            }
            // 2- and 3-way tensors: factor rows of this thread's four columns (32 apart): one division per tile, then an
            // odometer — once per tile, not per chunk (the producers set the pace at rank 33..64)
            const bool three = p.ndim == 3;
            const float* r1[4] = {nullptr, nullptr, nullptr, nullptr};
            const float* r2[4] = {nullptr, nullptr, nullptr, nullptr};
            bool ok[4] = {false, false, false, false};
            if (!ND4 && p.ndim <= 3) {
                const int64_t I2 = three ? p.shape[2] : 1;
                const int64_t gc0 = ct * RT_N + rr;
                int64_t j, k;
                if (p.C < (1LL << 31)) {                  // 32-bit division: a fifth of the 64-bit one's instructions
                    const uint32_t g32 = (uint32_t)gc0, i32 = (uint32_t)I2;
                    const uint32_t q = three ? g32 / i32 : g32;
                    j = q;
                    k = three ? g32 - q * i32 : 0;
                } else {
                    j = three ? gc0 / I2 : gc0;
                    k = three ? gc0 - j * I2 : 0;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ok[i] = j < p.shape[1];
                    r1[i] = p.f[1] + (ok[i] ? j : 0) * p.rs[1];
                    r2[i] = three ? p.f[2] + k * p.rs[2] : p.f[1];
                    if (three) { k += 32; while (k >= I2) { k -= I2; ++j; } } else j += 32;
                }
            }
            // 4-way tensors (ND4 instances): the same with a two-level odometer
            const float* q3[4] = {nullptr, nullptr, nullptr, nullptr};
            if constexpr (ND4) {
                const int64_t I2 = p.shape[2], I3 = p.shape[3];
                const int64_t gc0 = ct * RT_N + rr;
                int64_t j, k, l;
                if (p.C < (1LL << 31)) {
                    const uint32_t g32 = (uint32_t)gc0, i2 = (uint32_t)I2, i3 = (uint32_t)I3;
                    const uint32_t t = g32 / i3;
                    l = g32 - t * i3;
                    const uint32_t q = t / i2;
                    k = t - q * i2;
                    j = q;
                } else {
                    const int64_t t = gc0 / I3;
                    l = gc0 - t * I3;
                    j = t / I2;
                    k = t - j * I2;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ok[i] = j < p.shape[1];
                    r1[i] = p.f[1] + (ok[i] ? j : 0) * p.rs[1];
                    r2[i] = p.f[2] + k * p.rs[2];
                    q3[i] = p.f[3] + l * p.rs[3];
                    l += 32;
                    while (l >= I3) { l -= I3; ++k; }
                    while (k >= I2) { k -= I2; ++j; }
                }
            }
#pragma unroll
            for (int st = 0; st < STEPS; ++st) {
            const uint32_t step = n * STEPS + st;            // one barrier step per CPS 32-wide chunks of the tile
            const uint32_t s = step % NST;
            if (step >= (uint32_t)NST) rt_mbar_wait(&b_empty[s], ((step / NST) - 1) & 1u);
            if (ND4 || p.ndim <= 3) {
                // all factor-row loads of a chunk are issued before the first product (two columns at a time was a
                // chain of exposed L2 round trips)
#pragma unroll
                for (int c2 = 0; c2 < CPS; ++c2) {
                    const int ch = st * CPS + c2;
                    unsigned char* stage = b_smem + (s * CPS + c2) * B_STAGE;
                    const int r0 = ch * 32 + 4 * c;
                    float4 u1[4], u2[4], u3[ND4 ? 4 : 1];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        u1[i] = ok[i] ? load4(r1[i], p.cs[1], r0, unit_cs) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ND4 || three) u2[i] = load4(r2[i], p.cs[2], r0, unit_cs);
                        if constexpr (ND4) u3[i] = load4(q3[i], p.cs[3], r0, unit_cs);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 v = u1[i];      // products left to right, like the reference's khatri_rao
                        if (ND4 || three) { v.x *= u2[i].x; v.y *= u2[i].y; v.z *= u2[i].z; v.w *= u2[i].w; }
                        if constexpr (ND4) { v.x *= u3[i].x; v.y *= u3[i].y; v.z *= u3[i].z; v.w *= u3[i].w; }
                        store_chunk(stage, stage + 16384, rr + 32 * i, v);
                    }
                }
            } else {
#pragma unroll
            for (int c2 = 0; c2 < CPS; ++c2) {
            const int ch = st * CPS + c2;
            unsigned char* stage = b_smem + (s * CPS + c2) * B_STAGE;
#pragma unroll 2
            for (int i = 0; i < 4; ++i) {
                const int col = rr + 32 * i;
                const int64_t gc = ct * RT_N + col;
                // factor rows of this column (last mode fastest)
                const float* kp1 = nullptr;
                const float* kpm[TLB200_MAX_NDIM - 2];
                {
                    int64_t rem = gc < p.C ? gc : 0;
#pragma unroll
                    for (int m = TLB200_MAX_NDIM - 1; m >= 2; --m) {
                        if (m < p.ndim) {
                            const int64_t idx = rem % p.shape[m];
                            rem /= p.shape[m];
                            kpm[m - 2] = p.f[m] + idx * p.rs[m];
                        }
                    }
                    kp1 = p.f[1] + rem * p.rs[1];
                }
                {
                    const int r0 = ch * 32 + 4 * c;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gc < p.C) {
                        v = load4(kp1, p.cs[1], r0, unit_cs);
#pragma unroll
                        for (int m = 2; m < TLB200_MAX_NDIM; ++m) {
                            if (m < p.ndim) {
                                const float4 u = load4(kpm[m - 2], p.cs[m], r0, unit_cs);
                                v.x *= u.x; v.y *= u.y; v.z *= u.z; v.w *= u.w;
                            }
                        }
                    }
                    store_chunk(stage, stage + 16384, col, v);
                }
            }
            }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            rt_mbar_arrive(&b_full[s]);
            }
        }
    } else if (warp == 13) {
        // ================= TMA loader of the x / mask boxes (imputation / masked variants, TMA epilogue only) ==========
        if (MODE != 0 && p.use_tma) {
            uint32_t g = 0;                                   // global part counter of this CTA
            for (int64_t t = t_begin; t < t_end; ++t) {
                const int64_t rt = t / p.col_tiles, ct = t - rt * p.col_tiles;
                for (int part = 0; part < RT_N / PC; ++part, ++g) {
                    const uint32_t k = g % NS;
                    if (g >= (uint32_t)NS) rt_mbar_wait(&slot_free[k], ((g / NS) - 1) & 1u);
                    if (rt_elect_one()) {
                        unsigned char* slot = slots + k * SLOT;
                        rt_mbar_expect_tx(&slot_full[k], MODE == 1 ? 2 * BOX : BOX);
                        const int cx = (int)(ct * RT_N + part * PC), cy = (int)(rt * RT_M);
                        if constexpr (W == 1) {
                            if (MODE == 1) rt_tma_load_2d(slot, &x_map, &slot_full[k], cx, cy);
                            rt_tma_load_2d(slot + BOX, &mask_map, &slot_full[k], cx, cy);
                        } else {
                            if (MODE == 1) rt_tma_load_3d(slot, &x_map, &slot_full[k], 0, cx / 32, cy);
                            rt_tma_load_3d(slot + BOX, &mask_map, &slot_full[k], 0, cx / 32, cy);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (p.use_tma) {
        // ================= epilogue through TMA: TMEM -> swizzled slot -> cp.async.bulk.tensor store =================
        // A lane owns one row of the tile (that is how tcgen05.ld hands it out): 8 chunks of 16 bytes per 32-column
        // part, written at chunk ^ (row & 7) — the SWIZZLE_128B pattern, conflict-free for lane = row — so the box
        // goes out (and x / mask come in) as whole 128-byte rows without a lane ever touching global memory.
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool storer = q == 0 && lane == 0;
        double s_out = 0.0, s_res = 0.0;
        uint32_t n = 0, g = 0;
        for (int64_t t = t_begin; t < t_end; ++t, ++n) {
            const int64_t rt = t / p.col_tiles, ct = t - rt * p.col_tiles;
            const uint32_t s = n & 1u;
            rt_mbar_wait(&d_full[s], (n >> 1) & 1u);
            rt_fence_after();
            const int64_t gi = rt * RT_M + row;
#pragma unroll 1
            for (int part = 0; part < RT_N / PC; ++part, ++g) {
                const uint32_t k = g % NS;
                unsigned char* slot = slots + k * SLOT;
                if (MODE != 0) rt_mbar_wait(&slot_full[k], (g / NS) & 1u);               // x / mask boxes have landed
#pragma unroll
                for (int ko = 0; ko < W; ++ko) {                  // the W lines of this row: 32 columns each
                    const int col0 = part * PC + ko * 32;
                    uint32_t r0[32], r1[32];
                    RT_TMEM_LD16(lane_addr + s * 256 + col0, r0);                        // hi*hi
                    RT_TMEM_LD16(lane_addr + s * 256 + col0 + 16, (r0 + 16));
                    RT_TMEM_LD16(lane_addr + s * 256 + RT_N + col0, r1);                 // hi*lo + lo*hi
                    RT_TMEM_LD16(lane_addr + s * 256 + RT_N + col0 + 16, (r1 + 16));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (col0 + 32 == RT_N) {                      // every column of the set has been read
                        rt_fence_before();
                        rt_mbar_arrive(&d_empty[s]);
                    }
                    const int line = W * row + ko;                // [128 rows][W lines][128 B], chunk c at c ^ (line & 7)
                    float po[4] = {0.f, 0.f, 0.f, 0.f}, pr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t off = (uint32_t)line * 128u + (uint32_t)((c ^ (line & 7)) << 4);
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r0[4 * c + j]) + __uint_as_float(r1[4 * c + j]);
                        if constexpr (MODE == 1) {
                            const float4 xv = *reinterpret_cast<const float4*>(slot + off);
                            const float4 mv = *reinterpret_cast<const float4*>(slot + BOX + off);
                            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w};
                            // chunks beyond the tensor were zero-filled by TMA (x = mask = 0) and are clipped on the
                            // way out; they must not reach the norms (C % 4 == 0: whole chunks)
                            const bool live = gi < p.I && ct * RT_N + col0 + 4 * c < p.C;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float rec = v[j];
                                const float o = xs[j] * ms[j] + rec * (1.f - ms[j]);       // the reference's expression
                                const float d = o - rec;
                                if (live) { po[j] = fmaf(o, o, po[j]); pr[j] = fmaf(d, d, pr[j]); }
                                v[j] = o;
                            }
                        } else if constexpr (MODE == 2) {
                            const float4 mv = *reinterpret_cast<const float4*>(slot + BOX + off);
                            v[0] *= mv.x; v[1] *= mv.y; v[2] *= mv.z; v[3] *= mv.w;
                        }
                        *reinterpret_cast<float4*>(slot + off) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                    if constexpr (MODE == 1) {
                        // fp32 partial sums over 32 elements (4 independent chains), folded into doubles per line
                        s_out += (double)((po[0] + po[1]) + (po[2] + po[3]));
                        s_res += (double)((pr[0] + pr[1]) + (pr[2] + pr[3]));
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the slot is read by the async proxy next
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (storer) {
                    const int cx = (int)(ct * RT_N + part * PC), cy = (int)(rt * RT_M);
                    if constexpr (W == 1) rt_tma_store_2d(slot, &out_map, cx, cy);
                    else rt_tma_store_3d(slot, &out_map, 0, cx / 32, cy);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    // A slot is free again as soon as its store has READ it (not when the data has landed).  With two
                    // slots the loader needs this one back at once, so the storer waits for the read; with three or
                    // more it only makes sure the PREVIOUS store has been read (it almost always has) and the slot
                    // being rewritten next — last used three or more parts ago — is then known free at the barrier.
                    if constexpr (NS == 2) {
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        if (MODE != 0) rt_mbar_arrive(&slot_free[k]);
                    } else {
                        static_assert(NS >= 3, "slot reuse distance");
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        if (MODE != 0 && g >= 1) rt_mbar_arrive(&slot_free[(g - 1) % NS]);
                    }
                }
            }
        }
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all boxes written before the kernel ends
        if constexpr (MODE == 1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s_out += __shfl_xor_sync(0xffffffffu, s_out, o);
                s_res += __shfl_xor_sync(0xffffffffu, s_res, o);
            }
            if (lane == 0) { red[0][q] = s_out; red[1][q] = s_res; }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (q == 0 && lane == 0) {
                double a = 0.0, b = 0.0;
                for (int i = 0; i < 4; ++i) { a += red[0][i]; b += red[1][i]; }
                p.partial[2 * blockIdx.x] = a;
                p.partial[2 * blockIdx.x + 1] = b;
            }
        }
    } else {
        // ================= epilogue: TMEM -> shared-memory transpose -> (blend) -> global =================
        // (pointers or extents TMA cannot describe: C % 4 != 0 or a misaligned base)
        // tcgen05.ld hands a lane one ROW of the tile; written like that every store instruction would touch 32 rows
        // that lie megabytes apart.  Each warp therefore stages its 32 rows x 32 columns in shared memory and writes
        // them out 4 rows x 128 contiguous bytes per instruction (and reads x / mask the same way).
        const int q = warp & 3;                                   // TMEM lane quarter of this warp
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* stg = reinterpret_cast<float*>(smem + OFF_STG) + q * (32 * RT_SLD);
        const bool vec_ok = (p.C % 4) == 0 && (reinterpret_cast<uintptr_t>(p.out) % 16) == 0 &&
                            (MODE != 1 || (reinterpret_cast<uintptr_t>(p.x) % 16) == 0) &&
                            (MODE == 0 || (reinterpret_cast<uintptr_t>(p.mask) % 16) == 0);
        const int orow = lane >> 3, ocol = (lane & 7) * 4;        // write-out: rows orow + 4 i, columns ocol .. ocol + 3
        double s_out = 0.0, s_res = 0.0;
        uint32_t n = 0;
        for (int64_t t = t_begin; t < t_end; ++t, ++n) {
            const int64_t rt = t / p.col_tiles, ct = t - rt * p.col_tiles;
            const uint32_t s = n & 1u;
            rt_mbar_wait(&d_full[s], (n >> 1) & 1u);
            rt_fence_after();
#pragma unroll 1
            for (int part = 0; part < RT_N / RT_SC; ++part) {
#pragma unroll
                for (int c0 = 0; c0 < RT_SC; c0 += 16) {
                    uint32_t r0[16], r1[16];
                    RT_TMEM_LD16(lane_addr + s * 256 + part * RT_SC + c0, r0);              // hi*hi
                    RT_TMEM_LD16(lane_addr + s * 256 + RT_N + part * RT_SC + c0, r1);       // hi*lo + lo*hi
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int h = 0; h < 4; ++h)
                        *reinterpret_cast<float4*>(stg + lane * RT_SLD + c0 + 4 * h) =
                            make_float4(__uint_as_float(r0[4 * h]) + __uint_as_float(r1[4 * h]),
                                        __uint_as_float(r0[4 * h + 1]) + __uint_as_float(r1[4 * h + 1]),
                                        __uint_as_float(r0[4 * h + 2]) + __uint_as_float(r1[4 * h + 2]),
                                        __uint_as_float(r0[4 * h + 3]) + __uint_as_float(r1[4 * h + 3]));
                }
                if (part == RT_N / RT_SC - 1) {                   // every column of the set has been read
                    rt_fence_before();
                    rt_mbar_arrive(&d_empty[s]);
                }
                __syncwarp();
                const int64_t gc = ct * RT_N + part * RT_SC + ocol;
                // the two norms: fp32 partial sums over this thread's 32 elements of the step (4 independent chains
                // each), folded into the double accumulators once per step — a double per element was a chain of
                // dependent fp64 operations longer than the rest of the epilogue
                float po[4] = {0.f, 0.f, 0.f, 0.f}, pr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                for (int i = 0; i < 8; ++i) {
                    const int lr = orow + 4 * i;
                    const int64_t gi = rt * RT_M + q * 32 + lr;
                    if (gi >= p.I || gc >= p.C) continue;
                    const int64_t off = gi * p.C + gc;
                    const float4 rv = *reinterpret_cast<const float4*>(stg + lr * RT_SLD + ocol);
                    float v[4] = {rv.x, rv.y, rv.z, rv.w};
                    if (vec_ok && gc + 4 <= p.C) {
                        if constexpr (MODE == 1) {
                            const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + off));
                            const float4 mv = __ldg(reinterpret_cast<const float4*>(p.mask + off));
                            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float rec = v[j];
                                const float o = xs[j] * ms[j] + rec * (1.f - ms[j]);       // the reference's expression
                                const float d = o - rec;
                                po[j] = fmaf(o, o, po[j]);
                                pr[j] = fmaf(d, d, pr[j]);
                                v[j] = o;
                            }
                        } else if constexpr (MODE == 2) {
                            const float4 mv = __ldg(reinterpret_cast<const float4*>(p.mask + off));
                            v[0] *= mv.x; v[1] *= mv.y; v[2] *= mv.z; v[3] *= mv.w;
                        }
                        *reinterpret_cast<float4*>(p.out + off) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (gc + j >= p.C) continue;
                            float o = v[j];
                            if constexpr (MODE == 1) {
                                const float xs = p.x[off + j], ms = p.mask[off + j];
                                o = xs * ms + v[j] * (1.f - ms);
                                const float d = o - v[j];
                                po[j] = fmaf(o, o, po[j]);
                                pr[j] = fmaf(d, d, pr[j]);
                            } else if constexpr (MODE == 2) {
                                o = v[j] * p.mask[off + j];
                            }
                            p.out[off + j] = o;
                        }
                    }
                }
                if constexpr (MODE == 1) {
                    s_out += (double)((po[0] + po[1]) + (po[2] + po[3]));
                    s_res += (double)((pr[0] + pr[1]) + (pr[2] + pr[3]));
                }
                __syncwarp();
            }
        }
        if constexpr (MODE == 1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s_out += __shfl_xor_sync(0xffffffffu, s_out, o);
                s_res += __shfl_xor_sync(0xffffffffu, s_res, o);
            }
            if (lane == 0) { red[0][q] = s_out; red[1][q] = s_res; }
            asm volatile("bar.sync 1, 128;" ::: "memory");           // the four epilogue warps only
            if (q == 0 && lane == 0) {
                double a = 0.0, b = 0.0;
                for (int i = 0; i < 4; ++i) { a += red[0][i]; b += red[1][i]; }
                p.partial[2 * blockIdx.x] = a;
                p.partial[2 * blockIdx.x + 1] = b;
            }
        }
    }

    rt_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

template <int CHUNKS, int MODE, int W, bool ND4>
int launch_nd(const ReconTcParams& p, const CUtensorMap* maps, int grid, cudaStream_t stream) {
    // at least 120 KB: two of these CTAs must never share an SM (each allocates all 512 TMEM columns)
    constexpr int need = RtCfg<CHUNKS, MODE, W>::SMEM;
    constexpr int smem = need > 120 * 1024 ? need : 120 * 1024;
    static std::atomic<uint64_t> attr_done{0};
    if (ensure_dynamic_smem(recon_tc_kernel<CHUNKS, MODE, W, ND4>, smem, attr_done)) return TLB200_ECUDA;
    recon_tc_kernel<CHUNKS, MODE, W, ND4><<<grid, RT_THREADS, smem, stream>>>(maps[0], maps[1], maps[2], p);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <int CHUNKS, int MODE, int W>
int launch_one(const ReconTcParams& p, const CUtensorMap* maps, int grid, cudaStream_t stream) {
    return p.ndim == 4 ? launch_nd<CHUNKS, MODE, W, true>(p, maps, grid, stream)
                       : launch_nd<CHUNKS, MODE, W, false>(p, maps, grid, stream);
}

}  // namespace

bool recon_tc_supported(const int64_t* shape, int ndim, int64_t rank, int dtype) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("TLB200_RECON_SIMT"); off = (e && atoi(e) != 0) ? 1 : 0; }
    if (off || dtype != TLB200_F32 || rank > 64 || rank < 1 || ndim < 2) return false;
    int64_t C = 1;
    for (int n = 1; n < ndim; ++n) C *= shape[n];
    // small problems are launch-bound: the SIMT kernel's grid of independent tiles has the shorter prologue
    return shape[0] >= 64 && C >= 1024 && shape[0] * C >= (1LL << 20);
}

int recon_tc_grid(const int64_t* shape, int ndim) {
    int64_t C = 1;
    for (int n = 1; n < ndim; ++n) C *= shape[n];
    const int64_t tiles = ceil_div(shape[0], RT_M) * ceil_div(C, RT_N);
    return (int)(tiles < kNumSMs ? tiles : kNumSMs);
}

int recon_tc_launch(const void* const* factors, const int64_t* shape, const int64_t* frs, const int64_t* fcs, int ndim,
                    int64_t rank, const float* w, const float* x, const float* mask, float* out, double* partial,
                    cudaStream_t stream) {
    ReconTcParams p;
    p.ndim = ndim;
    p.I = shape[0];
    p.C = 1;
    for (int n = 0; n < ndim; ++n) {
        p.shape[n] = shape[n]; p.rs[n] = frs[n]; p.cs[n] = fcs[n]; p.f[n] = static_cast<const float*>(factors[n]);
        if (n >= 1) p.C *= shape[n];
    }
    for (int n = ndim; n < TLB200_MAX_NDIM; ++n) { p.shape[n] = 1; p.rs[n] = 0; p.cs[n] = 0; p.f[n] = nullptr; }
    p.R = (int)rank;
    p.w = w; p.x = x; p.mask = mask; p.out = out; p.partial = partial;
    p.row_tiles = ceil_div(p.I, RT_M);
    p.col_tiles = ceil_div(p.C, RT_N);
    const int grid = recon_tc_grid(shape, ndim);
    p.tiles_per_cta = ceil_div(p.row_tiles * p.col_tiles, grid);
    const int mode = x ? 1 : (mask ? 2 : 0);
    // the epilogue goes through TMA when the [I][C] views of out / x / mask can be described to it: 16-byte aligned
    // bases and row pitch; boxes of 32 columns x 128 rows, SWIZZLE_128B (edges are clipped / zero-filled by TMA)
    CUtensorMap maps[3];
    memset(maps, 0, sizeof(maps));
    p.use_tma = tc_available() && p.C % 4 == 0 && p.C < (1LL << 31) && p.I < (1LL << 31) &&
                reinterpret_cast<uintptr_t>(out) % 16 == 0 && (!x || reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                (!mask || reinterpret_cast<uintptr_t>(mask) % 16 == 0) && !getenv("TLB200_RECON_NO_TMA");
    // two lines per row and box where the extents allow it (C % 32 == 0)
    static int w_cap = -1;
    if (w_cap < 0) { const char* e = getenv("TLB200_RECON_LINES"); w_cap = e ? atoi(e) : 2; }
    const int W = (p.use_tma && p.C % 32 == 0 && w_cap >= 2 && !(rank > 32 && mode == 0)) ? 2 : 1;
    if (p.use_tma) {
        int st;
        const void* bases[3] = {out, x, mask};
        for (int i = 0; i < 3; ++i) {
            if (!bases[i]) continue;
            if (W == 1) {
                const uint64_t dims[2] = {(uint64_t)p.C, (uint64_t)p.I}, strides[1] = {(uint64_t)p.C * 4};
                const uint32_t box[2] = {32, 128};
                st = tc_encode_map(&maps[i], bases[i], 2, dims, strides, box, true);
            } else {
                const uint64_t dims[3] = {32, (uint64_t)p.C / 32, (uint64_t)p.I}, strides[2] = {128, (uint64_t)p.C * 4};
                const uint32_t box[3] = {32, 2, 128};
                st = tc_encode_map(&maps[i], bases[i], 3, dims, strides, box, true);
            }
            if (st) { p.use_tma = 0; break; }
        }
    }
    const int Wk = p.use_tma ? W : 1;
    if (rank <= 32) {
        if (Wk == 2) {
            if (mode == 0) return launch_one<1, 0, 2>(p, maps, grid, stream);
            if (mode == 1) return launch_one<1, 1, 2>(p, maps, grid, stream);
            return launch_one<1, 2, 2>(p, maps, grid, stream);
        }
        if (mode == 0) return launch_one<1, 0, 1>(p, maps, grid, stream);
        if (mode == 1) return launch_one<1, 1, 1>(p, maps, grid, stream);
        return launch_one<1, 2, 1>(p, maps, grid, stream);
    }
    if (Wk == 2) {
        if (mode == 1) return launch_one<2, 1, 2>(p, maps, grid, stream);
        return launch_one<2, 2, 2>(p, maps, grid, stream);
    }
    if (mode == 0) return launch_one<2, 0, 1>(p, maps, grid, stream);
    if (mode == 1) return launch_one<2, 1, 1>(p, maps, grid, stream);
    return launch_one<2, 2, 1>(p, maps, grid, stream);
}

}  // namespace tlb200
