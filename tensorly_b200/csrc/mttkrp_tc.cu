// tcgen05 MTTKRP for fp32 tensors: error-compensated 3xTF32 on the 5th-gen tensor cores.
//
// Why tensor cores for a "streaming" op: MTTKRP needs 2R flops per tensor element, i.e.
// 16-32 flop/byte at R = 32-64.  Streaming the tensor at HBM speed therefore needs
// 100-200 TFLOP/s of fp32-accurate math — more than the CUDA cores have — so the
// contraction runs as a skinny GEMM  D[128 x R] += X_tile[128 x 32] * KR_tile[32 x R]
// on tcgen05, with the accumulator in tensor memory.
//
// Why 3xTF32: kind::tf32 reads 32-bit operands and keeps 10 mantissa bits (~7e-4 relative
// error per product, far above the 1e-5 gate).  Each operand is therefore split exactly
// into hi = tf32(x) and lo = x - hi, and  x*y ~= hi_x*hi_y + lo_x*hi_y + hi_x*lo_y
// (dropped term ~2^-22).  Three MMAs per K step instead of one.
//
// Data flow per CTA (one 128-row tile of the kept mode x one K range, 10 warps):
//   warp 0      TMA producer : X tile (128 rows x 32 contraction elements, 16 KB) global -> smem ring
//   warps 2-5   convert      : smem X tile -> registers, split hi/lo, tcgen05.st both halves
//                              into a ring of TMEM A-operand buffers (the MMA reads A from TMEM,
//                              so the tile costs shared memory one write and one read only)
//   warps 6-9   KR synthesis : KR tile rows P[a,:] * Q[b,:] formed on the fly from the two small
//                              tables, split hi/lo, written K-major SWIZZLE_128B into a smem ring
//   warp 1      MMA issuer   : 12 tcgen05.mma (kind::tf32, M=128, N=R, K=8) per tile, commits
//                              release the A/B slots
//   warps 2-5   epilogue     : every `flush` tiles the TMEM accumulator is drained into fp32
//                              registers (round-to-nearest adds) so that the tensor core's own
//                              accumulation chain stays short; partials go to the split-K buffer
#include "mttkrp_tc.cuh"

#include <cuda.h>
#include <cstdlib>

namespace tlb200 {
namespace {

constexpr int KT = 32;            // contraction elements per tile (= one 128-byte swizzle row)
constexpr int TM = 128;           // rows per tile (MMA M)
constexpr int XS = 6;             // smem stages of X tiles
constexpr int AS = 4;             // TMEM stages of converted A operands (hi+lo = 64 columns each)
constexpr int BS = 3;             // smem stages of KR tiles
constexpr int X_TILE_BYTES = TM * KT * 4;
constexpr int NUM_THREADS = 320;
constexpr uint32_t SPIN_LIMIT = 1u << 22;   // x ~1 us per try_wait: seconds, then trap (never hang the GPU)

struct TcParams {
    const float* P;       // [A][RP] or null
    const float* Q;       // [B][RP]
    float* partial;       // [splits][J][RP]
    int64_t J, A, B;
    int64_t chunks_per_a, total_chunks, chunks_per_split;
    int j_tiles;
    int flush;            // tiles per TMEM accumulation group
    int x_jmajor;         // 1: last mode (tile arrives as [32 k][128 j]); 0: [128 j][32 k] swizzled
};

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
        if (++spins > SPIN_LIMIT) asm volatile("trap;");   // a lost arrival must not hang the GPU
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
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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

// ---- shared-memory carve-up (offsets from a 1 KB-aligned base) -------------------------
template <int RP>
struct Smem {
    static constexpr int B_TILE = RP * 128;                 // one hi or lo KR tile (RP rows x 128 B)
    static constexpr int STAGE_LD = RP;                      // staging rows are rotated, not padded
    static constexpr int OFF_X = 0;
    static constexpr int OFF_B = OFF_X + XS * X_TILE_BYTES;
    static constexpr int OFF_STAGE = OFF_B + BS * 2 * B_TILE;
    static constexpr int OFF_BAR = OFF_STAGE + 2 * KT * STAGE_LD * 4;
    static constexpr int NUM_BARS = 2 * XS + 2 * AS + 2 * BS + 4;
    static constexpr int TOTAL = OFF_BAR + NUM_BARS * 8 + 16;
};

template <int RP>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mttkrp_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
    using L = Smem<RP>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* x_smem = smem + L::OFF_X;
    unsigned char* b_smem = smem + L::OFF_B;
    float* stage = reinterpret_cast<float*>(smem + L::OFF_STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
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
    const int jt = blockIdx.x % p.j_tiles;
    const int64_t split = blockIdx.x / p.j_tiles;
    const int64_t c_begin = split * p.chunks_per_split;
    const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_split);
    const int n_chunks = (int)(c_end - c_begin);
    const int64_t j0 = (int64_t)jt * TM;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 128); }
        for (int i = 0; i < AS; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < BS; ++i) { mbar_init(&b_full[i], 128); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: D0 [0,RP)  D1 [RP,2RP)  A stage t: hi [2RP+64t, +32)  lo [2RP+64t+32, +32)
    const uint32_t a_col0 = 2 * RP;
    const int FL = p.flush;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int i = 0; i < n_chunks; ++i) {
                const int s = i % XS;
                mbar_wait(&x_empty[s], ((i / XS) & 1) ^ 1);
                const int64_t c = c_begin + i;
                const int64_t a = c / p.chunks_per_a;
                const int b0 = (int)((c - a * p.chunks_per_a) * KT);
                mbar_expect_tx(&x_full[s], X_TILE_BYTES);
                if (p.x_jmajor) tma_load_3d(x_smem + s * X_TILE_BYTES, &tmap, &x_full[s], (int)j0, b0, (int)a);
                else            tma_load_3d(x_smem + s * X_TILE_BYTES, &tmap, &x_full[s], b0, (int)j0, (int)a);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(TM, RP);
            for (int i = 0; i < n_chunks; ++i) {
                const int t = i % AS, u = i % BS;
                const int g = i / FL, buf = g & 1;
                const bool first = (i % FL) == 0;
                if (first && g >= 2) mbar_wait(&d_empty[buf], ((g >> 1) - 1) & 1);
                mbar_wait(&a_full[t], (i / AS) & 1);
                mbar_wait(&b_full[u], (i / BS) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * RP;
                const uint32_t a_hi = tmem_base + a_col0 + t * 64;
                const uint32_t a_lo = a_hi + 32;
                const uint32_t bhi = smem_u32(b_smem + (u * 2 + 0) * L::B_TILE);
                const uint32_t blo = smem_u32(b_smem + (u * 2 + 1) * L::B_TILE);
#pragma unroll
                for (int ks = 0; ks < KT / 8; ++ks) {
                    const uint64_t dbh = desc_kmajor_sw128(bhi + ks * 32);
                    const uint64_t dbl = desc_kmajor_sw128(blo + ks * 32);
                    mma_ts_tf32(d_tmem, a_hi + ks * 8, dbh, idesc, (first && ks == 0) ? 0u : 1u);
                    mma_ts_tf32(d_tmem, a_lo + ks * 8, dbh, idesc, 1u);
                    mma_ts_tf32(d_tmem, a_hi + ks * 8, dbl, idesc, 1u);
                }
                tc_commit(&a_empty[t]);
                tc_commit(&b_empty[u]);
                if (((i + 1) % FL) == 0 || i == n_chunks - 1) tc_commit(&d_full[buf]);
            }
        }
    } else if (warp < 6) {
        // ================= convert + epilogue (4 warps, one TMEM lane quarter each) =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        float acc[RP];
#pragma unroll
        for (int c = 0; c < RP; ++c) acc[c] = 0.f;

        auto flush_group = [&](int g) {
            const int buf = g & 1;
            mbar_wait(&d_full[buf], (g >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < RP; c0 += 32) {
                uint32_t r[32];
                TLB_TMEM_LD32(lane_addr + buf * RP + c0, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c0 + c] += __uint_as_float(r[c]);
            }
            tc_fence_before();
            mbar_arrive(&d_empty[buf]);
        };

        for (int i = 0; i < n_chunks; ++i) {
            if ((i % FL) == 0 && i >= 2 * FL) flush_group(i / FL - 2);
            const int s = i % XS;
            mbar_wait(&x_full[s], (i / XS) & 1);
            uint32_t v[32];
            const unsigned char* xt = x_smem + s * X_TILE_BYTES;
            if (p.x_jmajor) {
                const float* xr = reinterpret_cast<const float*>(xt) + row;     // tile is [32 k][128 j]
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(xr[k * TM]);
            } else {
                const unsigned char* xr = xt + row * 128;                       // tile is [128 j][32 k], 128B-swizzled
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 w = *reinterpret_cast<const uint4*>(xr + ((c ^ (row & 7)) << 4));
                    v[4 * c + 0] = w.x; v[4 * c + 1] = w.y; v[4 * c + 2] = w.z; v[4 * c + 3] = w.w;
                }
            }
            mbar_arrive(&x_empty[s]);
            const int t = i % AS;
            mbar_wait(&a_empty[t], ((i / AS) & 1) ^ 1);
            tc_fence_after();
            uint32_t hi[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) hi[k] = v[k] & 0xFFFFE000u;               // exact tf32 part
            TLB_TMEM_ST32(lane_addr + a_col0 + t * 64, hi);
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) - __uint_as_float(hi[k]));  // exact remainder
            TLB_TMEM_ST32(lane_addr + a_col0 + t * 64 + 32, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&a_full[t]);
        }
        const int ngroups = (n_chunks + FL - 1) / FL;
        for (int g = max(0, ngroups - 2); g < ngroups; ++g) flush_group(g);
        const int64_t gj = j0 + row;
        if (gj < p.J && n_chunks > 0) {
            float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)split * p.J + gj) * RP);
#pragma unroll
            for (int c = 0; c < RP / 4; ++c) dst[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        } else if (gj < p.J) {
            float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)split * p.J + gj) * RP);
            for (int c = 0; c < RP / 4; ++c) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        // ================= KR synthesis (4 warps) =================
        const int kt = tid - 192;             // 0..127
        const int kw = kt >> 5;               // warp within the group
        constexpr int F4_PER_THREAD = (KT * RP / 4) / 128;   // float4 of the Q chunk per thread
        constexpr int F4_PER_ROW = RP / 4;
        float4 qreg[F4_PER_THREAD];
        auto load_q = [&](int i) {
            const int64_t c = c_begin + i;
            const int64_t a = c / p.chunks_per_a;
            const int64_t b0 = (c - a * p.chunks_per_a) * KT;
#pragma unroll
            for (int f = 0; f < F4_PER_THREAD; ++f) {
                const int idx = kt + f * 128;
                const int k = idx / F4_PER_ROW;
                qreg[f] = (b0 + k < p.B) ? __ldg(reinterpret_cast<const float4*>(p.Q + (b0 + k) * RP) + (idx % F4_PER_ROW))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (n_chunks > 0) load_q(0);
        for (int i = 0; i < n_chunks; ++i) {
            float* st = stage + (i & 1) * KT * L::STAGE_LD;
            // rotated staging: element (k, r) lives at st[k*RP + (r + k) % RP] so that both the
            // row-wise writes and the column-wise reads below are (nearly) conflict-free
#pragma unroll
            for (int f = 0; f < F4_PER_THREAD; ++f) {
                const int idx = kt + f * 128;
                const int k = idx / F4_PER_ROW, r = (idx % F4_PER_ROW) * 4;
                float* rowp = st + k * L::STAGE_LD;
                rowp[(r + 0 + k) & (RP - 1)] = qreg[f].x;
                rowp[(r + 1 + k) & (RP - 1)] = qreg[f].y;
                rowp[(r + 2 + k) & (RP - 1)] = qreg[f].z;
                rowp[(r + 3 + k) & (RP - 1)] = qreg[f].w;
            }
            const int64_t c = c_begin + i;
            const int64_t a = c / p.chunks_per_a;
            if (i + 1 < n_chunks) load_q(i + 1);        // prefetch the next Q chunk from L2
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int u = i % BS;
            mbar_wait(&b_empty[u], ((i / BS) & 1) ^ 1);
            unsigned char* bhi = b_smem + (u * 2 + 0) * L::B_TILE;
            unsigned char* blo = b_smem + (u * 2 + 1) * L::B_TILE;
            const float* prow = p.P ? p.P + a * RP : nullptr;
#pragma unroll 4
            for (int r = kw; r < RP; r += 4) {
                const float qv = st[lane * L::STAGE_LD + ((r + lane) & (RP - 1))];   // lane = k
                const float kr = prow ? __fmul_rn(__ldg(prow + r), qv) : qv;
                const uint32_t h = __float_as_uint(kr) & 0xFFFFE000u;
                const float lo = kr - __uint_as_float(h);
                const uint32_t off = r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4;
                *reinterpret_cast<uint32_t*>(bhi + off) = h;
                *reinterpret_cast<float*>(blo + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core reads
            mbar_arrive(&b_full[u]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
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

int flush_period() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TLB200_TC_FLUSH");
        v = e ? atoi(e) : 4;
        if (v < AS) v = AS;     // the deferred drain assumes a group is at least as long as the A ring
    }
    return v;
}

template <int RP>
int launch_rp(const CUtensorMap& tmap, const TcParams& p, int64_t ctas, cudaStream_t stream) {
    const int smem = Smem<RP>::TOTAL + 1024;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(mttkrp_tc_kernel<RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return TLB200_ECUDA;
        attr = true;
    }
    mttkrp_tc_kernel<RP><<<(unsigned)ctas, NUM_THREADS, smem, stream>>>(tmap, p);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

}  // namespace

bool mttkrp_tc_supported(const tlb200_mttkrp_plan_t& pl, int64_t rank, int dtype) {
    if (dtype != TLB200_F32 || rank > 64) return false;
    if (getenv("TLB200_DISABLE_TC")) return false;
    // TMA: global strides must be multiples of 16 bytes, extents must fit 32 bits
    if (pl.sb == 1) { if ((pl.sj % 4) || (pl.sa % 4)) return false; }
    else            { if ((pl.sb % 4) || (pl.sa % 4)) return false; }
    if (pl.A >= (1LL << 31) || pl.B >= (1LL << 31) || pl.J >= (1LL << 31)) return false;
    // tiny problems are launch-bound: the SIMT kernel has the shorter prologue
    if (pl.A * ceil_div(pl.B, KT) < 8 || pl.J < 32) return false;
    return get_encode_fn() != nullptr;
}

void mttkrp_tc_fill_plan(tlb200_mttkrp_plan_t* pl, int64_t rank) {
    pl->rank_padded = rank <= 32 ? 32 : 64;
    const int64_t j_tiles = ceil_div(pl->J, TM);
    const int64_t total = pl->A * ceil_div(pl->B, KT);
    int64_t splits = j_tiles >= kNumSMs ? 1 : kNumSMs / j_tiles;
    if (splits > total / 8) splits = total / 8;
    if (splits < 1) splits = 1;
    const int64_t per = ceil_div(total, splits);
    pl->splits = ceil_div(total, per);
}

size_t mttkrp_tc_extra_workspace(const tlb200_mttkrp_plan_t&) { return 0; }

int mttkrp_tc_launch(const float* x, const tlb200_mttkrp_plan_t& pl, int64_t /*rank*/, const float* P, const float* Q,
                     float* partial, void* /*extra_ws*/, cudaStream_t stream) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return TLB200_EUNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(x) % 16) return TLB200_EUNSUPPORTED;
    CUtensorMap tmap;
    cuuint64_t dims[3], strides[2];
    cuuint32_t box[3], estr[3] = {1, 1, 1};
    CUtensorMapSwizzle swz;
    const bool jmajor = pl.sb != 1;
    if (!jmajor) {            // [a][j][b], b contiguous: tile = 128 j-rows x 32 b, 128B-swizzled rows
        dims[0] = (cuuint64_t)pl.B; dims[1] = (cuuint64_t)pl.J; dims[2] = (cuuint64_t)pl.A;
        strides[0] = (cuuint64_t)pl.sj * 4; strides[1] = (cuuint64_t)pl.sa * 4;
        box[0] = KT; box[1] = TM; box[2] = 1;
        swz = CU_TENSOR_MAP_SWIZZLE_128B;
    } else {                  // [a][b][j], j contiguous: tile = 32 b-rows x 128 j, plain layout
        dims[0] = (cuuint64_t)pl.J; dims[1] = (cuuint64_t)pl.B; dims[2] = (cuuint64_t)pl.A;
        strides[0] = (cuuint64_t)pl.sb * 4; strides[1] = (cuuint64_t)pl.sa * 4;
        box[0] = TM; box[1] = KT; box[2] = 1;
        swz = CU_TENSOR_MAP_SWIZZLE_NONE;
    }
    if (pl.A == 1) strides[1] = strides[0] * dims[1];   // unused dim: any valid multiple of 16
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return TLB200_ECUDA;

    TcParams p;
    p.P = P; p.Q = Q; p.partial = partial;
    p.J = pl.J; p.A = pl.A; p.B = pl.B;
    p.chunks_per_a = ceil_div(pl.B, KT);
    p.total_chunks = pl.A * p.chunks_per_a;
    p.chunks_per_split = ceil_div(p.total_chunks, pl.splits);
    p.j_tiles = (int)ceil_div(pl.J, TM);
    p.flush = flush_period();
    p.x_jmajor = jmajor ? 1 : 0;
    const int64_t ctas = (int64_t)p.j_tiles * pl.splits;
    if (pl.rank_padded == 32) return launch_rp<32>(tmap, p, ctas, stream);
    return launch_rp<64>(tmap, p, ctas, stream);
}

}  // namespace tlb200
