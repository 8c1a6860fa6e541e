// fp64 "stream GEMM" on the double-precision tensor cores (DMMA.8x8x4 via mma.sync.m8n8k4.f64) — the float64 path of
// MTTKRP and TTM (north star: "the FP64 path where float64 is requested").
//
// Same contract as the SIMT kernel in stream_gemm.cuh (StreamGemmParams<double>: A = a strided, possibly batched
// view of the tensor streamed from HBM once; B = P[a, :] * Q[b, :] formed on the fly, or a factor matrix;
// deterministic split-K partials).  Why a second fp64 kernel: on B200 DMMA and DFMA have the same peak
// (64 FMA/clk/SM, probes/dmma_rate.cu), but the SIMT kernel needs 6 LDS.128 per 32 DFMA per thread and reaches
// 29 % of it — shared-memory bound.  A DMMA warp tile of 32 x (8 NT) takes 4 + NT 8-byte fragment loads per lane
// for 4 NT DMMAs (256 FMA each): 6-8x less shared-memory traffic per FMA, so the tensor pipe, not the LDS
// pipe, is the limit.  At rank 32 the fp64 MTTKRP is 8 flop/B: 37 TFLOP/s needs 4.6 TB/s of HBM — both walls close.
//
// CTA = 4 warps, tile 128 (m) x {32, 64} (n) x 16 (k); warp w owns rows [32 w, 32 w + 32).  Fragment layouts
// (PTX ISA, mma.m8n8k4.f64): A[row = lane / 4][k = lane % 4], B[k = lane % 4][n = lane / 4],
// C[row = lane / 4][n = 2 (lane % 4) + {0, 1}].  Shared-memory leading dimensions are chosen so that a half-warp's
// 16 fragment loads hit 32 distinct banks (LD x 2 words = 8 mod 32).
#include "stream_gemm.cuh"

#include <cstdlib>

namespace tlb200 {
namespace {

constexpr int DM_TM = 128, DM_KT = 16, DM_THREADS = 128;
constexpr int DM_LDA_K = 20;      // As[m][k]  (A k-contiguous):  20 doubles = 40 words = 8 mod 32
constexpr int DM_LDA_M = 132;     // As[k][m]  (A m-contiguous): 132 doubles = 264 words = 8 mod 32

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NT, bool A_KMAJOR>          // NT = 8-column tiles per warp (4: TN = 32, 8: TN = 64)
__global__ void __launch_bounds__(DM_THREADS, NT == 4 ? 4 : 2)      // NT = 4: 128 registers, 4 CTAs per SM = the plan's grid of 4 x 148
stream_gemm_dmma_kernel(const StreamGemmParams<double> p) {
    constexpr int TN = 8 * NT;
    constexpr int LDB = TN + 4;           // (TN + 4) * 2 words = 8 mod 32 for TN = 32 and 64
    constexpr int A_PER_THREAD = DM_TM * DM_KT / DM_THREADS;      // 16
    constexpr int B_PER_THREAD = DM_KT * TN / DM_THREADS;         // 4 or 8
    __shared__ __align__(16) double As[A_KMAJOR ? DM_TM * DM_LDA_K : DM_KT * DM_LDA_M];
    __shared__ __align__(16) double Bs[DM_KT * LDB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int64_t batch = blockIdx.x / p.m_tiles;
    const int64_t m0 = ((int64_t)blockIdx.x - batch * p.m_tiles) * DM_TM;
    const int64_t n0 = (int64_t)blockIdx.y * TN;
    const int64_t split = blockIdx.z;
    const int64_t c_begin = split * p.chunks_per_split;
    const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_split);
    const double* __restrict__ X = p.X + batch * p.sXbatch;

    double acc[4][NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    double a_reg[A_PER_THREAD], b_reg[B_PER_THREAD];
    auto load_chunk = [&](int64_t c) {
        const int64_t a = c / p.chunks_per_a;
        const int64_t b0 = (c - a * p.chunks_per_a) * DM_KT;
        const double* __restrict__ xa = X + a * p.sXa;
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; ++i) {
            const int e = tid + i * DM_THREADS;
            int m, k;
            if (A_KMAJOR) { k = e % DM_KT; m = e / DM_KT; } else { m = e % DM_TM; k = e / DM_TM; }
            const int64_t gm = m0 + m, gb = b0 + k;
            a_reg[i] = (gm < p.M && gb < p.KB) ? __ldg(xa + gm * p.sXm + gb * p.sXb) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < B_PER_THREAD; ++i) {
            const int e = tid + i * DM_THREADS;
            const int n = e % TN, k = e / TN;
            const int64_t gn = n0 + n, gb = b0 + k;
            double v = 0.0;
            if (gn < p.N && gb < p.KB) {
                v = __ldg(p.Q + gb * p.sQb + gn * p.sQn);
                if (p.P) v *= __ldg(p.P + a * p.ldP + gn);
            }
            b_reg[i] = v;
        }
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; ++i) {
            const int e = tid + i * DM_THREADS;
            if (A_KMAJOR) As[(e / DM_KT) * DM_LDA_K + (e % DM_KT)] = a_reg[i];
            else As[(e / DM_TM) * DM_LDA_M + (e % DM_TM)] = a_reg[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER_THREAD; ++i) {
            const int e = tid + i * DM_THREADS;
            Bs[(e / TN) * LDB + (e % TN)] = b_reg[i];
        }
    };

    if (c_begin < c_end) load_chunk(c_begin);
    for (int64_t c = c_begin; c < c_end; ++c) {
        __syncthreads();          // previous chunk fully consumed
        store_chunk();
        __syncthreads();
        if (c + 1 < c_end) load_chunk(c + 1);   // overlaps with the DMMAs below
#pragma unroll
        for (int k4 = 0; k4 < DM_KT / 4; ++k4) {
            double af[4], bf[NT];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = warp * 32 + i * 8 + g, k = k4 * 4 + q;
                af[i] = A_KMAJOR ? As[m * DM_LDA_K + k] : As[k * DM_LDA_M + m];
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = Bs[(k4 * 4 + q) * LDB + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }

    double* __restrict__ C = p.C + batch * p.sCbatch + split * p.sCsplit;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + warp * 32 + i * 8 + g;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int64_t gn = n0 + j * 8 + 2 * q;
            if (gn < p.N) C[gm * p.sCm + gn * p.sCn] = acc[i][j][0];
            if (gn + 1 < p.N) C[gm * p.sCm + (gn + 1) * p.sCn] = acc[i][j][1];
        }
    }
}


// ---- cp.async variant ---------------------------------------------------------------------------------------------
// The register-staged kernel above spends ~300 integer / LDG / STS instructions per thread and K chunk on moving the
// A tile (one element at a time, 64-bit index math each) for 64 DMMAs per warp: the DMMA pipe was 42-50 % active (ncu)
// and 160 registers allowed 3 CTAs per SM against a grid planned for 4.  Here the A tile goes global -> shared with
// 16-byte cp.async (8 per thread and chunk, zero fill at the edges) into a double buffer: no staging registers, one
// barrier per chunk, 4 CTAs per SM.  Needs 16-byte aligned pairs: even strides / offsets (checked on the host).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}

template <int NT, bool A_KMAJOR>
__global__ void __launch_bounds__(DM_THREADS, NT == 4 ? 4 : 2)
stream_gemm_dmma_async_kernel(const StreamGemmParams<double> p) {
    constexpr int TN = 8 * NT;
    constexpr int LDB = TN + 4;
    constexpr int A_STAGE = A_KMAJOR ? DM_TM * DM_LDA_K : DM_KT * DM_LDA_M;      // doubles
    constexpr int B_STAGE = DM_KT * LDB;
    constexpr int B_PER_THREAD = DM_KT * TN / DM_THREADS;
    extern __shared__ __align__(16) unsigned char dm_smem[];
    double* As = reinterpret_cast<double*>(dm_smem);                 // [2][A_STAGE]
    double* Bs = As + 2 * A_STAGE;                                   // [2][B_STAGE]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    const int64_t batch = blockIdx.x / p.m_tiles;
    const int64_t m0 = ((int64_t)blockIdx.x - batch * p.m_tiles) * DM_TM;
    const int64_t n0 = (int64_t)blockIdx.y * TN;
    const int64_t split = blockIdx.z;
    const int64_t c_begin = split * p.chunks_per_split;
    const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_split);
    const double* __restrict__ X = p.X + batch * p.sXbatch;

    double acc[4][NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // this thread's 8 pieces of 2 doubles: K-major piece e = (row e / 8, k pair e % 8); M-major (k e / 64, m pair e % 64)
    auto issue_a = [&](int64_t c, int stage) {
        const int64_t a = c / p.chunks_per_a;
        const int64_t b0 = (c - a * p.chunks_per_a) * DM_KT;
        const double* __restrict__ xa = X + a * p.sXa;
        double* dst = As + stage * A_STAGE;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * DM_THREADS;
            if (A_KMAJOR) {
                const int m = e >> 3, k = (e & 7) * 2;
                const int64_t gm = m0 + m, gb = b0 + k;
                const int bytes = gm < p.M ? (gb + 1 < p.KB ? 16 : (gb < p.KB ? 8 : 0)) : 0;
                cp_async16(dst + m * DM_LDA_K + k, bytes ? xa + gm * p.sXm + gb : X, bytes);
            } else {
                const int k = e >> 6, m = (e & 63) * 2;
                const int64_t gm = m0 + m, gb = b0 + k;
                const int bytes = gb < p.KB ? (gm + 1 < p.M ? 16 : (gm < p.M ? 8 : 0)) : 0;
                cp_async16(dst + k * DM_LDA_M + m, bytes ? xa + gb * p.sXb + gm : X, bytes);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double b_reg[B_PER_THREAD];
    auto load_b = [&](int64_t c) {
        const int64_t a = c / p.chunks_per_a;
        const int64_t b0 = (c - a * p.chunks_per_a) * DM_KT;
#pragma unroll
        for (int i = 0; i < B_PER_THREAD; ++i) {
            const int e = tid + i * DM_THREADS;
            const int n = e % TN, k = e / TN;
            const int64_t gn = n0 + n, gb = b0 + k;
            double v = 0.0;
            if (gn < p.N && gb < p.KB) {
                v = __ldg(p.Q + gb * p.sQb + gn * p.sQn);
                if (p.P) v *= __ldg(p.P + a * p.ldP + gn);
            }
            b_reg[i] = v;
        }
    };
    auto store_b = [&](int stage) {
        double* dst = Bs + stage * B_STAGE;
#pragma unroll
        for (int i = 0; i < B_PER_THREAD; ++i) {
            const int e = tid + i * DM_THREADS;
            dst[(e / TN) * LDB + (e % TN)] = b_reg[i];
        }
    };

    if (c_begin < c_end) {
        issue_a(c_begin, 0);
        load_b(c_begin);
        store_b(0);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int stage = 0;
    for (int64_t c = c_begin; c < c_end; ++c) {
        const bool more = c + 1 < c_end;
        if (more) { issue_a(c + 1, stage ^ 1); load_b(c + 1); }      // both in flight under the DMMAs below
        const double* as = As + stage * A_STAGE;
        const double* bs = Bs + stage * B_STAGE;
#pragma unroll
        for (int k4 = 0; k4 < DM_KT / 4; ++k4) {
            double af[4], bf[NT];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = warp * 32 + i * 8 + g, k = k4 * 4 + q;
                af[i] = A_KMAJOR ? as[m * DM_LDA_K + k] : as[k * DM_LDA_M + m];
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = bs[(k4 * 4 + q) * LDB + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (more) {
            store_b(stage ^ 1);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();          // stage^1 is complete and visible; everyone is done reading `stage`
        stage ^= 1;
    }

    double* __restrict__ C = p.C + batch * p.sCbatch + split * p.sCsplit;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + warp * 32 + i * 8 + g;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int64_t gn = n0 + j * 8 + 2 * q;
            if (gn < p.N) C[gm * p.sCm + gn * p.sCn] = acc[i][j][0];
            if (gn + 1 < p.N) C[gm * p.sCm + (gn + 1) * p.sCn] = acc[i][j][1];
        }
    }
}

// 16-byte pairs must be aligned: base pointer, and every stride that moves a pair's start, even
template <bool KM>
bool dmma_async_ok(const StreamGemmParams<double>& p) {
    if (const char* e = getenv("TLB200_FP64_NO_ASYNC")) { if (atoi(e) != 0) return false; }
    if (reinterpret_cast<uintptr_t>(p.X) % 16) return false;
    if ((p.sXa | p.sXbatch) & 1) return false;
    if (KM) return p.sXb == 1 && (p.sXm & 1) == 0;
    return p.sXm == 1 && (p.sXb & 1) == 0;
}

template <int NT, bool KM>
int launch_dmma_one(const StreamGemmParams<double>& p_in, cudaStream_t stream) {
    constexpr int TN = 8 * NT;
    StreamGemmParams<double> p = p_in;
    p.m_tiles = ceil_div(p.M, DM_TM);
    const int64_t gx = p.m_tiles * p.nbatch, gy = ceil_div(p.N, TN), gz = p.nsplit;
    if (gx <= 0 || gy <= 0 || gz <= 0) return TLB200_OK;
    if (gx > 0x7fffffffLL || gy > 65535 || gz > 65535) return TLB200_EUNSUPPORTED;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
    if (dmma_async_ok<KM>(p)) {
        constexpr int smem = 2 * ((KM ? DM_TM * DM_LDA_K : DM_KT * DM_LDA_M) + DM_KT * (TN + 4)) * (int)sizeof(double);
        static std::atomic<uint64_t> attr_done{0};
        if (ensure_dynamic_smem(stream_gemm_dmma_async_kernel<NT, KM>, smem, attr_done)) return TLB200_ECUDA;
        stream_gemm_dmma_async_kernel<NT, KM><<<grid, DM_THREADS, smem, stream>>>(p);
    } else {
        stream_gemm_dmma_kernel<NT, KM><<<grid, DM_THREADS, 0, stream>>>(p);
    }
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

}  // namespace

bool stream_gemm_dmma_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TLB200_FP64_SIMT"); on = (e && atoi(e) != 0) ? 0 : 1; }
    return on == 1;
}

// TN in {32, 64}: the caller planned its column blocks / padding with this value (stream_gemm_dmma_tn)
int launch_stream_gemm_dmma(const StreamGemmParams<double>& p, int TN, bool a_kmajor, cudaStream_t s) {
    set_last_path("dmma");
    if (TN == 32) return a_kmajor ? launch_dmma_one<4, true>(p, s) : launch_dmma_one<4, false>(p, s);
    if (TN == 64) return a_kmajor ? launch_dmma_one<8, true>(p, s) : launch_dmma_one<8, false>(p, s);
    return TLB200_EINVAL;
}

}  // namespace tlb200
