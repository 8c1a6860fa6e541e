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
__global__ void __launch_bounds__(DM_THREADS)
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

template <int NT, bool KM>
int launch_dmma_one(const StreamGemmParams<double>& p_in, cudaStream_t stream) {
    constexpr int TN = 8 * NT;
    StreamGemmParams<double> p = p_in;
    p.m_tiles = ceil_div(p.M, DM_TM);
    const int64_t gx = p.m_tiles * p.nbatch, gy = ceil_div(p.N, TN), gz = p.nsplit;
    if (gx <= 0 || gy <= 0 || gz <= 0) return TLB200_OK;
    if (gx > 0x7fffffffLL || gy > 65535 || gz > 65535) return TLB200_EUNSUPPORTED;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
    stream_gemm_dmma_kernel<NT, KM><<<grid, DM_THREADS, 0, stream>>>(p);
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
    if (TN == 32) return a_kmajor ? launch_dmma_one<4, true>(p, s) : launch_dmma_one<4, false>(p, s);
    if (TN == 64) return a_kmajor ? launch_dmma_one<8, true>(p, s) : launch_dmma_one<8, false>(p, s);
    return TLB200_EINVAL;
}

}  // namespace tlb200
