// SIMT "stream GEMM": C[m, n] = sum_k A(m, k) * B(k, n) where A is a (possibly strided,
// possibly batched) view of a large tensor that is streamed from HBM exactly once and B
// is a small operand synthesised on the fly in shared memory:
//
//   MTTKRP : A(m,(a,b)) = X[a*sXa + m*sXm + b*sXb],  B((a,b), n) = P[a, n] * Q[b, n]
//            (the Khatri-Rao rows, formed tile by tile, never materialised)
//   TTM    : A(m, b)    = X[batch*sXbatch + m*sXm + b*sXb],  B(b, n) = Mat[n, b]
//
// This is the general path: any dtype (fp32/fp64), any extents, any mode.  fp32 problems
// whose shape allows it are routed to the tcgen05 kernels instead (mttkrp_tc.cu).
// CTA tile TM x TN x KT with TM = 16*TJ, TN = 8*TR, 128 threads, a TJ x TR register
// micro-tile per thread, global->register prefetch of the next K chunk overlapping the
// FMAs of the current one, split-K over gridDim.z with partials reduced by a second,
// deterministic pass.
#pragma once
#include "common.cuh"

namespace tlb200 {

template <typename T>
struct StreamGemmParams {
    // streamed operand
    const T* X;
    int64_t M;            // extent of the streamed output dim
    int64_t KA, KB;       // contraction extent = KA (outer, "a") x KB (inner, "b")
    int64_t sXm, sXa, sXb;
    int64_t sXbatch;
    // small operand: B((a,b), n) = (P ? P[a*ldP + n] : 1) * Q[b*sQb + n*sQn]
    const T* P;
    int64_t ldP;
    const T* Q;
    int64_t sQb, sQn;
    int64_t N;            // valid output columns
    // output: C[batch*sCbatch + split*sCsplit + m*sCm + n*sCn]
    T* C;
    int64_t sCm, sCn, sCbatch, sCsplit;
    int64_t nbatch;
    int64_t m_tiles;          // ceil(M / TM); blockIdx.x = m_tile + m_tiles * batch
    int64_t chunks_per_a;     // ceil(KB / KT)
    int64_t total_chunks;     // KA * chunks_per_a
    int64_t chunks_per_split; // chunks handled by one gridDim.z slice
    int64_t nsplit;
};

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int W = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int W = 2; };

template <typename T> __device__ __forceinline__ void vec_unpack(const typename VecOf<T>::type& v, T* o);
template <> __device__ __forceinline__ void vec_unpack<float>(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <> __device__ __forceinline__ void vec_unpack<double>(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }

template <typename T> struct KTile { static constexpr int value = 128 / sizeof(T); };  // 32 fp32 / 16 fp64

// A_KMAJOR: the streamed operand is contiguous along the contraction index (sXb == 1);
// otherwise it is contiguous along m (sXm == 1).  Either way global loads are coalesced.
template <typename T, int TJ, int TR, bool A_KMAJOR>
__global__ void __launch_bounds__(128)
stream_gemm_kernel(const StreamGemmParams<T> p) {
    constexpr int VW = VecOf<T>::W;
    using Vec = typename VecOf<T>::type;
    constexpr int TM = 16 * TJ, TN = 8 * TR, KT = KTile<T>::value;
    constexpr int A_PER_THREAD = TM * KT / 128, B_PER_THREAD = KT * TN / 128;
    static_assert(TJ % VW == 0 && TR % VW == 0, "micro-tile must be a multiple of the vector width");
    static_assert(128 % TN == 0 || TN % 128 == 0, "TN must divide 128");

    __shared__ __align__(16) T As[TM * KT];   // A_KMAJOR: [m][k]   else: [k][m]
    __shared__ __align__(16) T Bs[KT * TN];   // [k][n]

    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;  // 8 threads across n, 16 across m
    const int64_t batch = blockIdx.x / p.m_tiles;
    const int64_t m0 = ((int64_t)blockIdx.x - batch * p.m_tiles) * TM;
    const int64_t n0 = (int64_t)blockIdx.y * TN;
    const int64_t split = blockIdx.z;
    const int64_t c_begin = split * p.chunks_per_split;
    const int64_t c_end = min(p.total_chunks, c_begin + p.chunks_per_split);

    const T* __restrict__ X = p.X + batch * p.sXbatch;

    T acc[TJ][TR];
#pragma unroll
    for (int i = 0; i < TJ; ++i)
#pragma unroll
        for (int j = 0; j < TR; ++j) acc[i][j] = T(0);

    T a_reg[A_PER_THREAD], b_reg[B_PER_THREAD];

    auto load_chunk = [&](int64_t c) {
        const int64_t a = c / p.chunks_per_a;
        const int64_t b0 = (c - a * p.chunks_per_a) * KT;
        const T* __restrict__ xa = X + a * p.sXa;
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; ++i) {
            const int e = tid + i * 128;
            int m, k;
            if (A_KMAJOR) { k = e % KT; m = e / KT; } else { m = e % TM; k = e / TM; }
            const int64_t gm = m0 + m, gb = b0 + k;
            a_reg[i] = (gm < p.M && gb < p.KB) ? __ldg(xa + gm * p.sXm + gb * p.sXb) : T(0);
        }
#pragma unroll
        for (int i = 0; i < B_PER_THREAD; ++i) {
            const int e = tid + i * 128;
            const int n = e % TN, k = e / TN;
            const int64_t gn = n0 + n, gb = b0 + k;
            T v = T(0);
            if (gn < p.N && gb < p.KB) {
                v = __ldg(p.Q + gb * p.sQb + gn * p.sQn);
                if (p.P) v *= __ldg(p.P + a * p.ldP + gn);
            }
            b_reg[i] = v;
        }
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; ++i) As[tid + i * 128] = a_reg[i];
#pragma unroll
        for (int i = 0; i < B_PER_THREAD; ++i) Bs[tid + i * 128] = b_reg[i];
    };

    if (c_begin < c_end) load_chunk(c_begin);
    for (int64_t c = c_begin; c < c_end; ++c) {
        __syncthreads();          // previous chunk fully consumed
        store_chunk();
        __syncthreads();
        if (c + 1 < c_end) load_chunk(c + 1);   // overlaps with the FMAs below

        if (A_KMAJOR) {
#pragma unroll 2
            for (int kk = 0; kk < KT; kk += VW) {
                T av[TJ][VW];
#pragma unroll
                for (int i = 0; i < TJ; ++i) {
                    const int m = ((i / VW) * 16 + ty) * VW + (i % VW);
                    Vec v = *reinterpret_cast<const Vec*>(&As[m * KT + kk]);
                    vec_unpack<T>(v, av[i]);
                }
#pragma unroll
                for (int u = 0; u < VW; ++u) {
                    T bv[TR];
#pragma unroll
                    for (int j = 0; j < TR / VW; ++j) {
                        Vec v = *reinterpret_cast<const Vec*>(&Bs[(kk + u) * TN + (j * 8 + tx) * VW]);
                        vec_unpack<T>(v, &bv[j * VW]);
                    }
#pragma unroll
                    for (int i = 0; i < TJ; ++i)
#pragma unroll
                        for (int j = 0; j < TR; ++j) acc[i][j] += av[i][u] * bv[j];
                }
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < KT; ++k) {
                T av[TJ], bv[TR];
#pragma unroll
                for (int i = 0; i < TJ / VW; ++i) {
                    Vec v = *reinterpret_cast<const Vec*>(&As[k * TM + (i * 16 + ty) * VW]);
                    vec_unpack<T>(v, &av[i * VW]);
                }
#pragma unroll
                for (int j = 0; j < TR / VW; ++j) {
                    Vec v = *reinterpret_cast<const Vec*>(&Bs[k * TN + (j * 8 + tx) * VW]);
                    vec_unpack<T>(v, &bv[j * VW]);
                }
#pragma unroll
                for (int i = 0; i < TJ; ++i)
#pragma unroll
                    for (int j = 0; j < TR; ++j) acc[i][j] += av[i] * bv[j];
            }
        }
    }

    // epilogue: thread owns rows m_i = ((i/VW)*16 + ty)*VW + i%VW, cols n_j = ((j/VW)*8 + tx)*VW + j%VW
    T* __restrict__ C = p.C + batch * p.sCbatch + split * p.sCsplit;
#pragma unroll
    for (int i = 0; i < TJ; ++i) {
        const int64_t gm = m0 + ((i / VW) * 16 + ty) * VW + (i % VW);
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TR; ++j) {
            const int64_t gn = n0 + ((j / VW) * 8 + tx) * VW + (j % VW);
            if (gn < p.N) C[gm * p.sCm + gn * p.sCn] = acc[i][j];
        }
    }
}

// Deterministic split-K reduction: out[m*ld + n] = sum_s partial[s][m][n]  (n < N)
template <typename T>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const T* __restrict__ partial, int64_t nsplit, int64_t M, int64_t N, int64_t Npad,
                     T* __restrict__ out, int64_t out_ld) {
    const int64_t total = M * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = idx / N, n = idx - m * N;
        const T* src = partial + m * Npad + n;
        T s = T(0);
        for (int64_t k = 0; k < nsplit; ++k) s += src[k * M * Npad];
        out[m * out_ld + n] = s;
    }
}

// Same sum for fp32 with 16-byte rows, built for latency: 256 threads = 32 float4 columns x 8 split groups; group g
// adds the partials s = g, g+8, ... in order, then the 8 group sums are combined in a fixed order through shared
// memory.  Deterministic (the order depends only on nsplit); ~3x faster than one thread walking every split.
template <int kUnused = 0>   // template only so that the header can be included from several translation units
__global__ void __launch_bounds__(256)
splitk_reduce_f4_kernel(const float4* __restrict__ partial, int nsplit, int64_t M, int N4, int Npad4,
                        float4* __restrict__ out, int64_t out_ld4) {
    __shared__ float4 red[8][32];
    const int tx = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t total = M * N4;
    const int64_t plane = M * Npad4;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
        const int64_t idx = base + tx;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        int64_t m = 0;
        int n = 0;
        if (idx < total) {
            m = idx / N4; n = (int)(idx - m * N4);
            const float4* src = partial + m * Npad4 + n;
#pragma unroll 4
            for (int k = g; k < nsplit; k += 8) {
                const float4 v = src[(int64_t)k * plane];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        red[g][tx] = s;
        __syncthreads();
        if (g == 0 && idx < total) {
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                const float4 v = red[j][tx];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            out[m * out_ld4 + n] = s;
        }
        __syncthreads();
    }
}

// Launch helper: picks the template instance from (TR, A_KMAJOR).
template <typename T>
int launch_stream_gemm(const StreamGemmParams<T>& p, int TR, bool a_kmajor, cudaStream_t stream);

// fp64 on the double-precision tensor cores (stream_gemm_dmma.cu): same parameters, TN in {32, 64}.
// TLB200_FP64_SIMT=1 keeps the SIMT kernel (A/B comparison).
bool stream_gemm_dmma_enabled();
int launch_stream_gemm_dmma(const StreamGemmParams<double>& p, int TN, bool a_kmajor, cudaStream_t stream);

// Tile configuration chosen for N output columns.
inline int stream_gemm_tr_for(int64_t N, int dtype) {
    const int vw = dtype == TLB200_F64 ? 2 : 4;
    if (N <= 16 && vw == 2) return 2;
    if (N <= 32) return 4;
    return 8;
}

}  // namespace tlb200
