// Template instantiation + launcher for the SIMT stream GEMM (see stream_gemm.cuh).
#include "stream_gemm.cuh"

namespace tlb200 {

template <typename T, int TR, bool KM>
static int launch_one(const StreamGemmParams<T>& p_in, cudaStream_t stream) {
    constexpr int TJ = 8;
    constexpr int TM = 16 * TJ, TN = 8 * TR;
    StreamGemmParams<T> p = p_in;
    p.m_tiles = ceil_div(p.M, TM);
    const int64_t gx = p.m_tiles * p.nbatch, gy = ceil_div(p.N, TN), gz = p.nsplit;
    if (gx <= 0 || gy <= 0 || gz <= 0) return TLB200_OK;
    if (gx > 0x7fffffffLL || gy > 65535 || gz > 65535) return TLB200_EUNSUPPORTED;
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
    stream_gemm_kernel<T, TJ, TR, KM><<<grid, 128, 0, stream>>>(p);
    TLB_CHECK_LAUNCH();
    return TLB200_OK;
}

template <>
int launch_stream_gemm<float>(const StreamGemmParams<float>& p, int TR, bool km, cudaStream_t s) {
    if (TR == 4) return km ? launch_one<float, 4, true>(p, s) : launch_one<float, 4, false>(p, s);
    if (TR == 8) return km ? launch_one<float, 8, true>(p, s) : launch_one<float, 8, false>(p, s);
    return TLB200_EINVAL;
}

template <>
int launch_stream_gemm<double>(const StreamGemmParams<double>& p, int TR, bool km, cudaStream_t s) {
    // the DMMA kernel covers the SIMT tile widths 16 / 32 (TN = 32) and 64 (TN = 64) with the same column blocking
    if (stream_gemm_dmma_enabled() && (TR == 2 || TR == 4 || TR == 8)) return launch_stream_gemm_dmma(p, TR == 8 ? 64 : 32, km, s);
    if (TR == 2) return km ? launch_one<double, 2, true>(p, s) : launch_one<double, 2, false>(p, s);
    if (TR == 4) return km ? launch_one<double, 4, true>(p, s) : launch_one<double, 4, false>(p, s);
    if (TR == 8) return km ? launch_one<double, 8, true>(p, s) : launch_one<double, 8, false>(p, s);
    return TLB200_EINVAL;
}

}  // namespace tlb200
