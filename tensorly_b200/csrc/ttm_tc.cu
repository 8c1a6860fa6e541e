// tcgen05 TTM path (placeholder until the kernel lands: every call takes the SIMT path).
#include "ttm_tc.cuh"

namespace tlb200 {
bool ttm_tc_supported(int64_t, int64_t, int64_t, int64_t) { return false; }
int ttm_tc_launch(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, int64_t, float*,
                  cudaStream_t) { return TLB200_EUNSUPPORTED; }
}  // namespace tlb200
