// tcgen05 MTTKRP path (placeholder until the kernel lands: reports "unsupported" so that
// every plan resolves to the SIMT path).
#include "mttkrp_tc.cuh"

namespace tlb200 {
bool mttkrp_tc_supported(const tlb200_mttkrp_plan_t&, int64_t, int) { return false; }
void mttkrp_tc_fill_plan(tlb200_mttkrp_plan_t*, int64_t) {}
size_t mttkrp_tc_extra_workspace(const tlb200_mttkrp_plan_t&) { return 0; }
int mttkrp_tc_launch(const float*, const tlb200_mttkrp_plan_t&, int64_t, const float*, const float*, float*, void*,
                     cudaStream_t) { return TLB200_EUNSUPPORTED; }
}  // namespace tlb200
