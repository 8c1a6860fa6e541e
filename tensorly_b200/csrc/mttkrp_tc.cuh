// tcgen05 (5th-gen tensor core) MTTKRP path — interface used by mttkrp.cu.
#pragma once
#include "common.cuh"

namespace tlb200 {

// True when the fp32 3xTF32 tcgen05 engine (tc_stream.cu) can run this plan.
bool mttkrp_tc_supported(const tlb200_mttkrp_plan_t& pl, int64_t rank, int dtype);
// Fills rank_padded and splits for the tcgen05 engine's tiling.
void mttkrp_tc_fill_plan(tlb200_mttkrp_plan_t* pl, int64_t rank);
size_t mttkrp_tc_extra_workspace(const tlb200_mttkrp_plan_t& pl);
// Writes partial[split][J][rank_padded]; the caller runs the deterministic reduction.
// x_absmax != null: the fp16-split engine — Q then holds fp16 hi / lo tables [rank_padded][Bpad] followed (at float
// offset rank_padded * Bpad) by the inverse column scales (mttkrp.cu builds them); needs mttkrp_tc_hf_ok(pl).
int mttkrp_tc_launch(const float* x, const tlb200_mttkrp_plan_t& pl, int64_t rank, const float* P,
                     const float* Q, float* partial, void* extra_ws, cudaStream_t stream, const float* x_absmax = nullptr);
// the fp16-split engine streams 64-element tiles only
bool mttkrp_tc_hf_ok(const tlb200_mttkrp_plan_t& pl);

}  // namespace tlb200
