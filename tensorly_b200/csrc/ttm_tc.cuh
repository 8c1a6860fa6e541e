// tcgen05 TTM path — interface used by ttm.cu.
#pragma once
#include "common.cuh"

namespace tlb200 {
// out[l, i, t] = sum_j M[i*mrs + j*mcs] * X[l, j, t], fp32 in/out, 3xTF32 on tcgen05.
bool ttm_tc_supported(int64_t L, int64_t J, int64_t T, int64_t I);
size_t ttm_tc_workspace(int64_t L, int64_t J, int64_t T, int64_t I);
int ttm_tc_launch(const float* x, int64_t L, int64_t J, int64_t T, const float* m, int64_t I, int64_t mrs,
                  int64_t mcs, float* out, void* workspace, cudaStream_t stream);
}  // namespace tlb200
