// tcgen05 reconstruction / imputation path — interface used by recon.cu.
#pragma once
#include "common.cuh"

namespace tlb200 {
// fp32, rank <= 64, problems large enough to fill the machine
bool recon_tc_supported(const int64_t* shape, int ndim, int64_t rank, int dtype);
// persistent grid size = number of per-CTA partial sums the imputation variant writes
int recon_tc_grid(const int64_t* shape, int ndim);
// x == null: out = rec (mask == null) or rec * mask; else out = x*mask + rec*(1-mask), partial[cta] = {sum out^2, sum (out-rec)^2}
int recon_tc_launch(const void* const* factors, const int64_t* shape, const int64_t* frs, const int64_t* fcs, int ndim,
                    int64_t rank, const float* w, const float* x, const float* mask, float* out, double* partial,
                    cudaStream_t stream);
}  // namespace tlb200
