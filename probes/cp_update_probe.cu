// Phase timing of cp_update_kernel (CTA 0): nvcc -DTLB_CP_TRACE ... probes/cp_update_probe.cu
#define TLB_CP_TRACE 1
#include "../tensorly_b200/csrc/cp_als.cu"
#include <cstdio>
#include <vector>
#include <cstdlib>
namespace tlb200 { void set_last_path(const char*) {} void count_launch() {} }
int main(int argc, char** argv) {
    int R = argc > 1 ? atoi(argv[1]) : 32, rows = argc > 2 ? atoi(argv[2]) : 1024;
    std::vector<float> f((size_t)rows * R), g((size_t)R * R), mm((size_t)rows * R);
    srand(1);
    for (auto& v : f) v = rand() / (float)RAND_MAX;
    for (auto& v : mm) v = rand() / (float)RAND_MAX;
    for (int i = 0; i < R; ++i) for (int j = 0; j < R; ++j) { double s = 0; for (int r = 0; r < rows; ++r) s += f[(size_t)r * R + i] * f[(size_t)r * R + j]; g[i * R + j] = (float)s; }
    float *dg, *dm, *dout; cudaMalloc(&dg, g.size() * 4); cudaMalloc(&dm, mm.size() * 4); cudaMalloc(&dout, mm.size() * 4);
    cudaMemcpy(dg, g.data(), g.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dm, mm.data(), mm.size() * 4, cudaMemcpyHostToDevice);
    const void* grams[3] = {dg, dg, dg};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        int st = tlb200::cp_update_launch<float>(grams, 3, 0, R, nullptr, 0.0, dm, R, rows, dout, R, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long t[16]; cudaMemcpyFromSymbol(t, tlb200::g_cp_trace, sizeof(t));
        printf("R=%d rows=%d st=%d %.1f us | formV %lld  LU %lld  permute %lld  fwd %lld  back %lld  store %lld clk\n", R, rows, st, ms * 1e3,
               t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5]);
    }
    return 0;
}
