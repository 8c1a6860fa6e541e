// tcgen05 "stream GEMM": the tensor-core engine shared by MTTKRP and TTM.
//
//   C[kr][m, n] = sum over the K chunks of item (mt, kr) of  X_tile[128 x KS] * B_tile[KS x RP]
//
// X is the big tensor, streamed from HBM exactly once by TMA; B is small (Khatri-Rao rows
// synthesised on the fly, or a pre-split factor matrix).  fp32 in, fp32 out, error-
// compensated 3xTF32 on the tensor cores (see tc_stream.cu for the full design notes).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace tlb200 {

enum TcXLayout {
    TC_X_KMAJOR_1 = 0,   // tile = 128 rows x 32 k (one 128-byte line per row)         [any B]
    TC_X_KMAJOR_2 = 1,   // tile = 128 rows x 64 k (two adjacent lines per row: DRAM-friendly) [B % 32 == 0]
    TC_X_MMAJOR = 2      // tile = 64 k-rows x 128 m (m contiguous in memory)
};
enum TcBMode {
    TC_B_MAT = 1         // B(k=b, n) from pre-split hi/lo matrices via TMA ([RP rows][K], K-major)
};

struct TcStreamParams {
    // K space: A x B, chunked along B in steps of KS
    int64_t M;                 // extent of the streamed (row) dim
    int64_t A, B;
    int64_t chunks_per_a;      // ceil(B / KS)
    // items: (mt, bb, kr) with mt < m_tiles, bb < n_bblocks, kr < k_ranges; item (bb, kr) covers the chunks
    // [bb * nb, min(chunks_per_a, (bb+1) * nb)) of every a in [kr * a_per_range, min(A, (kr+1) * a_per_range))
    // and writes partial result number kr * n_bblocks + bb
    int m_tiles;
    int nb, n_bblocks;
    int64_t k_ranges;
    int64_t a_per_range;
    // 1: the b block of an item (nb chunks of the small operand) is loaded once and stays in shared memory
    //    (needs nb * KS / 32 <= tc_b_slots(rp)); 0: the small operand is streamed with the tiles
    int b_resident;
    int group_units;           // 32-element K units per TMEM accumulation group (RZ accumulate => keep short)
    // optional per-`a` scaling of the result (MTTKRP: the outer Khatri-Rao table), applied by the epilogue
    const float* P;            // [A][RP] or null
    // output: out[kr * sOk + m * sOm + n * sOn], n < n_valid
    float* out;
    int64_t sOk, sOm, sOn;
    int n_valid;
    // fp16-split engine only (TcStreamLaunch::hf): device scalar max |x| over the tensor (the range hint) and the
    // inverse power-of-two column scales of the small operand [RP]
    const float* x_absmax;
    const float* col_inv;
    long long* trace;          // perf triage only: per-role clock64 timestamps of CTA 0 (null = off)
    int debug;                 // TLB200_TC_DEBUG bitmask (perf triage only): 1 skip KR math, 2 skip MMAs, 4 skip TMEM stores, 8 skip epilogue loads
};

struct TcStreamLaunch {
    TcStreamParams p;
    CUtensorMap x_map;         // see tc_stream.cu for the dims per layout
    CUtensorMap bhi_map, blo_map;   // pre-split small operand: [RP rows][Kpad] hi / lo, box {32, RP}, SWIZZLE_128B
    int rp;                    // 32 or 64
    int x_layout;              // TcXLayout
    int b_mode;                // TcBMode
    int hf;                    // 1: fp16-split engine (64-element tiles only; B maps are fp16 [RP rows][Kpad], box {64, RP})
};

// true when cuTensorMapEncodeTiled could be resolved (a driver is present)
bool tc_available();
// encode helper (returns TLB200_ECUDA on failure)
int tc_encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128, bool half = false);
int tc_stream_launch(const TcStreamLaunch& l, cudaStream_t stream);
// K extent of one chunk for a layout
inline int tc_chunk_k(int x_layout) { return x_layout == TC_X_KMAJOR_1 ? 32 : 64; }
int tc_group_units(bool hf = false);
// shared-memory slots (32-element units) for the small operand
inline int tc_b_slots(int rp) { return rp == 32 ? 8 : 6; }

}  // namespace tlb200
