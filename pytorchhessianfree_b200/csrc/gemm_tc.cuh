// tcgen05 / TMEM / TMA contraction engines (split precision: TF32 main term + BF16 corrections) -- interface.
//   gemm_tc.cu   128x128 tiles, one CTA each, BF16 forms derived in the main loop: any FP32 operand, small grids
//   gemm_tc2.cu  256x256 tiles on CTA pairs, operands pre-split into images (Image16): large contractions
#pragma once
#include "gemm_simt.cuh"

namespace hf {
// below this many multiply-adds a 128x128 tensor tile is mostly padding: stay on the SIMT tiles
constexpr int64_t kTcMinWork = 1 << 20;
bool tc_supported(const GemmArgs& g);
int launch_gemm_tc(const GemmArgs& g, cudaStream_t stream);
int set_tc_trace(void* d_buf);
int set_tc_trace_iters(void* d_buf);

// pre-split pair engine: every operand must carry its image
bool tc2_supported(const GemmArgs& g);
int launch_gemm_tc2(const GemmArgs& g, cudaStream_t stream);
struct Tc2Choice {
  double us_pair, us_single;  // modelled time on the pair engine / on the 128x128 engine
};
Tc2Choice tc2_estimate(int M, int N, int K, int n_pairs, int splits);
int set_tc2_trace(void* d_buf);  // per-CTA phase timestamps of the pair kernel (d_buf: 4096 x 8 uint64), NULL = off
int tc2_mode();  // HF_TC2: 0 = never, 1 = where the model prefers it (default), 2 = wherever it is supported

// img = split(src) (+ optional 16-byte-pitched FP32 copy) for several matrices in one launch
struct SplitSegment {
  const float* src;
  int64_t rows;
  int cols;
  int64_t ld_src;
  float* dst32;  // optional FP32 copy with pitch ld32 (multiple of 4)
  int64_t ld32;
  Image16 img;   // optional
  int square;    // 1: every element is squared first (operands of the empirical-Fisher contraction)
  // perm_taps > 0: the columns are re-ordered on the way, dst column tap*perm_cin + c = src column c*perm_taps + tap:
  // a convolution weight [C_out, C_in, k_h, k_w] as PyTorch flattens it -> the tap-major operand conv.cuh contracts with
  int perm_cin, perm_taps;
};
constexpr int kMaxSplitSegments = 20;
struct SplitTable {
  SplitSegment seg[kMaxSplitSegments];
  int count;
  const int32_t* skip;
};
int launch_split(const SplitTable& t, cudaStream_t stream);
}  // namespace hf
