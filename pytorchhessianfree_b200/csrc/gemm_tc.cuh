// tcgen05 / TMEM / TMA contraction engine (split precision: TF32 + BF16 corrections) -- interface.
#pragma once
#include "gemm_simt.cuh"

namespace hf {
// below this many multiply-adds a 128x128 tensor tile is mostly padding: stay on the SIMT tiles
constexpr int64_t kTcMinWork = 1 << 20;
bool tc_supported(const GemmArgs& g);
int launch_gemm_tc(const GemmArgs& g, cudaStream_t stream);
int set_tc_trace(void* d_buf);
int set_tc_trace_iters(void* d_buf);
}  // namespace hf
