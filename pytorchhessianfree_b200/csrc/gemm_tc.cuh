// tcgen05 / TMEM / TMA contraction engine (3xTF32 split precision) -- interface.
#pragma once
#include "gemm_simt.cuh"

namespace hf {
bool tc_supported(const GemmArgs& g);
int launch_gemm_tc(const GemmArgs& g, cudaStream_t stream);
}  // namespace hf
