#include "gemm_tc.cuh"
namespace hf {
bool tc_supported(const GemmArgs&) { return false; }
int launch_gemm_tc(const GemmArgs&, cudaStream_t) {
  set_error("tcgen05 engine not built");
  return HF_ERR_UNSUPPORTED;
}
}  // namespace hf
