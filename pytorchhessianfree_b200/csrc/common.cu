// Error reporting and device queries shared by the hf_b200 translation units.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace hf {

static thread_local char g_error[512] = "";
static long long g_launches = 0;
void note_launch() { ++g_launches; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return HF_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 1;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace hf

extern "C" {

int hf_abi_version(void) { return HF_ABI_VERSION; }
const char* hf_last_error_string(void) { return hf::g_error; }
int hf_device_sm_count(void) { return hf::sm_count(); }
long long hf_debug_launch_count(void) { return hf::g_launches; }

}  // extern "C"
