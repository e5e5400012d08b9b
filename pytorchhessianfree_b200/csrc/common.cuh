// Shared helpers for the hf_b200 kernels: error reporting across the C ABI, deterministic
// block/grid reductions, a sense-reversing grid barrier for cooperative launches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hf_b200.h"

namespace hf {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
void note_launch();

#define HF_CUDA(expr)                                            \
  do {                                                           \
    cudaError_t e__ = (expr);                                    \
    if (e__ != cudaSuccess) return ::hf::cuda_fail(e__, #expr);  \
  } while (0)

// every kernel launch goes through one of these two, so hf_debug_launch_count() is exact
#define HF_STR2(x) #x
#define HF_STR(x) HF_STR2(x)
#define HF_LAUNCH_CHECK()                                                                          \
  do {                                                                                             \
    ::hf::note_launch();                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess) return ::hf::cuda_fail(e__, "kernel launch at " __FILE__ ":" HF_STR(__LINE__)); \
  } while (0)

#define HF_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::hf::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

// cudaFuncSetAttribute opt-ins (large dynamic shared memory) are per device: `seen` is a static table owned by the call
// site; true the first time that site runs on the current device (a process may drive several GPUs).
inline bool first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

constexpr int kMaxCtas = 256;  // upper bound on the persistent grid (148 SMs on B200)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Sense-reversing barrier over all CTAs of a cooperative launch.  bar[0] = arrival count,
// bar[1] = generation.  Self-resetting, so it survives any number of launches and skipped launches.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned n_ctas) {
  __syncthreads();
  if (n_ctas > 1 && threadIdx.x == 0) {
    const unsigned gen = ld_acquire_u32(bar + 1);
    __threadfence();
    const unsigned prev = atomicAdd(bar, 1u);
    if (prev == n_ctas - 1) {
      atomicExch(bar, 0u);
      __threadfence();
      st_release_u32(bar + 1, gen + 1);
    } else {
      while (ld_acquire_u32(bar + 1) == gen) __nanosleep(32);
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum NV doubles per thread over the block, fixed order; result valid in every thread of warp 0
// (and broadcast through `scratch`, NV*33 doubles, to all threads after the trailing barrier).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // scratch may still be read from a previous call
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 33 + warp] = v[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = lane < nwarp ? scratch[i * 33 + lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) scratch[i * 33 + 32] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = scratch[i * 33 + 32];
}

// Grid-wide all-reduce of NV doubles for a cooperative (co-resident) launch, fused with the barrier it implies.
// Low-latency flag-in-data protocol: every 64-bit word a CTA publishes carries 32 payload bits and the 32-bit
// generation, so a reader that sees the right generation has the payload too -- one store, one poll round trip, no
// fences, no atomics, no contended counter.  Each of the first kMaxCtas/32 warps polls 32 CTAs, the per-warp sums are
// combined in a fixed order, so every CTA, every run and every data-parallel rank derives bit-identical scalars.
// `gen` must be a value the slot array has not seen since it was zeroed (callers use iteration + 1); each array is
// used at most once per launch, so a kernel boundary separates consecutive uses.
struct alignas(64) ReduceSlot {
  unsigned long long w[8];  // w[2i] = {low 32 bits of value i, gen}, w[2i+1] = {high 32 bits of value i, gen}
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <int NV>
__device__ __forceinline__ void grid_allreduce(ReduceSlot* slots, unsigned n_ctas, unsigned gen, double (&v)[NV],
                                               double* scratch) {
  static_assert(NV <= 3, "slot holds three values");
  block_sum<NV>(v, scratch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ReduceSlot* mine = slots + blockIdx.x;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v[i]);
      st_relaxed_u64(&mine->w[2 * i], ((unsigned long long)gen << 32) | (bits & 0xffffffffull));
      st_relaxed_u64(&mine->w[2 * i + 1], ((unsigned long long)gen << 32) | (bits >> 32));
    }
  }
  double* wsum = scratch + 100;  // [kMaxCtas/32][3], clear of block_sum's area
  if (warp < kMaxCtas / 32) {
    const unsigned c = warp * 32 + lane;
    double t[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) t[i] = 0.0;
    if (c < n_ctas) {
      unsigned long long w[2 * NV];
      bool ready;
      do {
        ready = true;
#pragma unroll
        for (int j = 0; j < 2 * NV; ++j) {
          w[j] = ld_relaxed_u64(&slots[c].w[j]);
          ready = ready && ((unsigned)(w[j] >> 32) == gen);
        }
      } while (!ready);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        t[i] = __longlong_as_double((long long)((w[2 * i] & 0xffffffffull) | (w[2 * i + 1] << 32)));
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      t[i] = warp_sum(t[i]);
      if (lane == 0) wsum[warp * 3 + i] = t[i];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kMaxCtas / 32; ++w) t += wsum[w * 3 + i];
    v[i] = t;
  }
  __syncthreads();
}

}  // namespace hf
