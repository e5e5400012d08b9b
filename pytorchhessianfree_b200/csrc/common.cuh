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
#define HF_LAUNCH_CHECK()            \
  do {                               \
    ::hf::note_launch();             \
    HF_CUDA(cudaGetLastError());     \
  } while (0)

#define HF_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::hf::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

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

// Grid-wide all-reduce of NV doubles for a cooperative (co-resident) launch, fused with the barrier it implies:
// every CTA publishes {values, generation} in its own 32-byte slot (plain stores + one release store: no atomics, no
// contended counter), then polls the generation words of all CTAs and sums the values in one fixed order.  Every
// CTA, every run and every data-parallel rank therefore derives bit-identical scalars.  `gen` must be a value the
// slot array has not seen since it was zeroed (callers use iteration + 1); each array is used at most once per
// launch, so a kernel boundary separates consecutive uses.
struct alignas(32) ReduceSlot {
  double v[3];
  unsigned gen;
  unsigned pad_;
};

template <int NV>
__device__ __forceinline__ void grid_allreduce(ReduceSlot* slots, unsigned n_ctas, unsigned gen, double (&v)[NV],
                                               double* scratch) {
  static_assert(NV <= 3, "slot holds three values");
  block_sum<NV>(v, scratch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ReduceSlot* mine = slots + blockIdx.x;
#pragma unroll
    for (int i = 0; i < NV; ++i) mine->v[i] = v[i];
    __threadfence();
    st_release_u32(&mine->gen, gen);
  }
  if (warp == 0) {
    for (unsigned c = lane; c < n_ctas; c += 32)
      while (ld_acquire_u32(&slots[c].gen) != gen) {
      }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = 0.0;
      for (unsigned c = lane; c < n_ctas; c += 32) t += __ldcg(&slots[c].v[i]);
      t = warp_sum(t);
      if (lane == 0) scratch[i] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = scratch[i];
  __syncthreads();
}

}  // namespace hf
