// The one exchange step of the data-parallel path, hand-written for NVLink 5 / NVSwitch: all-reduce(sum) of the flat FP32
// vector through the switch's multicast + in-switch reduction (NVLS), SURVEY.md section 8e.
//
// The vector lives in a SYMMETRIC buffer (same size on every rank, mapped into every peer and behind one multicast
// address; torch.distributed._symmetric_memory does the allocation and the handle exchange -- plumbing).  Two-shot:
//   barrier            every rank's partial vector is complete and visible            (flags in the peers' signal pads)
//   reduce-scatter     rank r owns slice r: multimem.ld_reduce.add  reads the 8 partials of an element THROUGH the
//                      switch, which adds them on the way -- one load returns the sum
//   all-gather         multimem.st writes the sum to the same offset of ALL ranks' buffers -- one store, 8 copies
//   barrier            every slice has landed everywhere
// Each element is summed exactly once, by its owner, and the same bits are broadcast to everybody: the replicas stay
// bit-identical (which the CG replicas rely on: no scalar collectives).  Every rank ships (W-1)/W of its vector out
// (serving the owners' reductions) and takes (W-1)/W of the result in; the 11.3 MB vector of the autoencoder config
// takes 41-43 us on 8 GPUs (12.7 us of it launch + the two rendezvous) where ncclAllReduce, latency-bound at this
// size, takes 85 us (tools/scale_probe.py, profiles/r2_summary.md).
#include "common.cuh"

namespace hf {

__device__ __forceinline__ unsigned cas_release_sys(unsigned* p, unsigned expect, unsigned desired) {
  unsigned old;
  asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(expect), "r"(desired) : "memory");
  return old;
}
__device__ __forceinline__ unsigned cas_relaxed_sys(unsigned* p, unsigned expect, unsigned desired) {
  unsigned old;
  asm volatile("atom.relaxed.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(expect), "r"(desired) : "memory");
  return old;
}
__device__ __forceinline__ unsigned cas_acquire_sys(unsigned* p, unsigned expect, unsigned desired) {
  unsigned old;
  asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(expect), "r"(desired) : "memory");
  return old;
}

// All CTAs with the same blockIdx.x on all ranks meet here.  pads[r] = rank r's signal pad (zero-initialised, peer
// mapped); CTA b uses words [b*world, (b+1)*world) of every pad: word `rank` of the peer's pad is "rank has arrived".
// Signals are consumed (1 -> 0) by the waiter, so the barrier resets itself and can be reused back to back.
// ORDERED = false: pure rendezvous (what precedes it became visible at a kernel boundary already, and what follows are
// strong system-scope accesses); ORDERED = true: release on the way in (cumulative over the CTA through the bar.sync
// in front of it), acquire on the way out.
template <bool ORDERED>
__device__ __forceinline__ void rank_barrier(unsigned* const* pads, int rank, int world) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    unsigned* send = pads[peer] + (size_t)blockIdx.x * world + rank;
    unsigned* recv = pads[rank] + (size_t)blockIdx.x * world + peer;
    if (ORDERED) {
      while (cas_release_sys(send, 0u, 1u) != 0u) {
      }
      while (cas_acquire_sys(recv, 1u, 0u) != 1u) {
      }
    } else {
      while (cas_relaxed_sys(send, 0u, 1u) != 0u) {
      }
      while (cas_relaxed_sys(recv, 1u, 0u) != 1u) {
      }
    }
  }
  __syncthreads();
}

static int g_ar_variant = 3;  // bit 0: relaxed entry barrier; bit 1: no per-thread system fence in front of the exit barrier

template <int VARIANT>
__global__ void __launch_bounds__(512) allreduce_multimem_kernel(float* mc, unsigned* const* pads, int rank, int world, int64_t offset,
                                                                 int64_t count, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;  // uniform on all ranks: the solver state is replicated bit for bit
  // the partial vectors were written by earlier kernels of the launching streams: complete and in the owners' L2 (where
  // peer and multicast reads are served) when this kernel starts; the entry barrier only has to be a rendezvous
  if (VARIANT & 1)
    rank_barrier<false>(pads, rank, world);
  else
    rank_barrier<true>(pads, rank, world);
  // slice of this rank, in float4 units (count % (4 * world) == 0 is the caller's job: the buffer is padded)
  const int64_t per = count / 4 / world;
  const int64_t begin = offset / 4 + per * rank;
  float4* base = reinterpret_cast<float4*>(mc) + begin;
  constexpr int U = 4;  // independent ld_reduce in flight per thread: NVLink round trips overlap
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      if (j < per)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(base + j)
                     : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t j = i + u * stride;
      if (j < per)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(base + j), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z),
                     "f"(v[u].w)
                     : "memory");
    }
  }
  if (!(VARIANT & 2)) __threadfence_system();
  rank_barrier<true>(pads, rank, world);
}

}  // namespace hf

using namespace hf;

extern "C" void hf_debug_allreduce_variant(int32_t v) { g_ar_variant = v & 3; }

extern "C" int hf_allreduce_multimem(void* d_multicast, void* d_signal_pads, int32_t rank, int32_t world, int64_t offset, int64_t count,
                                     int32_t max_blocks, const int32_t* d_skip, void* stream) {
  HF_REQUIRE(d_multicast && d_signal_pads && world >= 2 && world <= 32 && rank >= 0 && rank < world, HF_ERR_INVALID,
             "hf_allreduce_multimem: bad arguments");
  HF_REQUIRE(offset % 4 == 0 && count % (4 * world) == 0 && (reinterpret_cast<uintptr_t>(d_multicast) & 15u) == 0, HF_ERR_INVALID,
             "hf_allreduce_multimem: offset must be a multiple of 4 floats and count of 4*world floats (pad the symmetric buffer)");
  if (count == 0) return HF_OK;
  const int64_t per = count / 4 / world;
  int64_t blocks = (per + 512 * 4 - 1) / (512 * 4);
  if (blocks > max_blocks) blocks = max_blocks;  // signal-pad words: blocks * world
  if (blocks > 64) blocks = 64;                  // co-resident by a wide margin (the barrier spins)
  if (blocks < 1) blocks = 1;
  auto* kern = allreduce_multimem_kernel<3>;
  if (g_ar_variant == 0) kern = allreduce_multimem_kernel<0>;
  if (g_ar_variant == 1) kern = allreduce_multimem_kernel<1>;
  if (g_ar_variant == 2) kern = allreduce_multimem_kernel<2>;
  kern<<<(unsigned)blocks, 512, 0, (cudaStream_t)stream>>>(static_cast<float*>(d_multicast), static_cast<unsigned* const*>(d_signal_pads), rank,
                                                           world, offset, count, d_skip);
  HF_LAUNCH_CHECK();
  return HF_OK;
}
