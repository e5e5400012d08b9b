// Issue-rate microbenchmark for tcgen05.mma on B200 (timing only: operands are whatever shared memory holds).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
//
// One CTA (or CTA pair) per SM.  The issuing thread runs the k-block program of the split-precision contraction
// (TF32 main term + BF16 corrections) for a given instruction shape and reports cycles per k-block of 32 floats
// of K.  A second elected thread can stream bulk copies (L2 -> shared memory) of a given size per k-block into the
// same CTA at the same time, so the interference of the operand feed with the MMA operand reads is visible.
// Decides: N = 128 vs 256 per instruction, cta_group 1 vs 2, A from shared memory vs TMEM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                    \
    }                                                                             \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// K-major FP32 tile [rows][32 floats], SWIZZLE_128B; k-step of 8 floats = +32 B
__device__ __forceinline__ uint64_t desc_f32(uint32_t base, int ks) { return smem_desc(base + ks * 32, 16, 1024, 2); }
// K-major BF16 tile [rows][32 bf16], SWIZZLE_64B; k-step of 16 bf16 = +32 B
__device__ __forceinline__ uint64_t desc_b16(uint32_t base, int ks) { return smem_desc(base + ks * 32, 16, 512, 4); }

template <int CG, bool TS, bool F16>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a_desc, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if (TS) {
    if (F16)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  } else if (CG == 1) {
    if (F16)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  } else {
    if (F16)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

struct Params {
  int n_kb;        // k-blocks of 32 floats
  int n_tf32;      // tf32 MMAs per k-block (K = 8 each)
  int n_bf16;      // bf16 MMAs per k-block (K = 16 each)
  int N;           // instruction N: 128 or 256
  int tma_bytes;   // bulk-copy bytes per k-block per CTA (0 = none)
  int distinct;    // 1: every MMA of a k-block reads its own operand tiles (as the real loop does)
  const uint8_t* src;
  long long* out;  // [ctas][2]: cycles, ns
};

constexpr int STAGE = 96 * 1024, NSTAGE = 2;

template <int CG, bool TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Params p) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(tiles + NSTAGE * STAGE);
  uint64_t* ring = bars;       // [4] MMA commits, ring of 4 k-blocks in flight
  uint64_t* fin = bars + 4;
  uint64_t* tma_bar = bars + 5;  // [4]
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&ring[i], 1), mbar_init(&tma_bar[i], 1);
    mbar_init(fin, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 && lane == 0 && rank == 0) {
    const int M = 128 * CG;
    const uint32_t id32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t id16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int b_rows = p.N / CG;  // rows of B each CTA holds
    long long t0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int kb = 0; kb < p.n_kb; ++kb) {
      const int s = kb & 3;
      if (kb >= 4) mbar_wait(&ring[s], ((kb >> 2) - 1) & 1);
      const uint32_t base = smem_u32(tiles) + (kb & 1) * STAGE;
      const uint32_t a32 = base, alo = base + 16384, ahi = base + 24576;
      const uint32_t b32 = base + 32768, blo = b32 + b_rows * 128, bhi = blo + b_rows * 64;
      for (int i = 0; i < p.n_tf32; ++i) {
        const int ks = p.distinct ? (i & 3) : 0;
        mma<CG, TS, false>(tmem, desc_f32(a32, ks), tmem + 256 + 8 * ks, desc_f32(b32, ks), id32, (kb | i) != 0);
      }
      for (int i = 0; i < p.n_bf16; ++i) {
        const int ks = p.distinct ? ((i >> 1) & 1) : 0;
        const bool second = p.distinct && (i & 1);
        mma<CG, TS, true>(tmem, desc_b16(second ? ahi : alo, ks), tmem + 256 + 32 + 8 * i, desc_b16(second ? blo : bhi, ks), id16, 1);
      }
      commit<CG>(&ring[s]);
    }
    commit<CG>(fin);
    mbar_wait(fin, 0);
    long long t1 = clock64();
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    p.out[2 * blockIdx.x] = t1 - t0;
    p.out[2 * blockIdx.x + 1] = (long long)(g1 - g0);
  } else if (warp == 2 && lane == 0 && p.tma_bytes > 0) {
    // bulk copies global (L2-resident) -> shared, 4 batches in flight, same k-block count as the MMA loop
    const uint8_t* src = p.src + (size_t)blockIdx.x * (1 << 20);
    for (int kb = 0; kb < p.n_kb; ++kb) {
      const int s = kb & 3;
      if (kb >= 4) mbar_wait(&tma_bar[s], ((kb >> 2) - 1) & 1);
      mbar_expect_tx(&tma_bar[s], p.tma_bytes);
      for (int off = 0; off < p.tma_bytes; off += 16384) {
        const int n = min(16384, p.tma_bytes - off);
        const uint32_t dst = smem_u32(tiles) + ((kb & 1) * STAGE + off) % (NSTAGE * STAGE);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src + ((size_t)kb * 98304 + off) % (1 << 20)), "r"(n), "r"(smem_u32(&tma_bar[s]))
                     : "memory");
      }
    }
    for (int s = 0; s < 4; ++s) {
      const int last = ((p.n_kb - 1 - s) / 4) * 4 + s;  // last k-block that used slot s
      if (last >= 0 && last < p.n_kb) mbar_wait(&tma_bar[s], (last >> 2) & 1);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (warp == 1) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int CG, bool TS>
static void run(const char* name, Params p, int ctas) {
  const int smem = NSTAGE * STAGE + 1024 + 256;
  CK(cudaFuncSetAttribute(mma_rate_kernel<CG, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas), cfg.blockDim = dim3(128), cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG, TS>, p));
    CK(cudaDeviceSynchronize());
  }
  long long* h = (long long*)malloc(sizeof(long long) * 2 * ctas);
  CK(cudaMemcpy(h, p.out, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost));
  double cyc = 0, ns = 0;
  int n = 0;
  for (int c = 0; c < ctas; c += CG) cyc += (double)h[2 * c], ns += (double)h[2 * c + 1], ++n;
  cyc /= n, ns /= n;
  const double flop = 2.0 * 128 * CG * p.N * 32 * p.n_kb;  // algorithmic: one product per (m, n, k)
  printf("%-34s N=%3d tf32x%d bf16x%d feed %3d KB/kb: %7.1f cyc/kb %6.3f us/kb  (%5.1f cyc/MMA)  %6.1f TF/s algorithmic over 148 SMs\n", name, p.N,
         p.n_tf32, p.n_bf16, p.tma_bytes / 1024, cyc / p.n_kb, ns / p.n_kb * 1e-3, cyc / p.n_kb / (p.n_tf32 + p.n_bf16),
         flop / (ns * 1e-9) / 1e12 * (148 / CG));
  free(h);
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
  const int ctas = 148;
  Params p = {};
  uint8_t* src;
  CK(cudaMalloc(&src, (size_t)ctas << 20));
  CK(cudaMemset(src, 0, (size_t)ctas << 20));
  CK(cudaMalloc(&p.out, sizeof(long long) * 2 * ctas));
  p.src = src, p.n_kb = 512, p.distinct = 1;
  const int feeds[] = {0, 32 << 10, 64 << 10, 96 << 10};
  for (int N : {128, 256}) {
    p.N = N;
    for (int mix = 0; mix < 3; ++mix) {
      p.n_tf32 = mix == 0 ? 4 : (mix == 1 ? 8 : 0), p.n_bf16 = mix == 0 ? 4 : (mix == 1 ? 0 : 8);
      for (int f : feeds) {
        if (mix != 0 && f != 0 && f != (64 << 10)) continue;
        p.tma_bytes = f;
        run<1, false>("cta_group::1 SS", p, ctas);
      }
    }
    p.n_tf32 = 4, p.n_bf16 = 4;
    for (int f : feeds) {
      p.tma_bytes = f;
      run<1, true>("cta_group::1 TS (A in TMEM)", p, ctas);
    }
    for (int f : feeds) {
      p.tma_bytes = f;
      run<2, false>("cta_group::2 SS (M=256)", p, ctas);
    }
    p.distinct = 0, p.tma_bytes = 0;
    run<1, false>("cta_group::1 SS same tiles", p, ctas);
    p.distinct = 1;
  }
  return 0;
}
