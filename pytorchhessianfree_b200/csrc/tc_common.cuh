// Pieces shared by the two tcgen05 contraction kernels (gemm_tc.cu: 128x128 tiles with the in-kernel splitter;
// gemm_tc2.cu: 256x256 CTA-pair tiles on pre-split operand images): inline-PTX wrappers (mbarrier, TMA, tcgen05),
// the UMMA shared-memory descriptors of the operand layouts, and the fused epilogue.
#pragma once
#include <cuda.h>

#include "gemm_tc.cuh"

namespace hf {

constexpr int BKT = 32;  // k-block: 32 floats = one 128-byte swizzle row of the FP32 tiles
// (16-float k-blocks with two CTAs per SM were measured in round 1: 10 % slower, one barrier round trip per 16 columns)

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// CTA-pair form: the box lands in the issuing CTA's shared memory, the bytes are counted on `bar_cluster`, a
// shared::cluster address that may name the barrier of the other CTA of the pair (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- A operand from TMEM (TS mode): the instruction fetches only B from shared memory ----
__device__ __forceinline__ void umma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one 32-bit word per lane and column: lane = this thread's row, 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---- CTA pair (cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the object at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(p)), "r"(rank));
  return remote;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_cta(bar, rank)) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the pair -> one arrival on the barrier at this offset in every CTA of `cta_mask`
// (default: the two CTAs of a cluster that is one pair)
__device__ __forceinline__ void umma2_commit(uint64_t* bar, uint16_t cta_mask = 3) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA shared-memory matrix descriptors ------------------------------------------------------------
// cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46) version[46,48)=1 layout_type[61,64); offsets in
// 16-byte units.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// FP32 operand tile of 128 (M or N) x 32 (K) floats at `base`, k-step ks (8 floats):
//   K-major : 128-byte rows, SWIZZLE_128B (type 2), 8-row swizzle atoms 1024 B apart (SBO); step = +32 B in the row
//   MN-major: 32-bit operands only exist in SWIZZLE_128B_BASE32B (type 1; cute Layout_MN_SW128_32B_Atom, TMA
//             SWIZZLE_128B_ATOM_32B): atoms of [4 k-rows x 128 B] 512 B apart along K (SBO); 4 column blocks of
//             [32 k-rows x 128 B] 4096 B apart along MN (LBO); one k-step = 8 k-rows = +1024 B
__device__ __forceinline__ uint64_t operand_desc(uint32_t base, int mn_major, int ks) {
  return mn_major ? smem_desc(base + ks * 1024, BKT * 128, 512, 1) : smem_desc(base + ks * 32, 16, 8 * BKT * 4, 2);
}
// BF16 tile of 128 (M or N) x 32 (K) bf16 at `base`, k-step ks (16 bf16):
//   K-major : 64-byte rows, SWIZZLE_64B (type 4), 8-row atoms 512 B apart (SBO); step = +32 B in the row
//   MN-major: two column blocks of [32 k-rows x 128 B = 64 bf16] 4096 B apart along MN (LBO), SWIZZLE_128B (type 2),
//             atoms of 8 k-rows 1024 B apart along K (SBO); one k-step = 16 k-rows = +2048 B
__device__ __forceinline__ uint64_t corr_desc(uint32_t base, int mn_major, int ks) {
  return mn_major ? smem_desc(base + ks * 2048, 4096, 1024, 2) : smem_desc(base + ks * 32, 16, 512, 4);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a/b format [7,10)/[10,13) (TF32 = 2, BF16 = 1),
// majors 15/16, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc(uint32_t fmt, int a_mn, int b_mn, int M, int N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// the two BF16 words of the split-precision image of x: hi = bf16(x), lo = bf16(x - tf32_trunc(x)) (exact remainder)
__device__ __forceinline__ float tf32_rest(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {  // element 0 in the low half
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// ---- fused epilogue, one specialisation per (epilogue kind, activation): the row loop is straight-line vector code
// (with one warp per scheduler every branch and dependent ALU op is exposed latency, so nothing is decided per element)
template <int ACT>
__device__ __forceinline__ float d1(float s) {
  if (ACT == HF_ACT_RELU) return s > 0.f ? 1.f : 0.f;
  if (ACT == HF_ACT_SIGMOID) return s * (1.f - s);
  if (ACT == HF_ACT_TANH) return 1.f - s * s;
  return 1.f;
}
template <int ACT>
__device__ __forceinline__ float d2(float s) {
  if (ACT == HF_ACT_SIGMOID) return s * (1.f - s) * (1.f - 2.f * s);
  if (ACT == HF_ACT_TANH) return -2.f * s * (1.f - s * s);
  return 0.f;
}
template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
  if (ACT == HF_ACT_RELU) return z > 0.f ? z : 0.f;
  if (ACT == HF_ACT_SIGMOID) return 1.f / (1.f + expf(-z));
  if (ACT == HF_ACT_TANH) return tanhf(z);
  return z;
}

template <bool VEC>
__device__ __forceinline__ void ld4(const float* p, int cnt, float (&o)[4]) {
  if (VEC) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x, o[1] = t.y, o[2] = t.z, o[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = e < cnt ? p[e] : 0.f;
  }
}
template <bool VEC>
__device__ __forceinline__ void st4(float* p, int cnt, const float (&o)[4]) {
  if (VEC) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < cnt) p[e] = o[e];
  }
}
// split-precision image of four consecutive elements (the image row pitch is a multiple of 8 and n of 4: 8-byte stores)
__device__ __forceinline__ void st4_image(const Image16& im, int64_t m, int n, int cnt, const float (&o)[4]) {
  uint16_t* hi = im.hi + m * im.ld + n;
  uint16_t* lo = hi + im.plane;
  if (cnt == 4) {
    *reinterpret_cast<uint2*>(hi) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
    *reinterpret_cast<uint2*>(lo) = make_uint2(pack_bf16x2(tf32_rest(o[0]), tf32_rest(o[1])), pack_bf16x2(tf32_rest(o[2]), tf32_rest(o[3])));
  } else {
    for (int e = 0; e < cnt; ++e) {
      hi[e] = (uint16_t)(pack_bf16x2(o[e], 0.f) & 0xffffu);
      lo[e] = (uint16_t)(pack_bf16x2(tf32_rest(o[e]), 0.f) & 0xffffu);
    }
  }
}

// rows [m_base, m_base+32) x columns [n, n+4) of the tile; `stage` holds the warp's 32 accumulator rows with a pitch
// of LDS_ROW floats, and this lane's four columns start `col` floats into each row.  The lane visits rows row_first,
// row_first + row_step, ...: (0, 1) when the 32 lanes span 128 columns of one row at a time, (lane / 8, 4) when a pass
// is 32 columns wide and four rows are in flight per warp instruction (the persistent pair kernel's small staging).
template <int EPI, int ACT, bool VEC, int LDS_ROW>
__device__ __forceinline__ void epilogue_rows(const GemmArgs& g, uint32_t stage, int col, int m_base, int n, int cnt,
                                              float (&cs)[4], int kz, int row_first, int row_step) {
  constexpr bool NEED_AUX = ACT != HF_ACT_NONE && EPI >= EPI_BIAS_DACT;
  // Every field is copied into a register first: `g` lives in the kernel-parameter window and is read through a
  // generic pointer, which the compiler must otherwise re-load after every global store (possible aliasing).
  const int64_t ldc = g.ldc, ldaux = g.ldaux;
  float* const C = g.C + (g.split_k > 1 ? (int64_t)kz * g.M * g.ldc : 0) + n;
  float* const C2 = g.C2 ? g.C2 + n : nullptr;
  const float* const aux = g.aux ? g.aux + n : nullptr;
  const float* const hga = g.h_ga ? g.h_ga + n : nullptr;
  const float* const hrz = g.h_rz ? g.h_rz + n : nullptr;
  const Image16 img = g.c_img;
  const float alpha = g.alpha;
  const int rows = min(32, g.M - m_base);
  float bi[4] = {0.f, 0.f, 0.f, 0.f};
  if ((EPI == EPI_STORE || EPI == EPI_BIAS_ACT || EPI == EPI_BIAS_DACT) && g.bias) ld4<VEC>(g.bias + n, cnt, bi);
  constexpr int RB = 4;  // rows in flight: their global loads are all issued before the first use
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // column sums in registers (`cs` is memory: it crosses a call boundary)
#pragma unroll 1
  for (int r0 = 0; r0 * row_step + row_first < rows; r0 += RB) {
    float x[RB][4], au[RB][4], ga[RB][4], rz[RB][4];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int r = min((r0 + j) * row_step + row_first, rows - 1);  // clamp: tail slots re-read the last row and are not stored
      const int64_t m = m_base + r;
      float4 t;  // explicit shared-space load (the pointer's address space is not visible to the compiler here)
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(stage + (r * LDS_ROW + col) * 4));
      x[j][0] = t.x, x[j][1] = t.y, x[j][2] = t.z, x[j][3] = t.w;
      if (NEED_AUX) ld4<VEC>(aux + m * ldaux, cnt, au[j]);
      if (EPI == EPI_DACT_H && hga) {
        ld4<VEC>(hga + m * ldaux, cnt, ga[j]);
        ld4<VEC>(hrz + m * ldaux, cnt, rz[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      if ((r0 + j) * row_step + row_first >= rows) break;
      const int64_t m = m_base + (r0 + j) * row_step + row_first;
      if (EPI == EPI_STORE) {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[j][e] = alpha * x[j][e] + bi[e];
      } else if (EPI == EPI_BIAS_ACT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[j][e] = act_fwd<ACT>(x[j][e] + bi[e]);
      } else if (EPI == EPI_BIAS_DACT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[j][e] += bi[e];
        if (C2) st4<VEC>(C2 + m * ldc, cnt, x[j]);
        if (NEED_AUX)
#pragma unroll
          for (int e = 0; e < 4; ++e) x[j][e] *= d1<ACT>(au[j][e]);
#pragma unroll
        for (int e = 0; e < 4; ++e) x[j][e] *= alpha;  // 1, or the scale of a fused diagonal loss Hessian
      } else if (EPI == EPI_DACT) {
        if (C2) st4<VEC>(C2 + m * ldc, cnt, x[j]);
        if (NEED_AUX)
#pragma unroll
          for (int e = 0; e < 4; ++e) x[j][e] *= d1<ACT>(au[j][e]);
      } else {  // EPI_DACT_H
        if (NEED_AUX) {
#pragma unroll
          for (int e = 0; e < 4; ++e) x[j][e] *= d1<ACT>(au[j][e]);
          if (hga)
#pragma unroll
            for (int e = 0; e < 4; ++e) x[j][e] += ga[j][e] * d2<ACT>(au[j][e]) * rz[j][e];
        }
      }
      st4<VEC>(C + m * ldc, cnt, x[j]);
      if (img.hi) st4_image(img, m, n, cnt, x[j]);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += x[j][e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) cs[e] += acc[e];
}

template <int EPI, int ACT, int LDS_ROW>
__device__ __forceinline__ void epilogue_vec(const GemmArgs& g, uint32_t stage, int col, int m_base, int n, float (&cs)[4], int kz, int row_first, int row_step) {
  const int cnt = min(4, g.N - n);
  if (cnt <= 0 || m_base >= g.M) return;
  float* C = g.C + (g.split_k > 1 ? (int64_t)kz * g.M * g.ldc : 0);
  const bool al16 = ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(g.C2) | reinterpret_cast<uintptr_t>(g.aux) |
                      reinterpret_cast<uintptr_t>(g.h_ga) | reinterpret_cast<uintptr_t>(g.h_rz) |
                      reinterpret_cast<uintptr_t>(g.bias)) & 15u) == 0 && g.ldc % 4 == 0 && g.ldaux % 4 == 0;
  if (al16 && cnt == 4)
    epilogue_rows<EPI, ACT, true, LDS_ROW>(g, stage, col, m_base, n, cnt, cs, kz, row_first, row_step);
  else
    epilogue_rows<EPI, ACT, false, LDS_ROW>(g, stage, col, m_base, n, cnt, cs, kz, row_first, row_step);
}

template <int EPI, int LDS_ROW>
__device__ __forceinline__ void epilogue_act(const GemmArgs& g, uint32_t stage, int col, int m_base, int n, float (&cs)[4], int kz, int row_first, int row_step) {
  switch (g.act) {
    case HF_ACT_RELU: epilogue_vec<EPI, HF_ACT_RELU, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
    case HF_ACT_SIGMOID: epilogue_vec<EPI, HF_ACT_SIGMOID, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
    case HF_ACT_TANH: epilogue_vec<EPI, HF_ACT_TANH, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
    default: epilogue_vec<EPI, HF_ACT_NONE, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
  }
}

// Fused epilogue of one warp: its 32 accumulator rows (staged in shared memory at `stage`, pitch LDS_ROW floats), this
// lane's four columns `col`..`col`+3 of the staged rows = global columns n..n+3.  Adds the column sums of what the lane
// stored to cs (bias gradient of the layer below).
template <int LDS_ROW>
__device__ __forceinline__ void epilogue_dispatch(const GemmArgs& g, uint32_t stage, int col, int m_base, int n, float (&cs)[4], int kz, int row_first = 0,
                                                  int row_step = 1) {
  switch (g.epi) {
    case EPI_STORE: epilogue_vec<EPI_STORE, HF_ACT_NONE, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
    case EPI_BIAS_ACT: epilogue_act<EPI_BIAS_ACT, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
    case EPI_BIAS_DACT: epilogue_act<EPI_BIAS_DACT, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
    case EPI_DACT: epilogue_act<EPI_DACT, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
    default: epilogue_act<EPI_DACT_H, LDS_ROW>(g, stage, col, m_base, n, cs, kz, row_first, row_step); break;
  }
}

// ---- host side, shared ---------------------------------------------------------------------------------
// cached cuTensorMapEncodeTiled: the operands of a solve are the same buffers on every CG iteration, so after the
// first product every launch finds its maps here instead of encoding four to twelve of them on the host
const CUtensorMap* cached_tensor_map(CUtensorMapDataType dtype, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                                     uint32_t box0, uint32_t box1, CUtensorMapSwizzle swizzle);

}  // namespace hf
