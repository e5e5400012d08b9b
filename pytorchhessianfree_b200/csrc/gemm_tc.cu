// tcgen05 / TMEM / TMA contraction engine in 3xTF32 split precision (sm_100a).
//
//   C[M,N] = epilogue( sum_{s < n_pairs} A_s[M,K] * B_s[N,K]^T )        (same contract as gemm_simt.cuh)
//
// One CTA computes one 128x128 output tile (optionally one K-split of it):
//   warp 4   TMA producer: cp.async.bulk.tensor 2-D boxes with SWIZZLE_128B into a 3-stage ring.  K-contiguous
//            operands arrive as one [128 rows x 32 floats] box (K-major UMMA layout); MN-contiguous operands
//            (transposed use: d^T a, d W) arrive as four [32 k-rows x 32 floats] boxes (MN-major UMMA layout), so no
//            transposed copy of any activation or weight is ever made.  Out-of-bounds rows/columns are zero-filled
//            by the TMA unit, which is what makes ragged M, N, K (784 = 24.5 x 32) legal.
//   warps0-3 splitters, then epilogue.  kind::tf32 keeps the top 19 bits of each FP32 word (truncation), so the raw
//            tile already is the "hi" operand; the splitters write lo = x - trunc_tf32(x) (exact in FP32) into a
//            second buffer with the identical swizzled layout (the op is element-wise, so the swizzle is irrelevant).
//   warp 5   MMA issuer (one elected lane): per 8-wide k-step three tcgen05.mma.kind::tf32 into the same TMEM
//            accumulator: A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.  The dropped lo*lo term is ~2^-22 relative, so
//            products are FP32-faithful (rtol 1e-4 needs ~2^-13).  tcgen05.commit releases the smem stage.
//   epilogue tcgen05.ld 32x32b -> registers -> the same fused epilogues as the SIMT engine (bias, act', act'', raw
//            copy, split-K partials) -> global.
#include <cuda.h>

#include "gemm_tc.cuh"

namespace hf {

constexpr int BM = 128, BN = 128, BKT = 32;  // tile; BKT floats = 128 B = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BKT * 4;                  // 16 KB per operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;               // rawA | rawB | loA | loB
constexpr int TC_THREADS = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TcArgs {
  GemmArgs g;
  int a_mn[2], b_mn[2];  // 1 = operand is MN-contiguous in global memory (MN-major UMMA operand)
};

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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
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

// UMMA shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor: start[0,14) lbo[16,30) sbo[32,46)
// version[46,48)=1 layout_type[61,64)=2); all offsets in 16-byte units.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// operand tile of 128 (M or N) x 32 (K) floats at `base`, k-step ks (8 floats):
//   K-major : SWIZZLE_128B (type 2): rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO); step = +32 B in the row
//   MN-major: 32-bit operands only exist in SWIZZLE_128B_BASE32B (type 1; cute Layout_MN_SW128_32B_Atom, TMA
//             SWIZZLE_128B_ATOM_32B): atoms of [4 k-rows x 128 B] 512 B apart along K (SBO); 4 column blocks of
//             [32 k-rows x 128 B] 4096 B apart along MN (LBO); one k-step = 8 k-rows = +1024 B
__device__ __forceinline__ uint64_t operand_desc(uint32_t base, int mn_major, int ks) {
  return mn_major ? smem_desc(base + ks * 1024, BKT * 128, 512, 1) : smem_desc(base + ks * 32, 16, 1024, 2);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, const TcArgs p) {
  const GemmArgs& g = p.g;
  if (g.skip && *g.skip) return;  // uniform: solver already terminated
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(tiles + STAGES * STAGE_BYTES);
  uint64_t* full_raw = bars;              // [STAGES] TMA bytes landed
  uint64_t* full_lo = bars + STAGES;      // [STAGES] splitters done
  uint64_t* empty = bars + 2 * STAGES;    // [STAGES] MMAs of the stage retired
  uint64_t* acc_full = bars + 3 * STAGES; // accumulator complete
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int n_kb = k_end > k_begin ? (k_end - k_begin + BKT - 1) / BKT : 0;
  const int total = n_kb * g.n_pairs;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&full_lo[s], 4);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        const int pr = it / n_kb, k0 = k_begin + (it % n_kb) * BKT;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full_raw[s], 2 * TILE_BYTES);
        uint8_t* rawA = tiles + s * STAGE_BYTES;
        uint8_t* rawB = rawA + TILE_BYTES;
        const CUtensorMap* ma = pr ? &tmA1 : &tmA0;
        const CUtensorMap* mb = pr ? &tmB1 : &tmB0;
        if (p.a_mn[pr]) {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(rawA + j * (BKT * 128), ma, m0 + 32 * j, k0, &full_raw[s]);
        } else {
          tma_load_2d(rawA, ma, k0, m0, &full_raw[s]);
        }
        if (p.b_mn[pr]) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(rawB + j * (BKT * 128), mb, n0 + 32 * j, k0, &full_raw[s]);
        } else {
          tma_load_2d(rawB, mb, k0, n0, &full_raw[s]);
        }
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1, pr = it / n_kb;
        // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a/b=TF32 [7,10)/[10,13), majors 15/16,
        // N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn[pr] << 15) |
                               ((uint32_t)p.b_mn[pr] << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        mbar_wait(&full_raw[s], ph);
        mbar_wait(&full_lo[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t rawA = smem_u32(tiles + s * STAGE_BYTES), rawB = rawA + TILE_BYTES;
        const uint32_t loA = rawA + 2 * TILE_BYTES, loB = rawA + 3 * TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BKT / 8; ++ks) {
          const uint64_t a_hi = operand_desc(rawA, p.a_mn[pr], ks), b_hi = operand_desc(rawB, p.b_mn[pr], ks);
          const uint64_t a_lo = operand_desc(loA, p.a_mn[pr], ks), b_lo = operand_desc(loB, p.b_mn[pr], ks);
          umma_tf32(tmem_base, a_lo, b_hi, idesc, (it | ks) != 0);
          umma_tf32(tmem_base, a_hi, b_lo, idesc, 1);
          umma_tf32(tmem_base, a_hi, b_hi, idesc, 1);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ---------------- splitters ----------------
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(&full_raw[s], ph);
      const float4* src = reinterpret_cast<const float4*>(tiles + s * STAGE_BYTES);
      float4* dst = reinterpret_cast<float4*>(tiles + s * STAGE_BYTES + 2 * TILE_BYTES);
#pragma unroll 4
      for (int i = threadIdx.x; i < 2 * TILE_BYTES / 16; i += 128) {
        const float4 v = src[i];
        float4 lo;
        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        dst[i] = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA unit
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_lo[s]);
    }
    // ---------------- epilogue ----------------
    // TMEM -> registers (one accumulator row per lane) -> shared (the pipeline buffers are idle once acc_full fired)
    // -> row-wise, so that every global load/store of the fused epilogue is a coalesced 512-byte row segment.
    if (total > 0) {
      mbar_wait(acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    constexpr int LDS_ROW = BN + 4;
    float* stage = reinterpret_cast<float*>(tiles) + warp * 32 * LDS_ROW;
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      if (total > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(stage + lane * LDS_ROW + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    __syncwarp();
    float* C = g.C + (g.split_k > 1 ? (int64_t)blockIdx.z * g.M * g.ldc : 0);
    const bool al16 = ((reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(g.C2) | reinterpret_cast<uintptr_t>(g.aux) |
                        reinterpret_cast<uintptr_t>(g.h_ga) | reinterpret_cast<uintptr_t>(g.h_rz) |
                        reinterpret_cast<uintptr_t>(g.bias)) & 15u) == 0 && g.ldc % 4 == 0 && g.ldaux % 4 == 0;
    const int n = n0 + lane * 4;
    const int cnt = min(4, g.N - n);
    const bool vec = al16 && cnt == 4;
    auto load4 = [&](const float* p, float (&o)[4]) {
      if (vec) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        o[0] = t.x, o[1] = t.y, o[2] = t.z, o[3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = e < cnt ? p[e] : 0.f;
      }
    };
    auto store4 = [&](float* p, const float (&o)[4]) {
      if (vec) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e < cnt) p[e] = o[e];
      }
    };
    float cs[4] = {0.f, 0.f, 0.f, 0.f};  // column sums of what this lane stores (bias gradient of the next layer)
    if (cnt > 0) {
      float bi[4] = {0.f, 0.f, 0.f, 0.f};
      if (g.bias) load4(g.bias + n, bi);
      const bool need_aux = g.act != HF_ACT_NONE && g.epi >= EPI_BIAS_DACT;
      constexpr int RB = 8;  // rows in flight: all their global loads are issued before the first use
      for (int r0 = 0; r0 < 32; r0 += RB) {
        float aub[RB][4];
        if (need_aux) {
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            const int mj = m0 + warp * 32 + r0 + j;
            if (mj < g.M) load4(g.aux + (int64_t)mj * g.ldaux + n, aub[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int r = r0 + j;
          const int m = m0 + warp * 32 + r;
          if (m >= g.M) break;
          float x[4], au[4] = {0.f, 0.f, 0.f, 0.f};
          const float4 t = *reinterpret_cast<const float4*>(stage + r * LDS_ROW + lane * 4);
          x[0] = t.x, x[1] = t.y, x[2] = t.z, x[3] = t.w;
          if (need_aux) {
#pragma unroll
            for (int e = 0; e < 4; ++e) au[e] = aub[j][e];
          }
          switch (g.epi) {
            case EPI_STORE:
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] = g.alpha * x[e] + bi[e];
              break;
            case EPI_BIAS_ACT:
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] = act_apply(g.act, x[e] + bi[e]);
              break;
            case EPI_BIAS_DACT:
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] += bi[e];
              if (g.C2) store4(g.C2 + (int64_t)m * g.ldc + n, x);
              if (need_aux)
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] *= act_d1(g.act, au[e]);
              break;
            case EPI_DACT:
              if (g.C2) store4(g.C2 + (int64_t)m * g.ldc + n, x);
              if (need_aux)
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] *= act_d1(g.act, au[e]);
              break;
            case EPI_DACT_H: {
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] *= act_d1(g.act, au[e]);
              if (g.h_ga) {
                float ga[4], rz[4];
                load4(g.h_ga + (int64_t)m * g.ldaux + n, ga);
                load4(g.h_rz + (int64_t)m * g.ldaux + n, rz);
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] += ga[e] * act_d2(g.act, au[e]) * rz[e];
              }
            } break;
          }
          store4(C + (int64_t)m * g.ldc + n, x);
#pragma unroll
          for (int e = 0; e < 4; ++e) cs[e] += x[e];
        }
      }
    }
    if (g.colpart) {
      // 4 warps x 32 rows -> one row of column sums per CTA, fixed order (deterministic)
      float* red = reinterpret_cast<float*>(tiles) + 4 * 32 * (BN + 4);
      *reinterpret_cast<float4*>(red + warp * BN + lane * 4) = make_float4(cs[0], cs[1], cs[2], cs[3]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 0) {
        const int n = n0 + lane * 4;
  #pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int idx = lane * 4 + e;
          const float t = (red[idx] + red[BN + idx]) + (red[2 * BN + idx] + red[3 * BN + idx]);
          if (n + e < g.N) g.colpart[(int64_t)blockIdx.y * g.N + n + e] = t;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

static bool operand_ok(const Operand& op, int MN, int K) {
  if (!op.ptr || (reinterpret_cast<uintptr_t>(op.ptr) & 15u)) return false;
  if (op.s_k == 1) return op.s_mn % 4 == 0 && op.s_mn >= K;
  if (op.s_mn == 1) return op.s_k % 4 == 0 && op.s_k >= MN;
  return false;
}

bool tc_supported(const GemmArgs& g) {
  if (g.square || g.n_pairs < 1 || g.n_pairs > 2) return false;
  if ((int64_t)g.M * g.N * g.K < kTcMinWork) return false;  // tiny layers stay on the SIMT tiles
  for (int s = 0; s < g.n_pairs; ++s)
    if (!operand_ok(g.A[s], g.M, g.K) || !operand_ok(g.B[s], g.N, g.K)) return false;
  return encode_fn() != nullptr;
}

// 2-D tensor map over one operand.  K-contiguous: dims (K, MN), box (32, 128).  MN-contiguous: dims (MN, K), box (32, 32).
static int make_map(CUtensorMap* map, const Operand& op, int MN, int K) {
  const bool mn_major = op.s_k != 1;
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (mn_major) {
    dims[0] = (cuuint64_t)MN, dims[1] = (cuuint64_t)K, strides[0] = (cuuint64_t)op.s_k * 4;
    box[0] = 32, box[1] = BKT;
  } else {
    dims[0] = (cuuint64_t)K, dims[1] = (cuuint64_t)MN, strides[0] = (cuuint64_t)op.s_mn * 4;
    box[0] = BKT, box[1] = BM;
  }
  CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(op.ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HF_REQUIRE(r == CUDA_SUCCESS, HF_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return HF_OK;
}

int launch_gemm_tc(const GemmArgs& g_in, cudaStream_t stream) {
  HF_REQUIRE(tc_supported(g_in), HF_ERR_UNSUPPORTED, "tcgen05 engine: unsupported shape or alignment");
  TcArgs p;
  p.g = g_in;
  GemmArgs& g = p.g;
  if (g.split_k < 1) g.split_k = 1;
  if (g.split_k == 1) g.k_per_split = ((g.K + BKT - 1) / BKT) * BKT;
  HF_REQUIRE(g.k_per_split % BKT == 0, HF_ERR_INVALID, "tcgen05 engine: K split must be a multiple of %d", BKT);
  CUtensorMap maps[4];
  for (int s = 0; s < 2; ++s) {
    const int src = s < g.n_pairs ? s : 0;
    p.a_mn[s] = g.A[src].s_k != 1, p.b_mn[s] = g.B[src].s_k != 1;
    int rc = make_map(&maps[2 * s], g.A[src], g.M, g.K);
    if (rc) return rc;
    rc = make_map(&maps[2 * s + 1], g.B[src], g.N, g.K);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    HF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.split_k);
  gemm_tc_kernel<<<grid, TC_THREADS, SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  HF_LAUNCH_CHECK();
  return HF_OK;
}

}  // namespace hf
