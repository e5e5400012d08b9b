// tcgen05 contraction engine on PRE-SPLIT operand images: 256x256 output tiles computed by CTA pairs (sm_100a).
//
//   C[M,N] = epilogue( sum_{s < n_pairs} A_s[M,K] * B_s[N,K]^T )        (same contract as gemm_simt.cuh / gemm_tc.cu)
//
// Why a second tile kernel.  Measured on B200 (tools/mma_rate.cu, profiles/r2_mma_rate.md): a tcgen05.mma with M = 128
// per CTA costs ~130 cycles whether N is 128 or 256, so the 128x128 instructions of gemm_tc.cu run the tensor pipe at
// half rate no matter how the operands are fed; N = 256 doubles the work per instruction at the same cost.  A 128x256
// tile per CTA needs 96 KB of operands per 32-wide k-block, which neither fits a useful ring in 227 KB of shared
// memory nor the L2 -> SM feed (~165 KB/us per SM measured); a CTA PAIR (tcgen05 cta_group::2, M = 256) stages, per
// CTA, its own 128 rows of A and HALF of the 256 rows of B: 64 KB per k-block, three stages.
//
// Why pre-split.  The split-precision product  A B ~= A_t B_t (kind::tf32) + A_lo B_hi + A_hi B_lo (kind::f16)  needs
// three forms of every operand (gemm_simt.cuh: Image16).  gemm_tc.cu derives the BF16 forms in the main loop, on every
// k-block of every launch, with four warps -- for operands that are constant over a whole 50..250-iteration solve
// (inputs, activations, weights).  Here the BF16 planes are read from HBM/L2 ready-made: written once per
// linearisation for the constant operands, by the producing kernel's epilogue for the per-iteration ones (tangents,
// cotangents) and by one small kernel per product for the CG direction.  The main loop is then TMA + MMA only:
//   warp 4   TMA producer (both CTAs): per k-block six operand forms (A32 Ahi Alo B32 Bhi Blo), each CTA into its own
//            shared memory, all counted on the LEADER's mbarrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 5   MMA issuer (leader CTA, one lane): 4 x tcgen05.mma.cta_group::2.kind::tf32 + 4 x kind::f16 per k-block,
//            M = 256, N = 256, into one FP32 accumulator of 256 TMEM columns per CTA; tcgen05.commit (multicast)
//            frees the stage in both CTAs
//   warps0-3 epilogue only: tcgen05.ld -> shared -> row-wise fused epilogue (tc_common.cuh), which also stores the
//            split image of C when the next contraction will read it
// K-contiguous and MN-contiguous operands are both loaded in place (K-major / MN-major UMMA layouts, as in
// gemm_tc.cu); ragged M, N, K are zero-filled by the TMA unit.
#include <stdlib.h>

#include <algorithm>

#include "tc_common.cuh"

namespace hf {

constexpr int P2_ROWS = 128;                    // rows of A, and of B, one CTA stages
constexpr int P2_TILE = 256;                    // pair tile: 256 x 256
constexpr int P2_STAGES = 3;
constexpr int P2_THREADS = 320;                 // one-tile kernel: 8 epilogue warps + producer + MMA issuer
constexpr int P2P_THREADS = 192;                // persistent kernel: 4 epilogue warps + producer + MMA issuer
constexpr int P2_LDS_ROW = 128 + 4;             // staged accumulator rows, 128 columns per pass: 4 warps x 32 rows x 132 floats = 66 KB

// Stage geometry by k-block width BK (floats of K per pipeline stage).  The one-tile kernel runs BK = 32: 64 KB per
// stage, 193 KB per CTA, ONE CTA per SM; three stages hide the L2 latency of a lone CTA: with L2-resident operands the
// main loop runs at the MMA floor (0.50-0.60 us per k-block, tools/tc2_slope.py).  (A BK = 16 build with two CTAs per
// SM -- two pairs per TPC, 2 x 256 TMEM columns -- measured the same, 55 vs 53 us: co-resident pairs of one launch run
// in lockstep, so nothing overlapped; removed.)  BK = 16 survives as an option of the persistent kernel below.
template <int BK>
struct P2Cfg {
  static constexpr int OP32 = P2_ROWS * BK * 4;   // FP32 tile of one operand
  static constexpr int OP16 = P2_ROWS * BK * 2;   // one BF16 plane
  static constexpr int A32 = 0, AHI = OP32, ALO = OP32 + OP16, B32 = OP32 + 2 * OP16, BHI = B32 + OP32, BLO = BHI + OP16;
  static constexpr int STAGE_BYTES = 2 * (OP32 + 2 * OP16);  // 64 KB / 32 KB
  static constexpr int SMEM = P2_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// UMMA descriptors of the operand tiles of one stage, k-step ks (one MMA: 8 floats / 16 bf16 of K).
//   FP32 K-major : rows of BK*4 bytes, SWIZZLE_128B (BK = 32) / SWIZZLE_64B (BK = 16), 8-row atoms (SBO), step +32 B
//   FP32 MN-major: SWIZZLE_128B_BASE32B, column blocks of [BK k-rows x 128 B] (LBO), atoms of 4 k-rows (SBO 512),
//                  step = 8 k-rows = +1024 B
//   BF16 K-major : rows of BK*2 bytes, SWIZZLE_64B (BK = 32, step +32 B) / SWIZZLE_32B (BK = 16, a single step)
//   BF16 MN-major: SWIZZLE_128B, column blocks of [BK k-rows x 128 B = 64 bf16] (LBO), atoms of 8 k-rows (SBO 1024),
//                  step = 16 k-rows = +2048 B
template <int BK>
__device__ __forceinline__ uint64_t desc32(uint32_t base, int mn_major, int ks) {
  return mn_major ? smem_desc(base + ks * 1024, BK * 128, 512, 1) : smem_desc(base + ks * 32, 16, 8 * BK * 4, BK == 32 ? 2u : 4u);
}
template <int BK>
__device__ __forceinline__ uint64_t desc16(uint32_t base, int mn_major, int ks) {
  return mn_major ? smem_desc(base + ks * 2048, BK * 128, 1024, 2) : smem_desc(base + ks * 32, 16, 8 * BK * 2, BK == 32 ? 4u : 6u);
}

// optional phase trace (tools/tc2_trace.py): per CTA, %globaltimer at [0] entry [1] prologue done [2] first stage
// landed (leader's MMA thread) [3] accumulator complete [4] first 128 columns staged [5] first 128 columns stored
// [6] all stored [7] pair released
__device__ unsigned long long* g_tc2_trace = nullptr;
__device__ __forceinline__ void tc2_mark(int slot, bool who) {
  if (g_tc2_trace && who) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (cta < 4096) g_tc2_trace[(size_t)cta * 8 + slot] = t;
  }
}

struct Tc2Maps {
  CUtensorMap m[2][2][3];  // [pair][A | B][fp32 | hi | lo]
};
struct Tc2Args {
  GemmArgs g;
  int a_mn[2], b_mn[2];  // 1 = operand is MN-contiguous in global memory (MN-major UMMA operand)
};

// one operand form set of 128 rows at k0: FP32 tile + the two BF16 planes
template <int BK>
__device__ __forceinline__ void load_operand(uint32_t dst32, uint32_t dst_hi, uint32_t dst_lo, const CUtensorMap* maps, int mn_major,
                                             int row0, int k0, uint32_t bar) {
  if (mn_major) {
#pragma unroll
    for (int j = 0; j < P2_ROWS / 32; ++j) tma_load_2d_pair(dst32 + j * (BK * 128), &maps[0], row0 + 32 * j, k0, bar);
#pragma unroll
    for (int j = 0; j < P2_ROWS / 64; ++j) {
      tma_load_2d_pair(dst_hi + j * (BK * 128), &maps[1], row0 + 64 * j, k0, bar);
      tma_load_2d_pair(dst_lo + j * (BK * 128), &maps[2], row0 + 64 * j, k0, bar);
    }
  } else {
    tma_load_2d_pair(dst32, &maps[0], k0, row0, bar);
    tma_load_2d_pair(dst_hi, &maps[1], k0, row0, bar);
    tma_load_2d_pair(dst_lo, &maps[2], k0, row0, bar);
  }
}

template <int BK>
__global__ void __launch_bounds__(P2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ Tc2Maps maps, const __grid_constant__ Tc2Args p) {
  using Cfg = P2Cfg<BK>;
  static_assert(8 * 32 * P2_LDS_ROW * 4 + 4 * P2_TILE * 4 <= P2_STAGES * Cfg::STAGE_BYTES, "epilogue staging must fit in the idle ring");
  const GemmArgs& g = p.g;
  if (g.skip && *g.skip) return;  // uniform across the pair: solver already terminated
  const uint32_t rank = cluster_ctarank();  // 0 = leader: issues the MMAs for both CTAs
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(tiles + P2_STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES] leader's: bytes of BOTH CTAs landed
  uint64_t* empty = bars + P2_STAGES;       // [STAGES] MMAs reading the stage retired (commit multicast: both CTAs)
  uint64_t* acc_full = bars + 2 * P2_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  tc2_mark(0, threadIdx.x == 0);
  // the two CTAs of a cluster are neighbours in x: blockIdx.x counts 128-row blocks, blockIdx.y 256-column tiles
  const int m0 = blockIdx.x * P2_ROWS;
  const int n0 = blockIdx.y * P2_TILE;
  const int nb0 = n0 + (int)rank * P2_ROWS;  // first row of B this CTA stages
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int n_kb = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
  const int total = n_kb * g.n_pairs;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < P2_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 12; ++i)  // hide the descriptor fetch of the first loads
      asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i / 6][(i / 3) % 2][i % 3]) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(P2_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // the leader's barriers exist before the peer's TMA counts bytes on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  tc2_mark(1, threadIdx.x == 0);

  if (warp == 4) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      const uint32_t bar0 = map_to_cta(&full[0], 0);  // the leader's barriers, as shared::cluster addresses
      for (int it = 0; it < total; ++it) {
        const int s = it % P2_STAGES, ph = (it / P2_STAGES) & 1;
        const int pr = it / n_kb, k0 = k_begin + (it % n_kb) * BK;
        mbar_wait(&empty[s], ph ^ 1);
        // the leader arms its barrier for the bytes of both CTAs; the peer's complete_tx may arrive first (the phase
        // cannot complete before the leader's own arrival)
        if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
        const uint32_t bar = bar0 + s * 8;
        const uint32_t base = smem_u32(tiles + s * Cfg::STAGE_BYTES);
        load_operand<BK>(base + Cfg::A32, base + Cfg::AHI, base + Cfg::ALO, maps.m[pr][0], p.a_mn[pr], m0, k0, bar);
        load_operand<BK>(base + Cfg::B32, base + Cfg::BHI, base + Cfg::BLO, maps.m[pr][1], p.b_mn[pr], nb0, k0, bar);
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer (leader CTA) ----------------
    if (lane == 0 && rank == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % P2_STAGES, ph = (it / P2_STAGES) & 1, pr = it / n_kb;
        mbar_wait(&full[s], ph);
        if (it == 0) tc2_mark(2, true);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int a_mn = p.a_mn[pr], b_mn = p.b_mn[pr];
        const uint32_t idesc32 = umma_idesc(2u, a_mn, b_mn, P2_TILE, P2_TILE), idesc16 = umma_idesc(1u, a_mn, b_mn, P2_TILE, P2_TILE);
        const uint32_t base = smem_u32(tiles + s * Cfg::STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks)
          umma2_tf32(tmem_base, desc32<BK>(base + Cfg::A32, a_mn, ks), desc32<BK>(base + Cfg::B32, b_mn, ks), idesc32, (it | ks) != 0);
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          umma2_bf16(tmem_base, desc16<BK>(base + Cfg::ALO, a_mn, ks), desc16<BK>(base + Cfg::BHI, b_mn, ks), idesc16, 1);
          umma2_bf16(tmem_base, desc16<BK>(base + Cfg::AHI, a_mn, ks), desc16<BK>(base + Cfg::BLO, b_mn, ks), idesc16, 1);
        }
        umma2_commit(&empty[s]);  // frees the stage in both CTAs
      }
      umma2_commit(acc_full);
    }
  } else {
    // ---------------- epilogue (both CTAs: 128 rows x 256 columns each; warps 0-3 and 6-9) ----------------
    // TMEM -> registers (one accumulator row per lane) -> shared (the ring is idle once acc_full fired: every MMA of
    // the pair has retired and every TMA box was consumed) -> row-wise coalesced fused epilogue.  A warp reads the TMEM
    // lane quarter (warp % 4); two warps share a quarter and take 128 columns each: the chain tcgen05.ld -> st.shared ->
    // ld.shared -> (loads of the fused terms) -> st.global is latency-bound with one warp per scheduler (4 warps: 7.4 us
    // for the 128 KB of a CTA, tools/tc2_trace.py), so the second set of warps nearly halves it.
    if (total > 0) {
      mbar_wait(acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    tc2_mark(3, threadIdx.x == 0);
    const int q = warp & 3, h = warp >= 6 ? 1 : 0;
    const uint32_t stage = smem_u32(tiles) + (h * 4 + q) * 32 * P2_LDS_ROW * 4;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};  // column sums of what this lane stores
#pragma unroll 2
    for (int c = 0; c < 128; c += 16) {
      float v[16];
      if (total > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + h * 128 + c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + (lane * P2_LDS_ROW + c + j) * 4), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
    }
    __syncwarp();
    tc2_mark(4, threadIdx.x == 0);
    epilogue_dispatch<P2_LDS_ROW>(g, stage, lane * 4, m0 + q * 32, n0 + h * 128 + lane * 4, cs, blockIdx.z);
    tc2_mark(5 + h, lane == 0 && q == 0);
    if (g.colpart && m0 < g.M) {  // (the grid is padded to whole pairs: a CTA entirely below the matrix owns no row of colpart)
      // 4 quarters x 32 rows -> one row of column sums per CTA (= per 128-row block, like gemm_tc.cu), fixed order
      float* red = reinterpret_cast<float*>(tiles) + 8 * 32 * P2_LDS_ROW;
      *reinterpret_cast<float4*>(red + q * P2_TILE + h * 128 + lane * 4) = make_float4(cs[0], cs[1], cs[2], cs[3]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int idx = (h * 4 + q) * 32 + lane;  // one column per epilogue thread
      const float t = (red[idx] + red[P2_TILE + idx]) + (red[2 * P2_TILE + idx] + red[3 * P2_TILE + idx]);
      if (n0 + idx < g.N) g.colpart[(int64_t)blockIdx.x * g.N + n0 + idx] = t;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // neither CTA leaves (or frees TMEM, or lets its shared memory go) while the pair is in flight
  tc2_mark(7, threadIdx.x == 0);
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(P2_TILE) : "memory");
}

// ---- persistent variant -------------------------------------------------------------------------------
// Same tiles, same operand forms, same epilogue; what changes is what happens BETWEEN the tiles of a launch that has
// more pair tiles than the machine has pairs (the 7500-row contractions of the autoencoder: 120 tiles on 74 pairs).
// tools/tc2_trace.py: of the 31 us a pair spends on one such tile, 11 us are outside the main loop (prologue 1.3 us,
// pipeline fill 2 us, epilogue 7.4 us during which 148 CTAs store 128 KB each and nothing computes), and the second
// wave pays them again.  Here one pair per SM pair stays resident and walks the tile list:
//   * the accumulator is double-buffered in TMEM (2 x 256 columns = all of it): the MMA warp starts tile i+1 in the
//     other buffer as soon as its operands land, while warps 0-3 drain tile i;
//   * the epilogue stages through its OWN shared memory, not through the idle ring, so the producer keeps prefetching.
//     A first cut kept the 128-column passes (66 KB of staging) and paid for them with the ring (128 KB: 2 x 64 KB or
//     4 x 32 KB): parity-green and SLOWER than the one-tile kernel on two-wave shapes (62.8-69.4 vs 57.2 us on 7500 x
//     1000 x 784) -- the main loop lost more to the shallow ring than the overlap won.  So: passes of 32 columns (18 KB
//     of staging: 4 warps x 32 rows x 36 floats), four accumulator rows in flight per warp instruction (each lane
//     quarter stores 128 contiguous bytes of a row), and the full 192 KB ring of the one-tile kernel;
//   * barriers, TMEM allocation and the cluster rendezvous happen once per launch.
// acc_full[b]  : leader's MMA thread -> epilogue warps of both CTAs (tcgen05.commit, multicast)
// acc_empty[b] : 4 epilogue warps x 2 CTAs -> leader's MMA thread (remote mbarrier arrive), after their last tcgen05.ld
constexpr int P2_COLS_P = 32;  // columns per epilogue pass of the persistent kernel
template <int BK>
struct P2PCfg {
  static constexpr int STAGES = BK == 32 ? 3 : 6;
  static constexpr int STAGE_BYTES = P2Cfg<BK>::STAGE_BYTES;
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int COLS = P2_COLS_P;
  static constexpr int LDS_ROW = COLS + 4;  // 144-byte rows: quarter-warp stores / loads hit 8 distinct 16-byte bank groups
  static constexpr int STAGING = 4 * 32 * LDS_ROW * 4;
  static constexpr int RED = 4 * P2_TILE * 4;
  static constexpr int SMEM = RING + STAGING + RED + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM <= 227 * 1024, "persistent pair kernel: shared memory");
};

// One epilogue pass as a real call: inlined into the tile loop, the 40 specialisations of the row loop have their
// loop-invariant setup hoisted above it and the kernel spills at 255 registers; behind a call it needs 100.
template <int EPI>
__device__ __noinline__ void tc2p_epilogue_kind(const GemmArgs& g, uint32_t stage, int col, int m_base, int n, float* cs_out, int kz, int row_first) {
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  if (EPI == EPI_STORE)
    epilogue_vec<EPI_STORE, HF_ACT_NONE, P2_COLS_P + 4>(g, stage, col, m_base, n, cs, kz, row_first, 4);
  else
    epilogue_act<EPI, P2_COLS_P + 4>(g, stage, col, m_base, n, cs, kz, row_first, 4);
  cs_out[0] = cs[0], cs_out[1] = cs[1], cs_out[2] = cs[2], cs_out[3] = cs[3];
}
__device__ __forceinline__ void tc2p_epilogue_pass(const GemmArgs& g, uint32_t stage, int col, int m_base, int n, float* cs, int kz, int row_first) {
  switch (g.epi) {
    case EPI_STORE: tc2p_epilogue_kind<EPI_STORE>(g, stage, col, m_base, n, cs, kz, row_first); break;
    case EPI_BIAS_ACT: tc2p_epilogue_kind<EPI_BIAS_ACT>(g, stage, col, m_base, n, cs, kz, row_first); break;
    case EPI_BIAS_DACT: tc2p_epilogue_kind<EPI_BIAS_DACT>(g, stage, col, m_base, n, cs, kz, row_first); break;
    case EPI_DACT: tc2p_epilogue_kind<EPI_DACT>(g, stage, col, m_base, n, cs, kz, row_first); break;
    default: tc2p_epilogue_kind<EPI_DACT_H>(g, stage, col, m_base, n, cs, kz, row_first); break;
  }
}

struct Tc2Tiles {
  int tiles_m, tiles_n, n_tiles;  // pair tiles along M, along N, in all (x split_k)
};

template <int BK>
__global__ void __launch_bounds__(P2P_THREADS, 1)
gemm_tc2p_kernel(const __grid_constant__ Tc2Maps maps, const __grid_constant__ Tc2Args p, const __grid_constant__ Tc2Tiles tl) {
  using Cfg = P2Cfg<BK>;
  using PC = P2PCfg<BK>;
  constexpr int S = PC::STAGES;
  const GemmArgs& g = p.g;
  if (g.skip && *g.skip) return;
  const uint32_t rank = cluster_ctarank();
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = tiles + PC::RING;
  float* red = reinterpret_cast<float*>(staging + PC::STAGING);
  uint64_t* bars = (uint64_t*)(staging + PC::STAGING + PC::RED);
  uint64_t* full = bars;             // [S] leader's
  uint64_t* empty = bars + S;        // [S] both (commit multicast)
  uint64_t* acc_full = bars + 2 * S; // [2] both (commit multicast)
  uint64_t* acc_empty = acc_full + 2;  // [2] leader's: 8 arrivals
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  tc2_mark(0, threadIdx.x == 0);

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 12; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i / 6][(i / 3) % 2][i % 3]) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * P2_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  tc2_mark(1, threadIdx.x == 0);

  // tile t -> (pair-row block, column tile, K split); the same order the one-tile kernel's grid is rasterised in
  auto n_kb_of = [&](int z) {
    const int k_begin = z * g.k_per_split, k_end = min(g.K, k_begin + g.k_per_split);
    return k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
  };

  if (warp == 4) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      const uint32_t bar0 = map_to_cta(&full[0], 0);
      int it = 0;
      for (int t = cid; t < tl.n_tiles; t += n_clusters) {
        const int xm = t % tl.tiles_m, yn = (t / tl.tiles_m) % tl.tiles_n, z = t / (tl.tiles_m * tl.tiles_n);
        const int m0 = (2 * xm + (int)rank) * P2_ROWS, nb0 = yn * P2_TILE + (int)rank * P2_ROWS;
        const int k_begin = z * g.k_per_split, n_kb = n_kb_of(z), total = n_kb * g.n_pairs;
        for (int i = 0; i < total; ++i, ++it) {
          const int s = it % S, ph = (it / S) & 1;
          const int pr = i / n_kb, k0 = k_begin + (i % n_kb) * BK;
          mbar_wait(&empty[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
          const uint32_t bar = bar0 + s * 8;
          const uint32_t base = smem_u32(tiles + s * Cfg::STAGE_BYTES);
          load_operand<BK>(base + Cfg::A32, base + Cfg::AHI, base + Cfg::ALO, maps.m[pr][0], p.a_mn[pr], m0, k0, bar);
          load_operand<BK>(base + Cfg::B32, base + Cfg::BHI, base + Cfg::BLO, maps.m[pr][1], p.b_mn[pr], nb0, k0, bar);
        }
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer (leader CTA) ----------------
    if (lane == 0 && rank == 0) {
      int it = 0, lt = 0;
      for (int t = cid; t < tl.n_tiles; t += n_clusters, ++lt) {
        const int z = t / (tl.tiles_m * tl.tiles_n);
        const int n_kb = n_kb_of(z), total = n_kb * g.n_pairs;
        const int b = lt & 1;
        mbar_wait_cluster(&acc_empty[b], ((lt >> 1) & 1) ^ 1);  // both CTAs have drained what this buffer held two tiles ago
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + b * P2_TILE;
        for (int i = 0; i < total; ++i, ++it) {
          const int s = it % S, ph = (it / S) & 1, pr = i / n_kb;
          mbar_wait(&full[s], ph);
          if (it == 0) tc2_mark(2, true);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int a_mn = p.a_mn[pr], b_mn = p.b_mn[pr];
          const uint32_t idesc32 = umma_idesc(2u, a_mn, b_mn, P2_TILE, P2_TILE), idesc16 = umma_idesc(1u, a_mn, b_mn, P2_TILE, P2_TILE);
          const uint32_t base = smem_u32(tiles + s * Cfg::STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks)
            umma2_tf32(acc, desc32<BK>(base + Cfg::A32, a_mn, ks), desc32<BK>(base + Cfg::B32, b_mn, ks), idesc32, (i | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            umma2_bf16(acc, desc16<BK>(base + Cfg::ALO, a_mn, ks), desc16<BK>(base + Cfg::BHI, b_mn, ks), idesc16, 1);
            umma2_bf16(acc, desc16<BK>(base + Cfg::AHI, a_mn, ks), desc16<BK>(base + Cfg::BLO, b_mn, ks), idesc16, 1);
          }
          umma2_commit(&empty[s]);
        }
        umma2_commit(&acc_full[b]);
      }
    }
  } else {
    // ---------------- epilogue (both CTAs: 128 rows x 256 columns of every tile of the pair) ----------------
    constexpr int COLS = PC::COLS, LDS = PC::LDS_ROW, PASSES = P2_TILE / COLS;
    const uint32_t stage = smem_u32(staging) + warp * 32 * LDS * 4;
    const int sub = lane >> 3, col = (lane & 7) * 4;  // row residue (mod 4) and first column of this lane within a pass
    int lt = 0;
    for (int t = cid; t < tl.n_tiles; t += n_clusters, ++lt) {
      const int xm = t % tl.tiles_m, yn = (t / tl.tiles_m) % tl.tiles_n, z = t / (tl.tiles_m * tl.tiles_n);
      const int mb = 2 * xm + (int)rank;  // 128-row block of this CTA
      const int m0 = mb * P2_ROWS, n0 = yn * P2_TILE;
      const int total = n_kb_of(z) * g.n_pairs;
      const int b = lt & 1;
      mbar_wait(&acc_full[b], (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lt == 0) tc2_mark(3, threadIdx.x == 0);
#pragma unroll 1
      for (int h = 0; h < PASSES; ++h) {
        float cs[4];
#pragma unroll
        for (int c = 0; c < COLS; c += 16) {
          float v[16];
          if (total > 0) {
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + b * P2_TILE + h * COLS + c, v);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + (lane * LDS + c + j) * 4), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
        }
        if (h == PASSES - 1) {  // this warp has read the last of buffer b: hand it back before the last stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&acc_empty[b], 0);
        }
        __syncwarp();
        tc2p_epilogue_pass(g, stage, col, m0 + warp * 32, n0 + h * COLS + col, cs, z, sub);
        __syncwarp();
        if (g.colpart) {  // column sums of the warp's 32 rows: the four row residues, fixed order
          float4 v = make_float4(cs[0], cs[1], cs[2], cs[3]);
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o), v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, o), v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
          }
          if (sub == 0) *reinterpret_cast<float4*>(red + warp * P2_TILE + h * COLS + col) = v;
        }
      }
      if (g.colpart) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (m0 < g.M)
          for (int idx = threadIdx.x; idx < P2_TILE; idx += 128) {
            const float tsum = (red[idx] + red[P2_TILE + idx]) + (red[2 * P2_TILE + idx] + red[3 * P2_TILE + idx]);
            if (n0 + idx < g.N) g.colpart[(int64_t)mb * g.N + n0 + idx] = tsum;
          }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // `red` is rewritten by the next tile
      }
      if (lt == 0) tc2_mark(6, threadIdx.x == 0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  tc2_mark(7, threadIdx.x == 0);
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * P2_TILE) : "memory");
}

// ---- split-precision images ----------------------------------------------------------------------------
// img = split(src) for `count` segments in one launch (blockIdx.y = segment): hi = bf16(x), lo = bf16(x - tf32_trunc(x));
// optionally also a 16-byte-pitched FP32 copy (what pitch_rows did in round 1).
__global__ void __launch_bounds__(256) split_segments_kernel(SplitTable t) {
  if (t.skip && *t.skip) return;
  const SplitSegment& s = t.seg[blockIdx.y];
  const int groups = (s.cols + 3) >> 2;  // groups of 4 columns
  const int64_t total = (int64_t)s.rows * groups;
  const bool vec = s.ld_src % 4 == 0 && (reinterpret_cast<uintptr_t>(s.src) & 15u) == 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / groups;
    const int c = (int)(i % groups) * 4, cnt = min(4, s.cols - c);
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = s.src + r * s.ld_src + c;
    if (s.perm_taps > 0) {
      for (int e = 0; e < cnt; ++e) {
        const int j = c + e;  // destination column (tap, channel) <- source column (channel, tap)
        x[e] = s.src[r * s.ld_src + (j % s.perm_cin) * s.perm_taps + j / s.perm_cin];
      }
    } else if (vec && cnt == 4) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
    } else {
      for (int e = 0; e < cnt; ++e) x[e] = src[e];
    }
    if (s.square) {
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] *= x[e];
    }
    if (s.dst32) *reinterpret_cast<float4*>(s.dst32 + r * s.ld32 + c) = make_float4(x[0], x[1], x[2], x[3]);  // pitch % 4 == 0
    if (s.img.hi) st4_image(s.img, r, c, 4, x);  // image pitch % 8 == 0: the padding columns get zeros
  }
}

int launch_split(const SplitTable& t, cudaStream_t stream) {
  if (t.count == 0) return HF_OK;
  int64_t most = 0;
  for (int i = 0; i < t.count; ++i) most = std::max<int64_t>(most, (int64_t)t.seg[i].rows * ((t.seg[i].cols + 3) / 4));
  int64_t blocks = (most + 255) / 256;
  const int64_t cap = std::max(1, 4 * sm_count() / t.count);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  split_segments_kernel<<<dim3((unsigned)blocks, t.count), 256, 0, stream>>>(t);
  HF_LAUNCH_CHECK();
  return HF_OK;
}

// ---- host side ---------------------------------------------------------------------------------------
static bool fp32_ok(const Operand& op, int MN, int K) {
  if (!op.ptr || (reinterpret_cast<uintptr_t>(op.ptr) & 15u)) return false;
  if (op.s_k == 1) return op.s_mn % 4 == 0 && op.s_mn >= K;
  if (op.s_mn == 1) return op.s_k % 4 == 0 && op.s_k >= MN;
  return false;
}
static bool image_ok(const Operand& op, int MN, int K) {
  const Image16& im = op.img;
  if (!im.hi || (reinterpret_cast<uintptr_t>(im.hi) & 15u) || im.ld % 8 != 0 || im.plane % 8 != 0) return false;
  return im.ld >= (op.s_k == 1 ? K : MN);
}

bool tc2_supported(const GemmArgs& g) {
  if (g.square || g.n_pairs < 1 || g.n_pairs > 2) return false;
  for (int s = 0; s < g.n_pairs; ++s)
    if (!fp32_ok(g.A[s], g.M, g.K) || !fp32_ok(g.B[s], g.N, g.K) || !image_ok(g.A[s], g.M, g.K) || !image_ok(g.B[s], g.N, g.K))
      return false;
  return true;
}

// Rough cost model of the two tensor engines (us): waves x k-blocks x measured time per k-block.  The pair kernel
// wins when its 256x256 tiles still fill the machine; small grids stay on the 128x128 tiles.
Tc2Choice tc2_estimate(int M, int N, int K, int n_pairs, int splits) {
  const int64_t kb = ((int64_t)(K + BKT - 1) / BKT + splits - 1) / splits * n_pairs;
  const int64_t pairs = (int64_t)((M + 255) / 256) * ((N + 255) / 256) * splits;
  const int64_t ctas = (int64_t)((M + 127) / 128) * ((N + 127) / 128) * splits;
  const int n_pairs_hw = std::max(1, sm_count() / 2), n_sm = std::max(1, sm_count());
  Tc2Choice c;
  c.us_pair = (double)((pairs + n_pairs_hw - 1) / n_pairs_hw) * (double)kb * 0.60 + 6.0;
  c.us_single = (double)((ctas + n_sm - 1) / n_sm) * (double)kb * 0.64 + 6.0;
  return c;
}

int tc2_mode() {
  static const int mode = getenv("HF_TC2") ? atoi(getenv("HF_TC2")) : 1;  // 0 = off, 1 = where the model says so, 2 = wherever supported
  return mode;
}

static int operand_maps(CUtensorMap* out, const Operand& op, int MN, int K, int bk) {
  const bool mn_major = op.s_k != 1;
  const CUtensorMap* m[3];
  if (mn_major) {
    m[0] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, op.ptr, (uint64_t)MN, (uint64_t)K, (uint64_t)op.s_k * 4, 32, (uint32_t)bk,
                             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    for (int i = 0; i < 2; ++i)
      m[1 + i] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, op.img.hi + i * op.img.plane, (uint64_t)MN, (uint64_t)K,
                                   (uint64_t)op.img.ld * 2, 64, (uint32_t)bk, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    m[0] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, op.ptr, (uint64_t)K, (uint64_t)MN, (uint64_t)op.s_mn * 4, (uint32_t)bk,
                             P2_ROWS, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    for (int i = 0; i < 2; ++i)
      m[1 + i] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, op.img.hi + i * op.img.plane, (uint64_t)K, (uint64_t)MN,
                                   (uint64_t)op.img.ld * 2, (uint32_t)bk, P2_ROWS, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  }
  for (int i = 0; i < 3; ++i) {
    if (!m[i]) return HF_ERR_CUDA;
    out[i] = *m[i];
  }
  return HF_OK;
}

int set_tc2_trace(void* d_buf) {
  unsigned long long* p = static_cast<unsigned long long*>(d_buf);
  HF_CUDA(cudaMemcpyToSymbol(g_tc2_trace, &p, sizeof(p)));
  return HF_OK;
}

template <int BK>
static int launch_bk(const Tc2Maps& maps, const Tc2Args& p, dim3 grid, cudaStream_t stream) {
  static bool seen[64] = {};
  if (first_use_on_device(seen))
    HF_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2Cfg<BK>::SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(P2_THREADS), cfg.dynamicSmemBytes = P2Cfg<BK>::SMEM, cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  HF_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<BK>, maps, p));
  note_launch();
  return HF_OK;
}

template <int BK>
static int launch_persistent(const Tc2Maps& maps, const Tc2Args& p, const Tc2Tiles& tl, int n_clusters, cudaStream_t stream) {
  static bool seen[64] = {};
  if (first_use_on_device(seen))
    HF_CUDA(cudaFuncSetAttribute(gemm_tc2p_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2PCfg<BK>::SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * n_clusters), cfg.blockDim = dim3(P2P_THREADS), cfg.dynamicSmemBytes = P2PCfg<BK>::SMEM, cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  HF_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2p_kernel<BK>, maps, p, tl));
  note_launch();
  return HF_OK;
}

// HF_TC2_PERSIST: 0 = one tile per pair always, 1 (default) = the persistent kernel when a launch has at least four
// waves of pair tiles, 2 = always.  Measured (tools/tc_microbench.py): 60000 x 1000 x 784 (12.7 waves) 395 vs 451 us,
// 8192 x 8192 x 2048 (13.8 waves) 957 vs 1019 us, but 7500 x 1000 x 784 (1.6 waves) 63.9 vs 57.5 us: while the epilogue
// of tile i runs, the main loop of tile i+1 crawls (both live on shared-memory bandwidth: the MMAs read ~90 B/cycle of
// operands, the epilogue stages 256 KB per tile through it), so with two tiles per pair the overlap buys less than
// the narrower passes cost.  HF_TC2P_BK = 16 selects a six-stage / 16-wide ring (slower everywhere measured).
static int tc2_persist_mode() {  // read per launch (tests switch it in-process); a getenv is noise next to a launch
  const char* e = getenv("HF_TC2_PERSIST");
  return e ? atoi(e) : 1;
}
static int tc2p_block_k() {
  const char* e = getenv("HF_TC2P_BK");
  return e && atoi(e) == 16 ? 16 : 32;
}

int launch_gemm_tc2(const GemmArgs& g_in, cudaStream_t stream) {
  HF_REQUIRE(tc2_supported(g_in), HF_ERR_UNSUPPORTED, "pre-split tcgen05 engine: missing operand image, unsupported shape or alignment");
  Tc2Args p;
  p.g = g_in;
  GemmArgs& g = p.g;
  if (g.split_k < 1) g.split_k = 1;
  if (g.split_k == 1) g.k_per_split = ((g.K + BKT - 1) / BKT) * BKT;
  HF_REQUIRE(g.k_per_split % BKT == 0, HF_ERR_INVALID, "tcgen05 engine: K split must be a multiple of %d", BKT);
  Tc2Tiles tl;
  tl.tiles_m = (g.M + P2_TILE - 1) / P2_TILE, tl.tiles_n = (g.N + P2_TILE - 1) / P2_TILE;
  tl.n_tiles = tl.tiles_m * tl.tiles_n * g.split_k;
  const int hw_pairs = std::max(1, sm_count() / 2);
  // (a short main loop cannot hide the persistent kernel's four-warp, 32-column-pass epilogue: such launches are
  // epilogue-bound and the one-tile kernel's eight epilogue warps do better)
  const int kb_per_tile = (g.k_per_split + 31) / 32 * g.n_pairs;
  const int mode = tc2_persist_mode();
  const bool persistent = mode == 2 || (mode == 1 && tl.n_tiles >= 4 * hw_pairs && kb_per_tile >= 16);
  const int bk = persistent ? tc2p_block_k() : 32;
  Tc2Maps maps;
  for (int s = 0; s < 2; ++s) {
    const int src = s < g.n_pairs ? s : 0;
    p.a_mn[s] = g.A[src].s_k != 1, p.b_mn[s] = g.B[src].s_k != 1;
    int rc = operand_maps(maps.m[s][0], g.A[src], g.M, g.K, bk);
    if (rc) return rc;
    rc = operand_maps(maps.m[s][1], g.B[src], g.N, g.K, bk);
    if (rc) return rc;
  }
  if (persistent) {
    const int n_clusters = std::min(tl.n_tiles, hw_pairs);
    return bk == 32 ? launch_persistent<32>(maps, p, tl, n_clusters, stream) : launch_persistent<16>(maps, p, tl, n_clusters, stream);
  }
  const dim3 grid(2 * ((g.M + P2_TILE - 1) / P2_TILE), (g.N + P2_TILE - 1) / P2_TILE, g.split_k);
  return launch_bk<32>(maps, p, grid, stream);
}

}  // namespace hf
