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
constexpr int P2_A32 = 0, P2_AHI = 16384, P2_ALO = 24576, P2_B32 = 32768, P2_BHI = 49152, P2_BLO = 57344;
constexpr int P2_STAGE_BYTES = 65536;
constexpr int P2_SMEM = P2_STAGES * P2_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int P2_THREADS = 192;
constexpr int P2_LDS_ROW = P2_TILE + 4;         // staged accumulator rows: 4 warps x 32 rows x 260 floats = 130 KB

struct Tc2Maps {
  CUtensorMap m[2][2][3];  // [pair][A | B][fp32 | hi | lo]
};
struct Tc2Args {
  GemmArgs g;
  int a_mn[2], b_mn[2];  // 1 = operand is MN-contiguous in global memory (MN-major UMMA operand)
};

// Half of a B operand form set (64 rows at k0), multicast to the CTAs of `mask`: with two pairs per cluster the pairs
// compute vertically adjacent tiles, share the B tile, and each CTA fetches a quarter of it for itself and for its
// counterpart in the other pair -- 96 KB instead of 128 KB per pair and k-block come out of the L2, which is what
// bounds the single-pair kernel on a full chip (profiles/r2_summary.md).
__device__ __forceinline__ void load_half_multicast(uint32_t dst32, uint32_t dst_hi, uint32_t dst_lo, const CUtensorMap* maps, int mn_major,
                                                    int row0, int k0, uint32_t bar, uint16_t mask) {
  if (mn_major) {
#pragma unroll
    for (int j = 0; j < 2; ++j) tma_load_2d_pair_mc(dst32 + j * (BKT * 128), &maps[0], row0 + 32 * j, k0, bar, mask);
    tma_load_2d_pair_mc(dst_hi, &maps[1], row0, k0, bar, mask);
    tma_load_2d_pair_mc(dst_lo, &maps[2], row0, k0, bar, mask);
  } else {
    tma_load_2d_pair_mc(dst32, &maps[0], k0, row0, bar, mask);
    tma_load_2d_pair_mc(dst_hi, &maps[1], k0, row0, bar, mask);
    tma_load_2d_pair_mc(dst_lo, &maps[2], k0, row0, bar, mask);
  }
}

// one operand form set of 128 rows at k0: FP32 tile + the two BF16 planes
__device__ __forceinline__ void load_operand(uint32_t dst32, uint32_t dst_hi, uint32_t dst_lo, const CUtensorMap* maps, int mn_major,
                                             int row0, int k0, uint32_t bar) {
  if (mn_major) {
#pragma unroll
    for (int j = 0; j < P2_ROWS / 32; ++j) tma_load_2d_pair(dst32 + j * (BKT * 128), &maps[0], row0 + 32 * j, k0, bar);
#pragma unroll
    for (int j = 0; j < P2_ROWS / 64; ++j) {
      tma_load_2d_pair(dst_hi + j * (BKT * 128), &maps[1], row0 + 64 * j, k0, bar);
      tma_load_2d_pair(dst_lo + j * (BKT * 128), &maps[2], row0 + 64 * j, k0, bar);
    }
  } else {
    tma_load_2d_pair(dst32, &maps[0], k0, row0, bar);
    tma_load_2d_pair(dst_hi, &maps[1], k0, row0, bar);
    tma_load_2d_pair(dst_lo, &maps[2], k0, row0, bar);
  }
}

// PAIRS = 1: cluster = one CTA pair.  PAIRS = 2: cluster = two pairs on vertically adjacent tiles sharing B by multicast.
template <int PAIRS>
__global__ void __launch_bounds__(P2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ Tc2Maps maps, const __grid_constant__ Tc2Args p) {
  const GemmArgs& g = p.g;
  if (g.skip && *g.skip) return;  // uniform across the cluster: solver already terminated
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1;     // 0 = leader of its pair: issues the MMAs for both CTAs
  const uint32_t pair_id = crank >> 1;  // which pair of the cluster
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(tiles + P2_STAGES * P2_STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES] leader's: bytes of BOTH CTAs landed
  uint64_t* empty = bars + P2_STAGES;       // [STAGES] MMAs reading the stage retired (commit multicast: both CTAs)
  uint64_t* acc_full = bars + 2 * P2_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the two CTAs of a cluster are neighbours in x: blockIdx.x counts 128-row blocks, blockIdx.y 256-column tiles
  const int m0 = blockIdx.x * P2_ROWS;
  const int n0 = blockIdx.y * P2_TILE;
  const int nb0 = n0 + (int)rank * P2_ROWS;  // first row of B this CTA stages
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int n_kb = k_end > k_begin ? (k_end - k_begin + BKT - 1) / BKT : 0;
  const int total = n_kb * g.n_pairs;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < P2_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], PAIRS);  // a stage is refilled by multicast from both pairs: both must have drained it
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(P2_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // the leader's barriers exist before the peer's TMA counts bytes on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % P2_STAGES, ph = (it / P2_STAGES) & 1;
        const int pr = it / n_kb, k0 = k_begin + (it % n_kb) * BKT;
        mbar_wait(&empty[s], ph ^ 1);
        // the leader arms its barrier for the bytes of both CTAs; the peer's complete_tx may arrive first (the phase
        // cannot complete before the leader's own arrival)
        if (rank == 0) mbar_expect_tx(&full[s], 2 * P2_STAGE_BYTES);
        const uint32_t bar = map_to_cta(&full[s], crank & ~1u);  // this pair's leader
        const uint32_t base = smem_u32(tiles + s * P2_STAGE_BYTES);
        load_operand(base + P2_A32, base + P2_AHI, base + P2_ALO, maps.m[pr][0], p.a_mn[pr], m0, k0, bar);
        if (PAIRS == 1) {
          load_operand(base + P2_B32, base + P2_BHI, base + P2_BLO, maps.m[pr][1], p.b_mn[pr], nb0, k0, bar);
        } else {
          // rows [nb0 + 64 pair_id, +64) of B for this CTA and for the CTA of the same pair rank in the other pair
          const uint16_t mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));
          load_half_multicast(base + P2_B32 + pair_id * 8192, base + P2_BHI + pair_id * 4096, base + P2_BLO + pair_id * 4096,
                              maps.m[pr][1], p.b_mn[pr], nb0 + 64 * (int)pair_id, k0, bar, mask);
        }
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer (leader CTA) ----------------
    if (lane == 0 && rank == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % P2_STAGES, ph = (it / P2_STAGES) & 1, pr = it / n_kb;
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int a_mn = p.a_mn[pr], b_mn = p.b_mn[pr];
        const uint32_t idesc32 = umma_idesc(2u, a_mn, b_mn, P2_TILE, P2_TILE), idesc16 = umma_idesc(1u, a_mn, b_mn, P2_TILE, P2_TILE);
        const uint32_t base = smem_u32(tiles + s * P2_STAGE_BYTES);
#pragma unroll
        for (int ks = 0; ks < BKT / 8; ++ks)
          umma2_tf32(tmem_base, operand_desc(base + P2_A32, a_mn, ks), operand_desc(base + P2_B32, b_mn, ks), idesc32, (it | ks) != 0);
#pragma unroll
        for (int ks = 0; ks < BKT / 16; ++ks) {
          umma2_bf16(tmem_base, corr_desc(base + P2_ALO, a_mn, ks), corr_desc(base + P2_BHI, b_mn, ks), idesc16, 1);
          umma2_bf16(tmem_base, corr_desc(base + P2_AHI, a_mn, ks), corr_desc(base + P2_BLO, b_mn, ks), idesc16, 1);
        }
        umma2_commit(&empty[s], (uint16_t)((1u << (2 * PAIRS)) - 1));  // one arrival in every CTA of the cluster
      }
      umma2_commit(acc_full, (uint16_t)(3u << (2 * pair_id)));
    }
  } else {
    // ---------------- epilogue (both CTAs: 128 rows x 256 columns each) ----------------
    // TMEM -> registers (one accumulator row per lane) -> shared (the ring is idle once acc_full fired: every MMA of
    // the pair has retired and every TMA box was consumed) -> row-wise coalesced fused epilogue.
    if (total > 0) {
      mbar_wait(acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t stage = smem_u32(tiles) + warp * 32 * P2_LDS_ROW * 4;
#pragma unroll 2
    for (int c = 0; c < P2_TILE; c += 16) {
      float v[16];
      if (total > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + (lane * P2_LDS_ROW + c + j) * 4), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
    }
    __syncwarp();
    float cs[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};  // column sums of what this lane stores
#pragma unroll
    for (int h = 0; h < 2; ++h)
      epilogue_dispatch<P2_LDS_ROW>(g, stage, h * 128 + lane * 4, m0 + warp * 32, n0 + h * 128 + lane * 4, cs[h]);
    if (g.colpart && m0 < g.M) {  // (the grid is padded to whole pairs: a CTA entirely below the matrix owns no row of colpart)
      // 4 warps x 32 rows -> one row of column sums per CTA (= per 128-row block, like gemm_tc.cu), fixed order
      float* red = reinterpret_cast<float*>(tiles) + 4 * 32 * P2_LDS_ROW;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        *reinterpret_cast<float4*>(red + warp * P2_TILE + h * 128 + lane * 4) = make_float4(cs[h][0], cs[h][1], cs[h][2], cs[h][3]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int idx = threadIdx.x; idx < P2_TILE; idx += 128) {
        const float t = (red[idx] + red[P2_TILE + idx]) + (red[2 * P2_TILE + idx] + red[3 * P2_TILE + idx]);
        if (n0 + idx < g.N) g.colpart[(int64_t)blockIdx.x * g.N + n0 + idx] = t;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // neither CTA leaves (or frees TMEM, or lets its shared memory go) while the pair is in flight
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(P2_TILE) : "memory");
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
    if (vec && cnt == 4) {
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

static int operand_maps(CUtensorMap* out, const Operand& op, int MN, int K, int box_rows) {
  const bool mn_major = op.s_k != 1;
  const CUtensorMap* m[3];
  if (mn_major) {
    m[0] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, op.ptr, (uint64_t)MN, (uint64_t)K, (uint64_t)op.s_k * 4, 32, BKT,
                             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    for (int i = 0; i < 2; ++i)
      m[1 + i] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, op.img.hi + i * op.img.plane, (uint64_t)MN, (uint64_t)K,
                                   (uint64_t)op.img.ld * 2, 64, BKT, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    m[0] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, op.ptr, (uint64_t)K, (uint64_t)MN, (uint64_t)op.s_mn * 4, BKT,
                             (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
    for (int i = 0; i < 2; ++i)
      m[1 + i] = cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, op.img.hi + i * op.img.plane, (uint64_t)K, (uint64_t)MN,
                                   (uint64_t)op.img.ld * 2, BKT, (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_64B);
  }
  for (int i = 0; i < 3; ++i) {
    if (!m[i]) return HF_ERR_CUDA;
    out[i] = *m[i];
  }
  return HF_OK;
}

int launch_gemm_tc2(const GemmArgs& g_in, cudaStream_t stream) {
  HF_REQUIRE(tc2_supported(g_in), HF_ERR_UNSUPPORTED, "pre-split tcgen05 engine: missing operand image, unsupported shape or alignment");
  Tc2Args p;
  p.g = g_in;
  GemmArgs& g = p.g;
  if (g.split_k < 1) g.split_k = 1;
  if (g.split_k == 1) g.k_per_split = ((g.K + BKT - 1) / BKT) * BKT;
  HF_REQUIRE(g.k_per_split % BKT == 0, HF_ERR_INVALID, "tcgen05 engine: K split must be a multiple of %d", BKT);
  // HF_TC2_MC=1: two pairs per cluster with B shared by multicast, when the tile rows pair up without much padding.
  // Parity-green, but measured no faster on B200 (main loop 1.12 vs 0.97 us per k-block on a full chip, 132 instead of
  // 148 SMs usable by 4-CTA clusters of this size): a multicast to <= 4 CTAs does not lower the L2 -> SM traffic that
  // bounds the loop (profiles/r2_summary.md), so single-pair clusters are the default.
  static const bool mc_allowed = getenv("HF_TC2_MC") && atoi(getenv("HF_TC2_MC")) != 0;
  const int tiles_m = (g.M + P2_TILE - 1) / P2_TILE;
  const int pairs = (mc_allowed && (tiles_m % 2 == 0 || tiles_m >= 9)) ? 2 : 1;
  Tc2Maps maps;
  for (int s = 0; s < 2; ++s) {
    const int src = s < g.n_pairs ? s : 0;
    p.a_mn[s] = g.A[src].s_k != 1, p.b_mn[s] = g.B[src].s_k != 1;
    int rc = operand_maps(maps.m[s][0], g.A[src], g.M, g.K, P2_ROWS);
    if (rc) return rc;
    rc = operand_maps(maps.m[s][1], g.B[src], g.N, g.K, P2_ROWS / pairs);  // multicast: each CTA fetches 64 rows of B
    if (rc) return rc;
  }
  static bool seen[64] = {};
  if (first_use_on_device(seen)) {
    HF_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM));
    HF_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs * ((tiles_m + pairs - 1) / pairs), (g.N + P2_TILE - 1) / P2_TILE, g.split_k);
  cfg.blockDim = dim3(P2_THREADS), cfg.dynamicSmemBytes = P2_SMEM, cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2 * pairs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  if (pairs == 2) HF_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<2>, maps, p));
  else HF_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<1>, maps, p));
  note_launch();
  return HF_OK;
}

}  // namespace hf
