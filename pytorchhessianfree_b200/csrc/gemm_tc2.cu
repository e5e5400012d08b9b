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
constexpr int P2_THREADS = 192;
constexpr int P2_LDS_ROW = 128 + 4;             // staged accumulator rows, 128 columns per pass: 4 warps x 32 rows x 132 floats = 66 KB

// Two builds of the same kernel, by k-block width BK (floats of K per pipeline stage):
//   BK = 32: 64 KB per stage, 193 KB per CTA, ONE CTA per SM (the default).  Three stages hide the L2 latency of a lone
//            CTA: with L2-resident operands the main loop runs at the MMA floor (0.50-0.60 us per k-block,
//            tools/tc2_slope.py).  Everything outside the loop is exposed: per 128x256 CTA tile ~3.3 us of prologue and
//            pipeline fill and ~7.4 us of epilogue (tools/tc2_trace.py).
//   BK = 16: 32 KB per stage, 97 KB per CTA, TWO CTAs per SM (two pairs per TPC, 2 x 256 TMEM columns); HF_TC2_BK=16.
//            Parity-green; meant to run one pair's epilogue under the other pair's main loop, but co-resident pairs of one
//            launch run in lockstep, so it measures the same as BK = 32.  Kept for the persistent variant that staggers them.
template <int BK>
struct P2Cfg {
  static constexpr int OP32 = P2_ROWS * BK * 4;   // FP32 tile of one operand
  static constexpr int OP16 = P2_ROWS * BK * 2;   // one BF16 plane
  static constexpr int A32 = 0, AHI = OP32, ALO = OP32 + OP16, B32 = OP32 + 2 * OP16, BHI = B32 + OP32, BLO = BHI + OP16;
  static constexpr int STAGE_BYTES = 2 * (OP32 + 2 * OP16);  // 64 KB / 32 KB
  static constexpr int SMEM = P2_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int CTAS_PER_SM = BK == 32 ? 1 : 2;
  static_assert(4 * 32 * P2_LDS_ROW * 4 + 4 * P2_TILE * 4 <= P2_STAGES * STAGE_BYTES, "epilogue staging must fit in the idle ring");
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
__global__ void __launch_bounds__(P2_THREADS, P2Cfg<BK>::CTAS_PER_SM)
gemm_tc2_kernel(const __grid_constant__ Tc2Maps maps, const __grid_constant__ Tc2Args p) {
  using Cfg = P2Cfg<BK>;
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
    // ---------------- epilogue (both CTAs: 128 rows x 256 columns each) ----------------
    // TMEM -> registers (one accumulator row per lane) -> shared (the ring is idle once acc_full fired: every MMA of
    // the pair has retired and every TMA box was consumed) -> row-wise coalesced fused epilogue; 128 columns per pass.
    if (total > 0) {
      mbar_wait(acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    tc2_mark(3, threadIdx.x == 0);
    const uint32_t stage = smem_u32(tiles) + warp * 32 * P2_LDS_ROW * 4;
    float cs[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};  // column sums of what this lane stores
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll 2
      for (int c = 0; c < 128; c += 16) {
        float v[16];
        if (total > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + h * 128 + c, v);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + (lane * P2_LDS_ROW + c + j) * 4), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
      }
      __syncwarp();
      if (h == 0) tc2_mark(4, threadIdx.x == 0);
      epilogue_dispatch<P2_LDS_ROW>(g, stage, lane * 4, m0 + warp * 32, n0 + h * 128 + lane * 4, cs[h]);
      __syncwarp();  // the warp's staging rows are rewritten by the next pass
      tc2_mark(5 + h, threadIdx.x == 0);
    }
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
  tc2_mark(7, threadIdx.x == 0);
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

// k-block width of a launch.  BK = 32 (one CTA per SM) unless HF_TC2_BK=16 asks for the co-resident build: measured on
// B200 (tools/tc2_trace.py, 7500x1000x784) the two give the same time, 53 vs 55 us -- co-resident pairs start together,
// share the tensor pipe at half rate each, reach their epilogues together, and nothing overlaps.
int tc2_block_k(int, int, int) {
  static const int forced = getenv("HF_TC2_BK") ? atoi(getenv("HF_TC2_BK")) : 0;
  return forced == 16 ? 16 : 32;
}

int launch_gemm_tc2(const GemmArgs& g_in, cudaStream_t stream) {
  HF_REQUIRE(tc2_supported(g_in), HF_ERR_UNSUPPORTED, "pre-split tcgen05 engine: missing operand image, unsupported shape or alignment");
  Tc2Args p;
  p.g = g_in;
  GemmArgs& g = p.g;
  if (g.split_k < 1) g.split_k = 1;
  if (g.split_k == 1) g.k_per_split = ((g.K + BKT - 1) / BKT) * BKT;
  HF_REQUIRE(g.k_per_split % BKT == 0, HF_ERR_INVALID, "tcgen05 engine: K split must be a multiple of %d", BKT);
  const int bk = tc2_block_k(g.M, g.N, g.split_k);
  Tc2Maps maps;
  for (int s = 0; s < 2; ++s) {
    const int src = s < g.n_pairs ? s : 0;
    p.a_mn[s] = g.A[src].s_k != 1, p.b_mn[s] = g.B[src].s_k != 1;
    int rc = operand_maps(maps.m[s][0], g.A[src], g.M, g.K, bk);
    if (rc) return rc;
    rc = operand_maps(maps.m[s][1], g.B[src], g.N, g.K, bk);
    if (rc) return rc;
  }
  const dim3 grid(2 * ((g.M + P2_TILE - 1) / P2_TILE), (g.N + P2_TILE - 1) / P2_TILE, g.split_k);
  return bk == 32 ? launch_bk<32>(maps, p, grid, stream) : launch_bk<16>(maps, p, grid, stream);
}

}  // namespace hf
