// tcgen05 / TMEM / TMA contraction engine in split precision: TF32 main term + two BF16 correction terms (sm_100a).
//
//   C[M,N] = epilogue( sum_{s < n_pairs} A_s[M,K] * B_s[N,K]^T )        (same contract as gemm_simt.cuh)
//
// One CTA computes one 128x128 output tile (optionally one K-split of it):
//   warp 4   TMA producer: cp.async.bulk.tensor 2-D boxes with SWIZZLE_128B into a 4-slot ring of raw FP32 tiles.  K-contiguous
//            operands arrive as one [128 rows x 32 floats] box (K-major UMMA layout); MN-contiguous operands
//            (transposed use: d^T a, d W) arrive as four [32 k-rows x 32 floats] boxes (MN-major UMMA layout), so no
//            transposed copy of any activation or weight is ever made.  Out-of-bounds rows/columns are zero-filled
//            by the TMA unit, which is what makes ragged M, N, K (784 = 24.5 x 32) legal.
//   warps0-3 splitters, then epilogue.  kind::tf32 keeps the top 19 bits of each FP32 word (truncation), so the raw
//            tile already is the "hi" operand x_t.  The splitters turn each raw tile into two BF16 tiles of the same
//            major-ness in the canonical 16-bit UMMA layouts: lo16 = bf16(x - x_t) (the remainder is exact in FP32)
//            and hi16 = bf16(x).
//   warp 5   MMA issuer (one elected lane): per 32-wide k-block four tcgen05.mma.kind::tf32 (A_t*B_t, K = 8 each) and
//            four tcgen05.mma.kind::f16 (A_lo16*B_hi16 + A_hi16*B_lo16, K = 16 each) into the same TMEM accumulator.
//            Error budget per product: the dropped lo*lo term is 2^-22; the BF16 rounding of a correction operand is
//            2^-9 of a term that is itself 2^-11 of the product, i.e. ~2^-19 with random sign -- FP32-faithful at the
//            rtol 1e-4 the path needs (~2^-13).  All-TF32 corrections (3xTF32) measured the same accuracy in the
//            parity suite and 1.2x the main-loop time: the loop is bound by shared-memory bytes (TMA writes, splitter
//            read/write, MMA operand reads), and the BF16 tiles halve the correction terms' operand bytes.
//            tcgen05.commit releases the raw slot and the BF16 slot (a 2-slot ring of its own).
//   TS       gemm_tc_kernel<1, true> (HF_TC_TS=1): the splitters put the A operand (TF32 words, BF16 lo/hi pairs) into
//            TMEM with tcgen05.st and the MMAs run in TS mode, so the instruction fetches only B from shared memory.
//            Parity-green on the contraction suite; same speed for K-major A, 7-10 % faster when A is MN-major
//            (weight-gradient shapes).  Opt-in until it has been through the whole parity suite.
//   pair     gemm_tc_kernel<2, false>: the same roles in a cluster of two CTAs along M with tcgen05.mma.cta_group::2 (M = 256),
//            each CTA staging its A tile and half of B.  Parity-green, slower on B200, opt-in (HF_TC_PAIR=1).
//   epilogue tcgen05.ld 32x32b -> registers -> the same fused epilogues as the SIMT engine (bias, act', act'', raw
//            copy, split-K partials) -> global.
#include <stdlib.h>
#include <string.h>

#include <unordered_map>

#include "tc_common.cuh"

namespace hf {

#ifndef HF_TC_ITER_TRACE
#define HF_TC_ITER_TRACE 0  // compile with -DHF_TC_ITER_TRACE=1 for tools/tc_pipeline_trace.py (costs ~10 % of the loop)
#endif
#ifndef HF_TC_RS1
#define HF_TC_RS1 4
#define HF_TC_LS1 2
#endif

// Tile 128x128, k-block of BKT = 32 floats (128-byte swizzle rows), 193 KB of shared memory, one CTA per SM.
constexpr int BM = 128, BN = 128;
constexpr int TILE_BYTES = BM * BKT * 4;                  // 16 KB per 128-row FP32 operand tile
constexpr int TC_THREADS = 192;
// NCTA = 1: one CTA per tile.  NCTA = 2: a CTA pair (cluster of two along M, tcgen05 cta_group::2) computes two
// vertically adjacent tiles with one M = 256 MMA: each CTA stages its own A tile and HALF of the shared B tile, so
// per k-block a CTA moves 120 KB instead of 160 KB through its shared memory (the main loop's bound) and the
// smaller stages make room for a fourth ring slot.
template <int NCTA>
struct TcCfg {
  // Two rings with their own barriers: raw FP32 tiles (TMA -> splitters + TF32 MMAs) and BF16 tiles (splitters ->
  // BF16 MMAs).  Depths measured on B200 (us per k-block, M=4096 N=512 K=2048): 4 raw + 2 BF16 slots 0.64, 3 + 3
  // 0.66; issuing the TF32 term of block i ahead of the BF16 terms of block i-1 0.76 (the single issuing thread then
  // waits on the two rings in a fixed order).
  static constexpr int RAW_STAGES = NCTA == 1 ? HF_TC_RS1 : 4;
  static constexpr int LO_STAGES = NCTA == 1 ? HF_TC_LS1 : 4;
  static constexpr int B_ROWS = BN / NCTA;                     // rows of B this CTA stages
  static constexpr int B_BYTES = TILE_BYTES / NCTA;
  static constexpr int OFF_RAW_B = TILE_BYTES;                 // raw slot:  rawA | rawB
  static constexpr int RAW_BYTES = TILE_BYTES + B_BYTES;       // 32 KB / 24 KB
  static constexpr int OFF_B16 = TILE_BYTES;                   // BF16 slot: loA16 hiA16 | loB16 hiB16
  static constexpr int LO_BYTES = TILE_BYTES + B_BYTES;        // 32 KB / 24 KB
  static constexpr int RING_BYTES = RAW_STAGES * RAW_BYTES + LO_STAGES * LO_BYTES;  // 192 KB
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// optional phase trace (tools/tc_trace.py): per CTA, %globaltimer at [0] entry [1] prologue done [2] first stage landed
// [3] accumulator complete [4] epilogue done
// consecutive traced launches write consecutive blocks of the buffer (g_tc_epoch, bumped by the host per launch), so
// the gap between two kernels of a chain can be read off as well
__device__ unsigned long long* g_tc_trace = nullptr;
__device__ __forceinline__ void tc_mark(int slot, bool who, int epoch = 0) {
  if (g_tc_trace && who) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    g_tc_trace[((size_t)epoch * 1024 + cta) * 8 + slot] = t;  // 1024 CTA records per traced launch
  }
}

// per-k-block pipeline trace of CTA (0,0,0) (tools/tc_pipeline_trace.py): [it][0] TMA issued [1] raw seen by the
// splitters [2] split done [3] MMAs issued [4] slot wait of the producer done
__device__ unsigned long long* g_tc_trace_it = nullptr;
__device__ __forceinline__ void tc_mark_it(unsigned long long* buf, int slot, int it) {  // buf: read once per thread
  if (HF_TC_ITER_TRACE && buf && it < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    buf[it * 8 + slot] = t;
  }
}

struct TcArgs {
  GemmArgs g;
  int a_mn[2], b_mn[2];  // 1 = operand is MN-contiguous in global memory (MN-major UMMA operand)
  int trace_epoch;       // block of the trace buffer this launch writes (0 unless tracing)
};

constexpr int kTsAcols = 64;  // TMEM columns of one A slot: 32 TF32 words | 16 packed lo16 | 16 packed hi16

// Row `m` (= threadIdx.x = TMEM lane) of the raw FP32 A tile -> TMEM: x_t words, bf16(x - x_t) and bf16(x) packed in pairs.
template <bool MN>
__device__ __forceinline__ void split_row_to_tmem(uint32_t src, uint32_t taddr, int m) {
  float x[BKT];
  if (!MN) {  // [128 rows][8 x 16 B], chunk ^= row & 7
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(x[4 * j]), "=f"(x[4 * j + 1]), "=f"(x[4 * j + 2]), "=f"(x[4 * j + 3])
                   : "r"(src + m * 128 + ((j ^ (m & 7)) << 4)));
  } else {    // column blocks of 32 rows: [32 k-rows][4 x 32 B], 32-byte chunk ^= k-row & 3
    const uint32_t base = src + (m >> 5) * 4096 + ((m & 7) << 2);
    const int c = (m & 31) >> 3;
#pragma unroll
    for (int k = 0; k < BKT; ++k)
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[k]) : "r"(base + k * 128 + ((c ^ (k & 3)) << 5)));
  }
  uint32_t w[16];
#pragma unroll
  for (int h = 0; h < 2; ++h) {  // TF32 operand: the raw words (the MMA unit truncates)
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(x[16 * h + i]);
    tmem_st16(taddr + 16 * h, w);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float l0 = x[2 * i] - __uint_as_float(__float_as_uint(x[2 * i]) & 0xffffe000u);
    const float l1 = x[2 * i + 1] - __uint_as_float(__float_as_uint(x[2 * i + 1]) & 0xffffe000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(l1), "f"(l0));
  }
  tmem_st16(taddr + 32, w);
#pragma unroll
  for (int i = 0; i < 16; ++i) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
  tmem_st16(taddr + 48, w);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}


// Splitter: one FP32 operand tile (in its UMMA/TMA swizzled layout) -> two BF16 tiles of the same major-ness:
// hi16 = bf16(x) and lo16 = bf16(x - trunc_tf32(x)).  The FP32 tile is element (row, k-quad) addressed by undoing the
// TMA swizzle; the BF16 tiles are written in the canonical UMMA layout for 16-bit operands.
template <bool MN, int ROWS>
__device__ __forceinline__ void split_tile(uint32_t src, uint32_t lo16, uint32_t hi16) {  // shared-space addresses
  // element i = tid + 128 t (t = 0..7) of the FP32 tile is the float4 at src + 16 i; all swizzle arithmetic depends on
  // tid only and is hoisted, the t-dependence is a compile-time constant
  const int tid = threadIdx.x, r0 = tid >> 3, pc = tid & 7;
  int off0;
  if (!MN) {  // source [128 rows][8 x 16 B], chunk ^= row & 7 (SWIZZLE_128B); row = r0 + 16 t
    const int j = pc ^ (r0 & 7);
    off0 = r0 * 64 + ((((j >> 1) ^ (r0 >> 1)) & 3) << 4) + ((j & 1) << 3);
  } else {    // source: 4 column blocks of [32 k-rows][4 x 32 B], 32-byte chunk ^= k-row & 3 (SWIZZLE_128B_ATOM_32B);
              // column block = t >> 1, k-row = r0 + 16 (t & 1)
    const int m0 = (((pc >> 1) ^ r0) & 3) * 8 + (pc & 1) * 4;
    off0 = r0 * 128 + ((((m0 >> 3) ^ r0) & 7) << 4) + ((m0 & 4) << 1);
  }
  constexpr int NT = ROWS / 16;  // float4 per thread: ROWS x 8 quads over 128 threads
  float4 v[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t)
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[t].x), "=f"(v[t].y), "=f"(v[t].z), "=f"(v[t].w) : "r"(src + 16 * tid + 2048 * t));
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const float lx = v[t].x - __uint_as_float(__float_as_uint(v[t].x) & 0xffffe000u);
    const float ly = v[t].y - __uint_as_float(__float_as_uint(v[t].y) & 0xffffe000u);
    const float lz = v[t].z - __uint_as_float(__float_as_uint(v[t].z) & 0xffffe000u);
    const float lw = v[t].w - __uint_as_float(__float_as_uint(v[t].w) & 0xffffe000u);
    uint32_t h0, h1, l0, l1;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(v[t].y), "f"(v[t].x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(v[t].w), "f"(v[t].z));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l0) : "f"(ly), "f"(lx));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l1) : "f"(lw), "f"(lz));
    const int off = MN ? (t >> 2) * 4096 + (t & 1) * 2048 + (off0 ^ (((t >> 1) & 1) << 6)) : off0 + 1024 * t;
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(lo16 + off), "r"(l0), "r"(l1) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(hi16 + off), "r"(h0), "r"(h1) : "memory");
  }
}


template <int NCTA, bool TS>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ TcArgs p) {
  using Cfg = TcCfg<NCTA>;
  constexpr int RS = Cfg::RAW_STAGES, LS = Cfg::LO_STAGES;
  static_assert(!(TS && NCTA == 2), "the TMEM-operand variant is single-CTA");
  // columns: accumulator [0,128) (+ TS: LS slots of A)
  constexpr int TMEM_COLS = TS ? 256 : BN;
  const GemmArgs& g = p.g;
  if (g.skip && *g.skip) return;  // uniform (also across a CTA pair): solver already terminated
  // pair: rank 0 (even tile row) is the leader and issues the M = 256 MMAs for both CTAs
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint8_t* lo_tiles = tiles + RS * Cfg::RAW_BYTES;
  uint64_t* bars = (uint64_t*)(tiles + Cfg::RING_BYTES);
  uint64_t* full_raw = bars;                  // [RS] TMA bytes landed
  uint64_t* empty_raw = bars + RS;            // [RS] MMAs reading the raw slot retired
  uint64_t* full_lo = bars + 2 * RS;          // [LS] splitters done
  uint64_t* empty_lo = bars + 2 * RS + LS;    // [LS] MMAs reading the BF16 slot retired
  uint64_t* acc_full = bars + 2 * RS + 2 * LS; // accumulator complete
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* const trace_it = (HF_TC_ITER_TRACE && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) ? g_tc_trace_it : nullptr;
  tc_mark(0, threadIdx.x == 0, p.trace_epoch);
  // pair: the two CTAs of a cluster must be neighbours in x, so the pair kernel runs tile rows along x
  const int tile_m = NCTA == 2 ? blockIdx.x : blockIdx.y, tile_n = NCTA == 2 ? blockIdx.y : blockIdx.x;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int nb0 = n0 + (int)cta_rank * Cfg::B_ROWS;  // first row of B this CTA stages
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int n_kb = k_end > k_begin ? (k_end - k_begin + BKT - 1) / BKT : 0;
  const int total = n_kb * g.n_pairs;

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < RS; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&empty_raw[s], 1);
    }
    for (int s = 0; s < LS; ++s) {
      mbar_init(&full_lo[s], 4 * NCTA);  // the leader's barrier collects the splitter warps of both CTAs
      mbar_init(&empty_lo[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    if (NCTA == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (NCTA == 2) cluster_sync_all();  // the peer's barriers exist before anything arrives on them remotely
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  tc_mark(1, threadIdx.x == 0, p.trace_epoch);

  if (warp == 4) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % RS, ph = (it / RS) & 1;
        const int pr = it / n_kb, k0 = k_begin + (it % n_kb) * BKT;
        mbar_wait(&empty_raw[s], ph ^ 1);
        tc_mark_it(trace_it, 4, it);
        mbar_expect_tx(&full_raw[s], TILE_BYTES + Cfg::B_BYTES);
        uint8_t* rawA = tiles + s * Cfg::RAW_BYTES;
        uint8_t* rawB = rawA + Cfg::OFF_RAW_B;
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
          for (int j = 0; j < Cfg::B_ROWS / 32; ++j) tma_load_2d(rawB + j * (BKT * 128), mb, nb0 + 32 * j, k0, &full_raw[s]);
        } else {
          tma_load_2d(rawB, mb, k0, nb0, &full_raw[s]);  // box of B_ROWS rows (the B maps are encoded per NCTA)
        }
        tc_mark_it(trace_it, 0, it);
      }
    }
  } else if (warp == 5) {
    // ---------------- MMA issuer (pair: the leader CTA only) ----------------
    if (lane == 0 && cta_rank == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a/b format [7,10)/[10,13) (TF32 = 2,
      // BF16 = 1), majors 15/16, N>>3 [17,23), M>>4 [24,29)
      auto idesc_of = [&](int pr, uint32_t fmt) { return umma_idesc(fmt, p.a_mn[pr], p.b_mn[pr], BM * NCTA, BN); };
      for (int it = 0; it < total; ++it) {
        const int s = it % RS, ph = (it / RS) & 1, pr = it / n_kb;
        const int l = it % LS, phl = (it / LS) & 1;
        mbar_wait(&full_raw[s], ph);
        // pair: the peer's splitters arrive on the leader's barrier after they saw the peer's own TMA bytes land, so
        // this wait covers the peer's raw and BF16 tiles as well
        if (NCTA == 2) mbar_wait_cluster(&full_lo[l], phl);
        else mbar_wait(&full_lo[l], phl);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t rawA = smem_u32(tiles + s * Cfg::RAW_BYTES), rawB = rawA + Cfg::OFF_RAW_B;
        // corrections in BF16 (half the tensor time of a TF32 MMA): A_lo*B_hi + A_hi*B_lo with 8-bit operands is
        // 2^-10 * 2^-8 = 2^-18 relative per product, random sign
        const uint32_t idesc = idesc_of(pr, 2u), idesc16 = idesc_of(pr, 1u);
        const uint32_t loA = smem_u32(lo_tiles + l * Cfg::LO_BYTES), hiA = loA + TILE_BYTES / 2;
        const uint32_t loB = loA + Cfg::OFF_B16, hiB = loB + Cfg::B_BYTES / 2;
        if (TS) {
          // A from TMEM: always "K-major" (lane = row, columns = K), whatever the layout of A in global memory
          const uint32_t clr = ~(1u << 15);
          const uint32_t a_t = tmem_base + BN + l * kTsAcols;
#pragma unroll
          for (int ks = 0; ks < BKT / 8; ++ks)
            umma_ts_tf32(tmem_base, a_t + 8 * ks, operand_desc(rawB, p.b_mn[pr], ks), idesc & clr, (it | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < BKT / 16; ++ks) {
            umma_ts_bf16(tmem_base, a_t + 32 + 8 * ks, corr_desc(hiB, p.b_mn[pr], ks), idesc16 & clr, 1);
            umma_ts_bf16(tmem_base, a_t + 48 + 8 * ks, corr_desc(loB, p.b_mn[pr], ks), idesc16 & clr, 1);
          }
          umma_commit(&empty_raw[s]);
          umma_commit(&empty_lo[l]);
        } else if (NCTA == 1) {
#pragma unroll
          for (int ks = 0; ks < BKT / 8; ++ks)
            umma_tf32(tmem_base, operand_desc(rawA, p.a_mn[pr], ks), operand_desc(rawB, p.b_mn[pr], ks), idesc, (it | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < BKT / 16; ++ks) {
            umma_bf16(tmem_base, corr_desc(loA, p.a_mn[pr], ks), corr_desc(hiB, p.b_mn[pr], ks), idesc16, 1);
            umma_bf16(tmem_base, corr_desc(hiA, p.a_mn[pr], ks), corr_desc(loB, p.b_mn[pr], ks), idesc16, 1);
          }
          umma_commit(&empty_raw[s]);
          umma_commit(&empty_lo[l]);
          tc_mark_it(trace_it, 3, it);
        } else {
#pragma unroll
          for (int ks = 0; ks < BKT / 8; ++ks)
            umma2_tf32(tmem_base, operand_desc(rawA, p.a_mn[pr], ks), operand_desc(rawB, p.b_mn[pr], ks), idesc, (it | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < BKT / 16; ++ks) {
            umma2_bf16(tmem_base, corr_desc(loA, p.a_mn[pr], ks), corr_desc(hiB, p.b_mn[pr], ks), idesc16, 1);
            umma2_bf16(tmem_base, corr_desc(hiA, p.a_mn[pr], ks), corr_desc(loB, p.b_mn[pr], ks), idesc16, 1);
          }
          umma2_commit(&empty_raw[s]);  // frees the slots in both CTAs
          umma2_commit(&empty_lo[l]);
        }
      }
      if (NCTA == 1) umma_commit(acc_full);
      else umma2_commit(acc_full);
    }
  } else {
    // ---------------- splitters ----------------
    int sp_pr = 0, sp_kb = 0;
    for (int it = 0; it < total; ++it) {
      const int s = it % RS, ph = (it / RS) & 1;
      const int l = it % LS, phl = (it / LS) & 1;
      mbar_wait(&empty_lo[l], phl ^ 1);
      mbar_wait(&full_raw[s], ph);
      if (it == 0) tc_mark(2, threadIdx.x == 0, p.trace_epoch);
      if (threadIdx.x == 0) tc_mark_it(trace_it, 1, it);
      const uint32_t st = smem_u32(tiles + s * Cfg::RAW_BYTES), lo = smem_u32(lo_tiles + l * Cfg::LO_BYTES);
      if (TS) {
        const uint32_t a_t = tmem_base + ((uint32_t)(warp * 32) << 16) + BN + l * kTsAcols;
        if (p.a_mn[sp_pr]) split_row_to_tmem<true>(st, a_t, threadIdx.x);
        else split_row_to_tmem<false>(st, a_t, threadIdx.x);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      } else if (p.a_mn[sp_pr]) split_tile<true, BM>(st, lo, lo + TILE_BYTES / 2);
      else split_tile<false, BM>(st, lo, lo + TILE_BYTES / 2);
      if (p.b_mn[sp_pr])
        split_tile<true, Cfg::B_ROWS>(st + Cfg::OFF_RAW_B, lo + Cfg::OFF_B16, lo + Cfg::OFF_B16 + Cfg::B_BYTES / 2);
      else
        split_tile<false, Cfg::B_ROWS>(st + Cfg::OFF_RAW_B, lo + Cfg::OFF_B16, lo + Cfg::OFF_B16 + Cfg::B_BYTES / 2);
      if (++sp_kb == n_kb) sp_kb = 0, ++sp_pr;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA unit
      __syncwarp();
      if (threadIdx.x == 0) tc_mark_it(trace_it, 2, it);
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_cluster(&full_lo[l], 0);
        else mbar_arrive(&full_lo[l]);
      }
    }
    // ---------------- epilogue ----------------
    // TMEM -> registers (one accumulator row per lane) -> shared (the pipeline buffers are idle once acc_full fired)
    // -> row-wise, so that every global load/store of the fused epilogue is a coalesced 512-byte row segment.
    if (total > 0) {
      mbar_wait(acc_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    tc_mark(3, threadIdx.x == 0, p.trace_epoch);
    constexpr int LDS_ROW = BN + 4;
    const uint32_t stage = smem_u32(tiles) + warp * 32 * LDS_ROW * 4;
#pragma unroll 2
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
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + (lane * LDS_ROW + c + j) * 4), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
    }
    __syncwarp();
    tc_mark(5, threadIdx.x == 0, p.trace_epoch);
    float cs[4] = {0.f, 0.f, 0.f, 0.f};  // column sums of what this lane stores (bias gradient of the next layer)
    epilogue_dispatch<LDS_ROW>(g, stage, lane * 4, m0 + warp * 32, n0 + lane * 4, cs, blockIdx.z);
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
          if (n + e < g.N) g.colpart[(int64_t)tile_m * g.N + n + e] = t;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (NCTA == 2) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair is still in flight
  else __syncthreads();
  tc_mark(4, threadIdx.x == 0, p.trace_epoch);
  if (warp == 5) {
    if (NCTA == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
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

// Tensor maps are cached by content: the operands of a solve are the same buffers on every CG iteration, so after
// the first product a launch finds its maps here instead of paying cuTensorMapEncodeTiled four to twelve times.
struct MapKey {
  uint64_t w[6];
  bool operator==(const MapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    for (uint64_t v : k.w) h = (h ^ v) * 0xbf58476d1ce4e5b9ull, h ^= h >> 29;
    return (size_t)h;
  }
};

const CUtensorMap* cached_tensor_map(CUtensorMapDataType dtype, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                                     uint32_t box0, uint32_t box1, CUtensorMapSwizzle swizzle) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key = {{(uint64_t)(uintptr_t)ptr, dim0, dim1, stride1_bytes, ((uint64_t)box0 << 32) | box1, ((uint64_t)dtype << 32) | (uint64_t)swizzle}};
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  if (cache.size() > 8192) cache.clear();  // pointers of long-dead buffers: start over rather than grow without bound
  CUtensorMap map;
  const cuuint64_t dims[2] = {dim0, dim1}, strides[1] = {stride1_bytes};
  const cuuint32_t box[2] = {box0, box1}, estr[2] = {1, 1};
  CUresult r = encode_fn()(&map, dtype, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (ptr %p dims %llu x %llu stride %llu box %u x %u)", (int)r, ptr,
              (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)stride1_bytes, box0, box1);
    return nullptr;
  }
  return &cache.emplace(key, map).first->second;
}

// 2-D tensor map over one FP32 operand.  K-contiguous: dims (K, MN), box (32, rows).  MN-contiguous: dims (MN, K), box (32, 32).
static int make_map(CUtensorMap* map, const Operand& op, int MN, int K, int box_rows = BM) {
  const bool mn_major = op.s_k != 1;
  const CUtensorMap* m =
      mn_major ? cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, op.ptr, (uint64_t)MN, (uint64_t)K, (uint64_t)op.s_k * 4, 32, BKT,
                                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
               : cached_tensor_map(CU_TENSOR_MAP_DATA_TYPE_FLOAT32, op.ptr, (uint64_t)K, (uint64_t)MN, (uint64_t)op.s_mn * 4, BKT,
                                   (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!m) return HF_ERR_CUDA;
  *map = *m;
  return HF_OK;
}

static int g_trace_launches = -1;  // >= 0 while tracing: index of the next traced launch
int set_tc_trace(void* d_buf) {
  g_trace_launches = d_buf ? 0 : -1;
  unsigned long long* p = static_cast<unsigned long long*>(d_buf);
  HF_CUDA(cudaMemcpyToSymbol(g_tc_trace, &p, sizeof(p)));
  return HF_OK;
}

int set_tc_trace_iters(void* d_buf) {
  unsigned long long* p = static_cast<unsigned long long*>(d_buf);
  HF_CUDA(cudaMemcpyToSymbol(g_tc_trace_it, &p, sizeof(p)));
  return HF_OK;
}

int launch_gemm_tc(const GemmArgs& g_in, cudaStream_t stream) {
  HF_REQUIRE(tc_supported(g_in), HF_ERR_UNSUPPORTED, "tcgen05 engine: unsupported shape or alignment");
  TcArgs p;
  p.g = g_in;
  p.trace_epoch = g_trace_launches >= 0 ? g_trace_launches++ : 0;
  GemmArgs& g = p.g;
  if (g.split_k < 1) g.split_k = 1;
  if (g.split_k == 1) g.k_per_split = ((g.K + BKT - 1) / BKT) * BKT;
  HF_REQUIRE(g.k_per_split % BKT == 0, HF_ERR_INVALID, "tcgen05 engine: K split must be a multiple of %d", BKT);
  // CTA pairs need an even number of tile rows (a cluster of two along M).  Off unless HF_TC_PAIR=1: parity-green, but
  // measured slower than single-CTA tiles on B200 (the loop is latency-bound and the pair adds cluster round trips)
  const dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.split_k);
  static const bool pair_allowed = getenv("HF_TC_PAIR") && atoi(getenv("HF_TC_PAIR")) != 0;
  const bool pair = pair_allowed && grid.y % 2 == 0;
  CUtensorMap maps[4];
  for (int s = 0; s < 2; ++s) {
    const int src = s < g.n_pairs ? s : 0;
    p.a_mn[s] = g.A[src].s_k != 1, p.b_mn[s] = g.B[src].s_k != 1;
    int rc = make_map(&maps[2 * s], g.A[src], g.M, g.K);
    if (rc) return rc;
    rc = make_map(&maps[2 * s + 1], g.B[src], g.N, g.K, pair ? BN / 2 : BN);
    if (rc) return rc;
  }
  static bool seen[64] = {};
  if (first_use_on_device(seen)) {  // the opt-in is per device: a process may drive several GPUs
    HF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<1>::SMEM_BYTES));
    HF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<1>::SMEM_BYTES));
    HF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<2>::SMEM_BYTES));
  }
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid.y, grid.x, grid.z), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = TcCfg<2>::SMEM_BYTES, cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2, false>, maps[0], maps[1], maps[2], maps[3], p);
    if (le != cudaSuccess) {
      int nclusters = -1;
      cudaGetLastError();
      cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, gemm_tc_kernel<2, false>, &cfg);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, gemm_tc_kernel<2, false>);
      HF_REQUIRE(false, HF_ERR_CUDA, "pair launch failed: %s; grid (%u,%u,%u) smem %d; max active clusters %d (%s); regs %d static smem %zu maxdyn %d",
                 cudaGetErrorString(le), grid.x, grid.y, grid.z, TcCfg<2>::SMEM_BYTES, nclusters, cudaGetErrorString(oe), fa.numRegs,
                 fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
    }
    note_launch();
  } else {
    // HF_TC_TS=1: A operand through TMEM (experimental)
    static const bool ts = getenv("HF_TC_TS") && atoi(getenv("HF_TC_TS")) != 0;
    if (ts) gemm_tc_kernel<1, true><<<grid, TC_THREADS, TcCfg<1>::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    else gemm_tc_kernel<1, false><<<grid, TC_THREADS, TcCfg<1>::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    HF_LAUNCH_CHECK();
  }
  return HF_OK;
}

}  // namespace hf
