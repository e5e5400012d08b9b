// FP32 SIMT contraction tiles with fused epilogues (any shape, any alignment).
//
// C[M,N] = epilogue( sum_{s < n_pairs} A_s[M,K] * B_s[N,K]^T )
//
// Operands are addressed by element strides, so the three contractions of a curvature product map
// onto one kernel without transposed copies:
//   forward / R-op   z  = a W^T            A=(in,1)   B=(in,1)        (both K-contiguous)
//   backward data    da = d W              A=(out,1)  B=(1,in)        (B is N-contiguous)
//   backward weight  G  = d^T a            A=(1,out)  B=(1,in)        (both MN-contiguous, K = batch)
// Two pairs accumulate into the same tile (a V^T + Ra W^T of the R-op; the two bilinear terms of the
// Hessian product).  This engine serves every shape the tcgen05 engine cannot take (feature widths
// that are not multiples of 4 floats = 16 B for TMA, tiny layers) and is its on-device FP32 check.
#pragma once
#include "common.cuh"

namespace hf {

enum Epilogue {
  EPI_STORE = 0,      // C = alpha*acc (+bias)                      (partials, plain products)
  EPI_BIAS_ACT = 1,   // C = act(acc + bias)                        (forward pass)
  EPI_BIAS_DACT = 2,  // C = alpha * (acc + bias) * act'(aux); C2 = acc+bias (R-op forward)
  EPI_DACT = 3,       // C = acc * act'(aux); C2 = acc              (backward data)
  EPI_DACT_H = 4      // C = acc * act'(aux) + ga * act''(aux) * rz (Hessian backward data)
};

// Split-precision image of an FP32 matrix: two BF16 planes with the matrix's own orientation and a row pitch of
// their own (a multiple of 8 elements = 16 B, what TMA needs): hi = bf16(x), lo = bf16(x - tf32_trunc(x)).
// Together with the FP32 words themselves (kind::tf32 reads their top 19 bits) these are the three operand forms of
// the split-precision product  A B ~= A_t B_t + A_lo B_hi + A_hi B_lo.  The pre-split tensor engine (gemm_tc2.cu)
// loads them ready-made; they are written once per linearisation for operands that are constant during a solve
// (inputs, activations, weights) and by the producing kernel's epilogue for the per-iteration ones.
struct Image16 {
  uint16_t* hi;     // plane 0; nullptr = no image
  int64_t plane;    // element offset from the hi plane to the lo plane
  int64_t ld;       // row pitch in elements (multiple of 8)
};

__device__ __forceinline__ uint16_t bf16_bits(float x) {
  uint16_t r;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
// one element of an image (element-wise kernels; the tile epilogues store four at a time, tc_common.cuh)
__device__ __forceinline__ void store_image1(const Image16& im, int64_t row, int col, float x) {
  uint16_t* hi = im.hi + row * im.ld + col;
  hi[0] = bf16_bits(x);
  hi[im.plane] = bf16_bits(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}

struct Operand {
  const float* ptr;
  int64_t s_mn;
  int64_t s_k;
  Image16 img;  // optional (zero-initialised by the brace initialisers used everywhere)
};

struct GemmArgs {
  int M, N, K, n_pairs;
  Operand A[2], B[2];
  int square;  // square every operand element on load (empirical-Fisher diagonal)
  float* C;
  int64_t ldc;
  float* C2;
  int epi, act;
  const float* bias;
  const float* aux;
  int64_t ldaux;
  const float* h_ga;
  const float* h_rz;
  float alpha;
  int split_k;      // gridDim.z; partial z goes to C + z * M * ldc
  int k_per_split;  // multiple of BK
  const int32_t* skip;
  float* colpart;  // tensor-core engine only: colpart[blockIdx.y * N + n] = column sums of the stored tile rows
  Image16 c_img;   // optional: the tensor engines also store the split-precision image of C (row pitch c_img.ld)
};

__device__ __forceinline__ float act_apply(int act, float z) {
  switch (act) {
    case HF_ACT_RELU: return z > 0.f ? z : 0.f;
    case HF_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case HF_ACT_TANH: return tanhf(z);
    default: return z;
  }
}
// derivative expressed through the stored post-activation value s = act(z)
__device__ __forceinline__ float act_d1(int act, float s) {
  switch (act) {
    case HF_ACT_RELU: return s > 0.f ? 1.f : 0.f;
    case HF_ACT_SIGMOID: return s * (1.f - s);
    case HF_ACT_TANH: return 1.f - s * s;
    default: return 1.f;
  }
}
__device__ __forceinline__ float act_d2(int act, float s) {
  switch (act) {
    case HF_ACT_SIGMOID: return s * (1.f - s) * (1.f - 2.f * s);
    case HF_ACT_TANH: return -2.f * s * (1.f - s * s);
    default: return 0.f;
  }
}

constexpr int kBK = 16;
constexpr int kGemmThreads = 256;

// Stage one [BMN x BK] operand tile: global -> registers (so the loads overlap the FMAs of the previous
// tile), then registers -> shared as s[k][mn] (k-major rows, so fragments are contiguous in mn).
template <int BMN>
struct TileStage {
  static constexpr int ELEMS = BMN * kBK / kGemmThreads;
  static constexpr int LD = BMN + 4;
  float reg[ELEMS > 0 ? ELEMS : 1];

  __device__ __forceinline__ void load(const Operand& op, int mn0, int k0, int MN, int k_end, bool vec, bool square) {
    const bool kcontig = op.s_k == 1;
    if (vec && ELEMS >= 4) {
#pragma unroll
      for (int i = 0; i < ELEMS / 4; ++i) {
        const int c = threadIdx.x + i * kGemmThreads;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kcontig) {
          const int row = c >> 2, kg = c & 3;
          if (mn0 + row < MN && k0 + kg * 4 < k_end)
            v = *reinterpret_cast<const float4*>(op.ptr + (int64_t)(mn0 + row) * op.s_mn + k0 + kg * 4);
        } else {
          const int k = c / (BMN / 4), mg = c % (BMN / 4);
          if (k0 + k < k_end && mn0 + mg * 4 < MN)
            v = *reinterpret_cast<const float4*>(op.ptr + (int64_t)(k0 + k) * op.s_k + mn0 + mg * 4);
        }
        if (square) v = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
        reg[4 * i] = v.x, reg[4 * i + 1] = v.y, reg[4 * i + 2] = v.z, reg[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < ELEMS; ++i) {
        const int c = threadIdx.x + i * kGemmThreads;
        int mn, k;
        if (kcontig) {
          mn = c / kBK, k = c % kBK;
        } else {
          k = c / BMN, mn = c % BMN;
        }
        float v = 0.f;
        if (mn0 + mn < MN && k0 + k < k_end) v = op.ptr[(int64_t)(mn0 + mn) * op.s_mn + (int64_t)(k0 + k) * op.s_k];
        reg[i] = square ? v * v : v;
      }
    }
  }

  __device__ __forceinline__ void store(float* s, const Operand& op, bool vec) const {
    const bool kcontig = op.s_k == 1;
    if (vec && ELEMS >= 4) {
#pragma unroll
      for (int i = 0; i < ELEMS / 4; ++i) {
        const int c = threadIdx.x + i * kGemmThreads;
        if (kcontig) {
          const int row = c >> 2, kg = c & 3;
#pragma unroll
          for (int e = 0; e < 4; ++e) s[(kg * 4 + e) * LD + row] = reg[4 * i + e];
        } else {
          const int k = c / (BMN / 4), mg = c % (BMN / 4);
          *reinterpret_cast<float4*>(s + k * LD + mg * 4) =
              make_float4(reg[4 * i], reg[4 * i + 1], reg[4 * i + 2], reg[4 * i + 3]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < ELEMS; ++i) {
        const int c = threadIdx.x + i * kGemmThreads;
        int mn, k;
        if (kcontig) {
          mn = c / kBK, k = c % kBK;
        } else {
          k = c / BMN, mn = c % BMN;
        }
        s[k * LD + mn] = reg[i];
      }
    }
  }
};

// fragment index -> tile coordinate: 8-wide fragments are split into two 4-groups half a tile apart so
// that shared loads are conflict-free and global stores of neighbouring threads are contiguous
template <int B, int T>
__device__ __forceinline__ int frag_coord(int t, int i) {
  if (T == 8) return (i < 4) ? t * 4 + i : B / 2 + t * 4 + (i - 4);
  return t * T + i;
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(kGemmThreads) gemm_simt_kernel(GemmArgs g, int vecA0, int vecB0, int vecA1, int vecB1) {
  static_assert((BM / TM) * (BN / TN) == kGemmThreads, "tile/thread mismatch");
  if (g.skip && *g.skip) return;
  __shared__ __align__(16) float sA[2][kBK * (BM + 4)];
  __shared__ __align__(16) float sB[2][kBK * (BN + 4)];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int ktiles = k_end > k_begin ? (k_end - k_begin + kBK - 1) / kBK : 0;
  const int n_it = ktiles * g.n_pairs;
  const bool sq = g.square != 0;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  TileStage<BM> ra;
  TileStage<BN> rb;
  auto fetch = [&](int it) {
    const int pr = it / ktiles, k0 = k_begin + (it % ktiles) * kBK;
    ra.load(g.A[pr], m0, k0, g.M, k_end, pr ? vecA1 : vecA0, sq);
    rb.load(g.B[pr], n0, k0, g.N, k_end, pr ? vecB1 : vecB0, sq);
  };
  auto commit = [&](int it) {
    const int pr = it / ktiles;
    ra.store(sA[it & 1], g.A[pr], pr ? vecA1 : vecA0);
    rb.store(sB[it & 1], g.B[pr], pr ? vecB1 : vecB0);
  };
  if (n_it > 0) {
    fetch(0);
    commit(0);
  }
  __syncthreads();
  for (int it = 0; it < n_it; ++it) {
    if (it + 1 < n_it) fetch(it + 1);
    const float* a_s = sA[it & 1];
    const float* b_s = sB[it & 1];
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) af[i] = a_s[k * (BM + 4) + frag_coord<BM, TM>(ty, i)];
#pragma unroll
      for (int j = 0; j < TN; ++j) bf[j] = b_s[k * (BN + 4) + frag_coord<BN, TN>(tx, j)];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    if (it + 1 < n_it) commit(it + 1);
    __syncthreads();
  }

  float* C = g.C + (g.split_k > 1 ? (int64_t)blockIdx.z * g.M * g.ldc : 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + frag_coord<BM, TM>(ty, i);
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + frag_coord<BN, TN>(tx, j);
      if (n >= g.N) continue;
      float v = acc[i][j];
      switch (g.epi) {
        case EPI_STORE:
          v = g.alpha * v + (g.bias ? g.bias[n] : 0.f);
          break;
        case EPI_BIAS_ACT:
          v = act_apply(g.act, v + (g.bias ? g.bias[n] : 0.f));
          break;
        case EPI_BIAS_DACT: {
          v += g.bias ? g.bias[n] : 0.f;
          if (g.C2) g.C2[(int64_t)m * g.ldc + n] = v;
          if (g.act != HF_ACT_NONE) v *= act_d1(g.act, g.aux[(int64_t)m * g.ldaux + n]);
          v *= g.alpha;
        } break;
        case EPI_DACT: {
          if (g.C2) g.C2[(int64_t)m * g.ldc + n] = v;
          if (g.act != HF_ACT_NONE) v *= act_d1(g.act, g.aux[(int64_t)m * g.ldaux + n]);
        } break;
        case EPI_DACT_H: {
          const float s = g.act != HF_ACT_NONE ? g.aux[(int64_t)m * g.ldaux + n] : 0.f;
          v = v * act_d1(g.act, s);
          if (g.h_ga) v += g.h_ga[(int64_t)m * g.ldaux + n] * act_d2(g.act, s) * g.h_rz[(int64_t)m * g.ldaux + n];
        } break;
      }
      C[(int64_t)m * g.ldc + n] = v;
    }
  }
}

// out[i] = (accumulate ? out[i] : 0) + scale * sum_s part[s*stride + i]    (fixed order: deterministic)
static __global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, int64_t count, int64_t stride,
                                       float* __restrict__ out, float scale, int accumulate,
                                       const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * stride + i];
    out[i] = (accumulate ? out[i] : 0.f) + scale * s;
  }
}

// Two reductions in one launch: a weight-gradient slice and the bias slice that follows it in the flat layout.
// perm_taps > 0 (convolution layers, conv.cuh): the partial tiles hold the weight slice with columns (tap, c_in); it is
// written back with columns (c_in, tap), the order of PyTorch's flat parameter vector.
static __global__ void reduce_partials2_kernel(const float* __restrict__ partW, int splitsW, int64_t countW,
                                               float* __restrict__ outW, const float* __restrict__ partB, int splitsB,
                                               int64_t countB, float* __restrict__ outB, float scale, int accumulate,
                                               const int32_t* __restrict__ skip, int perm_cin = 0, int perm_taps = 0) {
  if (skip && *skip) return;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  // weight slice, 4 elements per thread: all split loads of a group are independent, so they are in flight together
  const bool vec = perm_taps == 0 && countW % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(partW) | reinterpret_cast<uintptr_t>(outW)) & 15u) == 0;
  const int64_t nvec = vec ? countW / 4 : 0;
  for (int64_t v = tid; v < nvec; v += nthreads) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int z = 0; z < splitsW; ++z) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(partW + (int64_t)z * countW) + v);
      s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
    }
    float4* o = reinterpret_cast<float4*>(outW) + v;
    float4 r = accumulate ? *o : make_float4(0.f, 0.f, 0.f, 0.f);
    r.x += scale * s.x, r.y += scale * s.y, r.z += scale * s.z, r.w += scale * s.w;
    *o = r;
  }
  // scalar tail of the weight slice (unaligned case) and the bias slice
  const int64_t scalar_w = countW - 4 * nvec, total = scalar_w + countB;
  for (int64_t i = tid; i < total; i += nthreads) {
    const bool w = i < scalar_w;
    const int64_t j = w ? 4 * nvec + i : i - scalar_w;
    const float* part = w ? partW : partB;
    const int64_t stride = w ? countW : countB;
    const int splits = w ? splitsW : splitsB;
    float s = 0.f;
#pragma unroll 8
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * stride + j];
    float* out = w ? outW : outB;
    int64_t dst = j;
    if (w && perm_taps > 0) {
      const int64_t K = (int64_t)perm_cin * perm_taps, o = j / K, col = j % K;  // col = tap * c_in + c
      dst = o * K + (col % perm_cin) * perm_taps + col / perm_cin;
    }
    out[dst] = (accumulate ? out[dst] : 0.f) + scale * s;
  }
}

// The same reduction for MANY partial sets of a SHORT vector (the fused head leaves one set per CTA): 8 groups of
// splits per element run in parallel (z = g, g + 8, ...) and are combined in fixed order, so the chain of dependent
// loads is splits / 8 long instead of splits.
static __global__ void __launch_bounds__(256) reduce_tall_kernel(const float* __restrict__ partW, int splitsW, int64_t countW,
                                                                 float* __restrict__ outW, const float* __restrict__ partB,
                                                                 int splitsB, int64_t countB, float* __restrict__ outB,
                                                                 float scale, int accumulate, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ float sm[8][33];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t total = countW + countB;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + e;
    const bool w = i < countW;
    const int64_t j = w ? i : i - countW;
    float s = 0.f;
    if (i < total) {
      const float* part = w ? partW : partB;
      const int64_t stride = w ? countW : countB;
      const int splits = w ? splitsW : splitsB;
#pragma unroll 4
      for (int z = g; z < splits; z += 8) s += part[(int64_t)z * stride + j];
    }
    sm[g][e] = s;
    __syncthreads();
    if (g == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][e];
      float* out = w ? outW : outB;
      out[j] = (accumulate ? out[j] : 0.f) + scale * t;
    }
    __syncthreads();
  }
}

// part[z][c] = sum over the z-th row range of d[n][c] (optionally squared): bias gradients
static __global__ void colsum_kernel(const float* __restrict__ d, int64_t rows, int cols, int64_t ld, int rows_per_split,
                              int square, float* __restrict__ part, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = min(rows, r0 + rows_per_split);
  float s = 0.f;
  if (c < cols)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float v = d[r * ld + c];
      s += square ? v * v : v;
    }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    part[(int64_t)blockIdx.y * cols + c] = t;
  }
}

struct TileChoice {
  int bm, bn;
};

inline TileChoice choose_tile(int M, int N) {
  if (N <= 16 && M > 16) return {128, 16};
  if (M <= 16 && N > 16) return {16, 128};
  if (M <= 64 || N <= 64) return {64, 64};
  // The large tile runs one 8-warp CTA per SM (147 registers): it only pays once there are several waves of it.
  // Measured at M=4096 N=512 K=784: 128x128 tiles 150 us, 64x64 tiles see profiles/.
  const int64_t big = (int64_t)((M + 127) / 128) * ((N + 127) / 128);
  return big >= 600 ? TileChoice{128, 128} : TileChoice{64, 64};
}

inline bool vec_ok(const Operand& op, int MN, int K) {
  if (!op.ptr) return false;
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15u) != 0) return false;
  if (op.s_k == 1) return op.s_mn % 4 == 0 && K % 4 == 0;
  return op.s_k % 4 == 0 && MN % 4 == 0;
}

// Launch the contraction described by g (g.split_k and g.k_per_split already set).
inline int launch_gemm_simt(GemmArgs g, cudaStream_t stream) {
  HF_REQUIRE(g.M > 0 && g.N > 0 && g.K >= 0 && g.n_pairs >= 1 && g.n_pairs <= 2, HF_ERR_INVALID, "gemm: bad shape");
  for (int s = 0; s < g.n_pairs; ++s)
    HF_REQUIRE((g.A[s].s_k == 1 || g.A[s].s_mn == 1) && (g.B[s].s_k == 1 || g.B[s].s_mn == 1), HF_ERR_INVALID,
               "gemm: each operand needs one unit stride");
  if (g.split_k < 1) g.split_k = 1;
  if (g.split_k == 1) g.k_per_split = ((g.K + kBK - 1) / kBK) * kBK;
  if (g.k_per_split < kBK) g.k_per_split = kBK;
  const TileChoice t = choose_tile(g.M, g.N);
  const dim3 grid((g.N + t.bn - 1) / t.bn, (g.M + t.bm - 1) / t.bm, g.split_k);
  const int vA0 = vec_ok(g.A[0], g.M, g.K), vB0 = vec_ok(g.B[0], g.N, g.K);
  const int vA1 = g.n_pairs > 1 ? vec_ok(g.A[1], g.M, g.K) : 0, vB1 = g.n_pairs > 1 ? vec_ok(g.B[1], g.N, g.K) : 0;
  if (t.bm == 128 && t.bn == 128)
    gemm_simt_kernel<128, 128, 8, 8><<<grid, kGemmThreads, 0, stream>>>(g, vA0, vB0, vA1, vB1);
  else if (t.bm == 64)
    gemm_simt_kernel<64, 64, 4, 4><<<grid, kGemmThreads, 0, stream>>>(g, vA0, vB0, vA1, vB1);
  else if (t.bm == 128)
    gemm_simt_kernel<128, 16, 8, 1><<<grid, kGemmThreads, 0, stream>>>(g, vA0, vB0, vA1, vB1);
  else
    gemm_simt_kernel<16, 128, 1, 8><<<grid, kGemmThreads, 0, stream>>>(g, vA0, vB0, vA1, vB1);
  HF_LAUNCH_CHECK();
  return HF_OK;
}

}  // namespace hf
