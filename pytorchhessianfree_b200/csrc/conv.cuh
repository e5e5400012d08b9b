// Convolution and pooling as layers of the layer program (first conv slice: SURVEY.md section 8, row N1).
//
// A Conv2d layer is the linear layer  z = U W^T + b  on the UNFOLDED input U = im2col(a_{l-1}):
//   rows of U   = output positions (n, oy, ox)            -> R_l = N * H_out * W_out
//   columns of U = (ky, kx, c_in), the tap slowest        -> K_l = k_h * k_w * C_in
// Activations are kept position-major ("NHWC"): [R_l, C_out] with the library's 16-byte row pitch, i.e. the same
// [rows, width] matrices the fully connected path uses, only with more rows than samples.  With the tap slowest, one
// tap of one output position is a CONTIGUOUS run of C_in floats of the input row it reads: im2col and its transpose
// move whole 16-byte groups, coalesced on both sides.  PyTorch flattens weight[C_out, C_in, k_h, k_w] with the channel
// slowest, so the layer's weight and direction slices are re-ordered into [C_out, (tap, c_in)] operand copies (by the
// split pass that makes the operand forms anyway: gemm_tc.cuh SplitSegment.perm_*), and the weight-gradient reduction
// writes its result back in PyTorch's order: the flat-vector contract of the ABI is untouched.  So the R-op, the
// transposed sweep and the weight gradients of a conv layer run on the tensor-core tile kernels unchanged; what this
// file adds is the data movement around them:
//   im2col   a_{l-1} (or its tangent)  -> U       (once per linearisation for a, once per product for the tangent)
//   fold     dU = cot W  -> cot_{l-1} = act'(a_{l-1}) * col2im(dU)   (gather form: deterministic, no atomics)
//   pool / unpool  for a global average pool (All-CNN-C: 6x6 -> 1x1 in front of the loss)
// 1x1 / stride 1 / unpadded convolutions have U = a_{l-1}: no copy, no fold, the fused epilogues apply directly.
// Every kernel can also write the split-precision image (gemm_simt.cuh: Image16) of what it produces, so that the
// pair engine reads conv operands like any other.
#pragma once
#include "tc_common.cuh"

namespace hf {

struct ConvGeom {
  int cin, hin, win, kh, kw, stride, pad, hout, wout;
};

// x[N, C, H, W] (PyTorch layout) -> dst[(n, y, x), c] with row pitch ld
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int c, int hw,
                                                           int ld, Image16 img) {
  const int64_t total = n * hw * (int64_t)c;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int64_t r = i / c;  // (n, y, x)
    const int64_t s = r / hw;
    const int p = (int)(r % hw);
    const float v = src[(s * c + ch) * hw + p];
    dst[r * ld + ch] = v;
    if (img.hi) store_image1(img, r, ch, v);
  }
}

// U[(n, oy, ox), (ky, kx, c)] = src[(n, oy*s - p + ky, ox*s - p + kx), c]  (0 outside the map); one thread per
// (row, group of 4 columns of the pitched row)
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                                                     int64_t n_samples, ConvGeom g, Image16 img, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  const int taps = g.kh * g.kw, K = taps * g.cin;
  const int groups = ld_dst >> 2;  // 4-column groups of the PITCHED row: the padding columns get zeros
  const int64_t rows = n_samples * g.hout * g.wout;
  const int64_t total = rows * groups;
  const bool vec = g.cin % 4 == 0;  // a group never straddles two taps, and the source run is 16-byte aligned
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % groups) * 4;
    const int64_t r = i / groups;
    const int ox = (int)(r % g.wout), oy = (int)((r / g.wout) % g.hout);
    const int64_t s = r / ((int64_t)g.wout * g.hout);
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {
      if (k0 < K) {
        const int tap = k0 / g.cin, c = k0 % g.cin;
        const int iy = oy * g.stride - g.pad + tap / g.kw, ix = ox * g.stride - g.pad + tap % g.kw;
        if (iy >= 0 && iy < g.hin && ix >= 0 && ix < g.win) {
          const float4 v = *reinterpret_cast<const float4*>(src + ((s * g.hin + iy) * g.win + ix) * ld_src + c);
          x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e;
        if (k >= K) continue;
        const int tap = k / g.cin, c = k % g.cin;
        const int iy = oy * g.stride - g.pad + tap / g.kw, ix = ox * g.stride - g.pad + tap % g.kw;
        if (iy >= 0 && iy < g.hin && ix >= 0 && ix < g.win) x[e] = src[((s * g.hin + iy) * g.win + ix) * ld_src + c];
      }
    }
    *reinterpret_cast<float4*>(dst + r * ld_dst + k0) = make_float4(x[0], x[1], x[2], x[3]);
    if (img.hi) st4_image(img, r, k0, 4, x);  // image pitch pad8(K) >= pad4(K): in range
  }
}

// cot_prev[(n, y, x), c] = act'(a_prev) * sum over the taps (ky, kx) that read input position (y, x):
//   dU[(n, oy, ox), (ky, kx, c)]  with  oy*s - p + ky = y,  ox*s - p + kx = x;   one thread per (input row, 4 channels)
// Optional, for Hessian products (Pearlmutter): ga_out receives the folded value BEFORE the activation derivative
// (dloss/da of the layer below, kept by the gradient pass); h_ga / h_rz add the second-order activation term
// ga * act''(a) * R{z} of the layer below (the conv analogue of EPI_DACT_H).
__global__ void __launch_bounds__(256) fold_kernel(const float* __restrict__ dU, int ld_du, const float* __restrict__ a_prev, int ld_a, int act_prev,
                                                   float* __restrict__ dst, int64_t n_samples, ConvGeom g, Image16 img,
                                                   const int32_t* __restrict__ skip, float* __restrict__ ga_out = nullptr,
                                                   const float* __restrict__ h_ga = nullptr, const float* __restrict__ h_rz = nullptr) {
  if (skip && *skip) return;
  const int groups = (g.cin + 3) >> 2;
  const int64_t rows_in = n_samples * g.hin * g.win;
  const int64_t total = rows_in * groups;
  const bool vec = g.cin % 4 == 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) * 4, cnt = min(4, g.cin - c0);
    const int64_t r = i / groups;
    const int x = (int)(r % g.win), y = (int)((r / g.win) % g.hin);
    const int64_t s = r / ((int64_t)g.win * g.hin);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < g.kh; ++ky) {
      const int ty = y + g.pad - ky;
      if (ty < 0 || ty % g.stride) continue;
      const int oy = ty / g.stride;
      if (oy >= g.hout) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int tx = x + g.pad - kx;
        if (tx < 0 || tx % g.stride) continue;
        const int ox = tx / g.stride;
        if (ox >= g.wout) continue;
        const float* p = dU + ((s * g.hout + oy) * g.wout + ox) * ld_du + (ky * g.kw + kx) * g.cin + c0;
        if (vec) {
          const float4 v = *reinterpret_cast<const float4*>(p);
          acc[0] += v.x, acc[1] += v.y, acc[2] += v.z, acc[3] += v.w;
        } else {
          for (int e = 0; e < cnt; ++e) acc[e] += p[e];
        }
      }
    }
    if (ga_out)
      for (int e = 0; e < cnt; ++e) ga_out[r * ld_a + c0 + e] = acc[e];
    if (act_prev != HF_ACT_NONE)
      for (int e = 0; e < cnt; ++e) {
        const float sa = a_prev[r * ld_a + c0 + e];
        acc[e] *= act_d1(act_prev, sa);
        if (h_ga) acc[e] += h_ga[r * ld_a + c0 + e] * act_d2(act_prev, sa) * h_rz[r * ld_a + c0 + e];
      }
    if (vec) {
      *reinterpret_cast<float4*>(dst + r * ld_a + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
      for (int e = 0; e < cnt; ++e) dst[r * ld_a + c0 + e] = acc[e];
    }
    if (img.hi) st4_image(img, r, c0, cnt, acc);
  }
}

// dst[n, c] = mean over the hw positions of src[(n, p), c]
__global__ void __launch_bounds__(256) avgpool_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, int64_t n_samples, int hw, int c,
                                                      Image16 img, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  const int64_t total = n_samples * c;
  const float inv = 1.f / (float)hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int64_t s = i / c;
    float acc = 0.f;
    for (int p = 0; p < hw; ++p) acc += src[(s * hw + p) * ld + ch];
    const float v = acc * inv;
    dst[s * ld + ch] = v;
    if (img.hi) store_image1(img, s, ch, v);
  }
}

// dst[(n, p), c] = cur[n, c] / hw * act'(a_prev[(n, p), c])   (transposed average pool, through the previous activation)
__global__ void __launch_bounds__(256) unpool_kernel(const float* __restrict__ cur, int ld, const float* __restrict__ a_prev, int act_prev,
                                                     float* __restrict__ dst, int64_t n_samples, int hw, int c, Image16 img,
                                                     const int32_t* __restrict__ skip, float* __restrict__ ga_out = nullptr,
                                                     const float* __restrict__ h_ga = nullptr, const float* __restrict__ h_rz = nullptr) {
  if (skip && *skip) return;
  const int64_t total = n_samples * hw * (int64_t)c;
  const float inv = 1.f / (float)hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int64_t r = i / c;
    const int64_t s = r / hw;
    float v = cur[s * ld + ch] * inv;
    if (ga_out) ga_out[r * ld + ch] = v;
    if (act_prev != HF_ACT_NONE) {
      const float sa = a_prev[r * ld + ch];
      v *= act_d1(act_prev, sa);
      if (h_ga) v += h_ga[r * ld + ch] * act_d2(act_prev, sa) * h_rz[r * ld + ch];
    }
    dst[r * ld + ch] = v;
    if (img.hi) store_image1(img, r, ch, v);
  }
}

// ---- empirical-Fisher diagonal of a convolution -----------------------------------------------------------
// For a layer applied at S positions per sample the per-sample gradient is a SUM over positions, and the square sits
// outside it:   F[co, k] = sum_n ( sum_pos cot[(n, pos), co] * U[(n, pos), k] )^2,   F_b[co] = sum_n ( sum_pos cot )^2
// -- not the (d^2)^T (a^2) contraction of the fully connected layers.  One CTA owns a 64 x 64 tile of (co, k) and a
// group of samples: per sample a [64 x S] x [S x 64] product in registers (4 x 4 per thread, 16 positions per shared
// stage), squared and added to the CTA's running tile; the groups' partial tiles go to part[group][co][k] and
// reduce_partials2_kernel sums them in a fixed order (deterministic; it also re-orders tap-major k into PyTorch's
// channel-major weight layout).  FP32 on the CUDA cores: once per optimizer step, ~2 curvature products' worth.
__global__ void __launch_bounds__(256) conv_fisher_w_kernel(const float* __restrict__ cot, int ld_c, const float* __restrict__ U, int ld_u,
                                                            int64_t n_samples, int S, int M, int K, int samples_per_group,
                                                            float* __restrict__ part, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ __align__(16) float sc[16][64];
  __shared__ __align__(16) float su[16][64];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int co0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int64_t s_begin = (int64_t)blockIdx.z * samples_per_group;
  const int64_t s_end = min(n_samples, s_begin + samples_per_group);
  const int lr = t >> 4, lc = (t & 15) * 4;  // this thread's slot of a 16 x 64 stage
  float sq[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sq[i][j] = 0.f;
  for (int64_t smp = s_begin; smp < s_end; ++smp) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int p0 = 0; p0 < S; p0 += 16) {
      const int pos = p0 + lr;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (pos < S) {
        const int64_t row = smp * S + pos;
        const float* pc = cot + row * ld_c + co0 + lc;
        const float* pu = U + row * ld_u + k0 + lc;
        if (co0 + lc + 4 <= M) {
          a = *reinterpret_cast<const float4*>(pc);
        } else {
          if (co0 + lc + 0 < M) a.x = pc[0];
          if (co0 + lc + 1 < M) a.y = pc[1];
          if (co0 + lc + 2 < M) a.z = pc[2];
        }
        if (k0 + lc + 4 <= K) {
          b = *reinterpret_cast<const float4*>(pu);
        } else {
          if (k0 + lc + 0 < K) b.x = pu[0];
          if (k0 + lc + 1 < K) b.y = pu[1];
          if (k0 + lc + 2 < K) b.z = pu[2];
        }
      }
      *reinterpret_cast<float4*>(&sc[lr][lc]) = a;
      *reinterpret_cast<float4*>(&su[lr][lc]) = b;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const float4 x = *reinterpret_cast<const float4*>(&sc[kk][ty * 4]);
        const float4 y = *reinterpret_cast<const float4*>(&su[kk][tx * 4]);
        const float xv[4] = {x.x, x.y, x.z, x.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sq[i][j] = fmaf(acc[i][j], acc[i][j], sq[i][j]);
  }
  float* dst = part + (int64_t)blockIdx.z * M * K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) dst[(int64_t)co * K + k] = sq[i][j];
    }
  }
}

// part[group][co] = sum over the group's samples of ( sum over the sample's positions of cot[(n, pos), co] )^2
__global__ void __launch_bounds__(256) conv_fisher_b_kernel(const float* __restrict__ cot, int ld_c, int64_t n_samples, int S, int M,
                                                            int samples_per_group, float* __restrict__ part, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t s_begin = (int64_t)blockIdx.y * samples_per_group;
  const int64_t s_end = min(n_samples, s_begin + samples_per_group);
  float total = 0.f;
  for (int64_t smp = s_begin; smp < s_end; ++smp) {
    float s = 0.f;
    if (c < M)
      for (int pos = threadIdx.y; pos < S; pos += 8) s += cot[(smp * S + pos) * ld_c + c];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
      float tsum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) tsum += sm[i][threadIdx.x];
      total = fmaf(tsum, tsum, total);
    }
    __syncthreads();
  }
  if (threadIdx.y == 0 && c < M) part[(int64_t)blockIdx.y * M + c] = total;
}

inline unsigned conv_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = 16 * (int64_t)sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace hf
