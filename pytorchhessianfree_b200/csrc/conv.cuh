// Convolution and pooling as layers of the layer program (first conv slice: SURVEY.md section 8, row N1).
//
// A Conv2d layer is the linear layer  z = U W^T + b  on the UNFOLDED input U = im2col(a_{l-1}):
//   rows of U   = output positions (n, oy, ox)            -> R_l = N * H_out * W_out
//   columns of U = (c_in, ky, kx), c_in slowest           -> K_l = C_in * k_h * k_w
// which is exactly how PyTorch flattens weight[C_out, C_in, k_h, k_w]: the layer's slice of the flat parameter
// vector IS the [C_out, K_l] operand the tile engines read, the direction slice IS V_l, and the weight-gradient
// contraction  cot^T U  lands in the flat layout without any reordering.  Activations are kept position-major
// ("NHWC"): [R_l, C_out] with the library's 16-byte row pitch, i.e. the same [rows, width] matrices the fully connected
// path uses, only with more rows than samples.  So the R-op, the transposed sweep and the weight gradients of a conv
// layer run on the tensor-core tile kernels unchanged; what this file adds is the data movement around them:
//   im2col   a_{l-1} (or its tangent)  -> U       (once per linearisation for a, once per product for the tangent)
//   fold     dU = cot W  -> cot_{l-1} = act'(a_{l-1}) * col2im(dU)   (gather form: deterministic, no atomics)
//   pool / unpool  for a global average pool (All-CNN-C: 6x6 -> 1x1 in front of the loss)
// 1x1 / stride 1 / unpadded convolutions have U = a_{l-1}: no copy, no fold, the fused epilogues apply directly.
// Every kernel can also write the split-precision image (gemm_simt.cuh: Image16) of what it produces, so that the
// pair engine reads conv operands like any other.
#pragma once
#include "gemm_simt.cuh"

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

// U[(n, oy, ox), (c, ky, kx)] = src[(n, oy*s - p + ky, ox*s - p + kx), c]  (0 outside the map)
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                                                     int64_t n_samples, ConvGeom g, Image16 img, const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  const int64_t rows = n_samples * g.hout * g.wout;
  const int64_t total = rows * g.cin;
  const int K = g.cin * g.kh * g.kw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.cin);
    const int64_t r = i / g.cin;
    const int ox = (int)(r % g.wout), oy = (int)((r / g.wout) % g.hout);
    const int64_t s = r / ((int64_t)g.wout * g.hout);
    float* d = dst + r * ld_dst + c * g.kh * g.kw;
    for (int ky = 0; ky < g.kh; ++ky) {
      const int iy = oy * g.stride - g.pad + ky;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int ix = ox * g.stride - g.pad + kx;
        float v = 0.f;
        if (iy >= 0 && iy < g.hin && ix >= 0 && ix < g.win) v = src[((s * g.hin + iy) * g.win + ix) * ld_src + c];
        d[ky * g.kw + kx] = v;
        if (img.hi) store_image1(img, r, c * g.kh * g.kw + ky * g.kw + kx, v);
      }
    }
    if (c == g.cin - 1)
      for (int k = K; k < ld_dst; ++k) dst[r * ld_dst + k] = 0.f;  // the 16-byte row pitch's padding
  }
}

__device__ __forceinline__ float conv_act_d1(int act, float s) { return act_d1(act, s); }

// cot_prev[(n, y, x), c] = act'(a_prev) * sum over the taps (ky, kx) that read input position (y, x):
//   dU[(n, oy, ox), (c, ky, kx)]  with  oy*s - p + ky = y,  ox*s - p + kx = x
__global__ void __launch_bounds__(256) fold_kernel(const float* __restrict__ dU, int ld_du, const float* __restrict__ a_prev, int ld_a, int act_prev,
                                                   float* __restrict__ dst, int64_t n_samples, ConvGeom g, Image16 img,
                                                   const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  const int64_t rows_in = n_samples * g.hin * g.win;
  const int64_t total = rows_in * g.cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.cin);
    const int64_t r = i / g.cin;
    const int x = (int)(r % g.win), y = (int)((r / g.win) % g.hin);
    const int64_t s = r / ((int64_t)g.win * g.hin);
    float acc = 0.f;
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
        acc += dU[((s * g.hout + oy) * g.wout + ox) * ld_du + (c * g.kh + ky) * g.kw + kx];
      }
    }
    const float v = act_prev == HF_ACT_NONE ? acc : acc * act_d1(act_prev, a_prev[r * ld_a + c]);
    dst[r * ld_a + c] = v;
    if (img.hi) store_image1(img, r, c, v);
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
                                                     const int32_t* __restrict__ skip) {
  if (skip && *skip) return;
  const int64_t total = n_samples * hw * (int64_t)c;
  const float inv = 1.f / (float)hw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    const int64_t r = i / c;
    const int64_t s = r / hw;
    float v = cur[s * ld + ch] * inv;
    if (act_prev != HF_ACT_NONE) v *= act_d1(act_prev, a_prev[r * ld + ch]);
    dst[r * ld + ch] = v;
    if (img.hi) store_image1(img, r, ch, v);
  }
}

inline unsigned conv_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = 16 * (int64_t)sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace hf
