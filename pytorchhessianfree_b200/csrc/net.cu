// Layer program, linearisation and curvature-matrix-vector products for fully connected nets.
//
// Replaces, behind the C ABI of include/hf_b200.h, what the reference obtains from BackPACK/autograd:
//   hf_lin_forward / hf_lin_gradient  <- forward() + autograd.grad, optimizer.py:223, :231-234
//   hf_ggn_matvec                     <- _Gv,  optimizer.py:457-462  (R-op, loss Hessian, L-op)
//   hf_hessian_matvec                 <- _Hv,  optimizer.py:450-455  (Pearlmutter R{backprop})
//   hf_fisher_diag                    <- diag_EF_backpack, preconditioners.py:11-60
// Activations are computed once per linearisation and stay resident in HBM for the whole solve (the
// reference re-runs the forward pass for every chunk on every CG iteration, optimizer.py:805-814).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "conv.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "head.cuh"

namespace hf {

struct Layer {
  int in, out, act, has_bias;
  int64_t w_off, b_off;
  const float* w_frozen;
  const float* b_frozen;
  int kind;       // hf_layer_kind.  CONV2D: in = c_in*k_h*k_w, the contraction runs on the unfolded input (conv.cuh)
  ConvGeom geom;  // CONV2D / AVGPOOL
  int s_in, s_out;  // rows (positions) per sample of the layer's input / output activation: 1 for fully connected layers
  bool unfold;    // CONV2D whose unfolded input is not the input itself (anything but 1x1, stride 1, no padding)
};

}  // namespace hf

struct hf_net {
  std::vector<hf::Layer> L;
  int loss, reduction;
  int64_t P;
  int first_trainable;
  int max_width;
  int classes;
  int engine;  // 0 = SIMT everywhere, 1 = tcgen05 where the shape allows
  bool has_relu;
  bool has_conv;   // any CONV2D / AVGPOOL layer: activations have more rows than samples
  int max_s;       // largest rows-per-sample of any activation
  int64_t max_act; // largest floats-per-sample of any [rows, width] matrix a product touches (activations, unfolded inputs)
};

struct hf_lin {
  const hf_net* net;
  int64_t N;
  int flags;
  const float* x;
  std::vector<float*> a;      // a[l]: output of layer l (post-activation); a[L-1] = network output
  std::vector<float*> delta;  // HESSIAN: dloss/dz_l
  std::vector<float*> ga;     // HESSIAN: dloss/da_l for layers whose activation has curvature
  std::vector<float*> ra;     // HESSIAN: R{a_l}
  std::vector<float*> rz;     // HESSIAN: R{z_l} for layers whose activation has curvature
  float* prob;                // [N,C] softmax / sigmoid probabilities
  float* deltaL;              // [N,C] dloss/dz of the last layer
  float* buf[2];              // ping-pong [N,max_width]
  float* partial_fwd;         // split-K partial tiles of the last R-op layer when the output is narrow
  int fwd_splits_max;
  float* partial;             // split-K partial tiles
  size_t partial_floats;
  bool head_ok;               // the fused output head (head.cuh) applies to GGN products of this linearisation
  hf::HeadPlan head;
  float* head_partW;          // [head.ctas][C*D]
  float* head_partB;          // [head.ctas][C]
  float* partial_main;        // second set for the first trainable layer, whose gradient runs on the caller's stream
  size_t partial_main_floats;
  bool loss_fused;            // the last rop_forward already applied the (diagonal) loss Hessian in its epilogue
  int fwd_splits;             // > 0: the last rop_forward left split partials in partial_fwd
  const float* fwd_bias;
  std::vector<float*> cot;    // cot[l]: cotangent dloss/dz_l (or its R-derivative) of the sweep in flight
  std::vector<float*> colbuf; // colbuf[l]: column-sum partials of cot[l] (bias gradient): [row blocks][out_l]
  size_t col_rows;            // row blocks each colbuf[l] can hold
  // Layers whose fan-in is not a multiple of 4 floats cannot be addressed by TMA in place (16-byte row pitch): the
  // tensor engine reads 16-byte-pitched copies instead.  wpad[l] is refreshed by hf_lin_forward (weights are fixed
  // for the life of a linearisation), vpad[l] at the start of every product.
  std::vector<float*> wpad, vpad;
  // Split-precision images (gemm_simt.cuh: Image16) for the pre-split pair engine: one per library-owned [rows, width]
  // matrix that a contraction may read, looked up by the FP32 base pointer.  Constant operands (inputs, activations,
  // weights) are split once in hf_lin_forward, the CG direction once per product, tangents and cotangents by the
  // epilogue of the kernel that produces them.
  struct ImgBuf {
    const float* base;
    uint16_t* hi;
    int64_t plane;  // elements between the hi and the lo plane
  };
  bool use_images;
  std::vector<ImgBuf> imgs;
  ImgBuf x_img;
  std::vector<ImgBuf> w_img, v_img;  // per layer; base filled in when the parameter / direction pointer is known
  // Empirical-Fisher diagonal on the tensor engines: (d^2)^T (a^2) is an ordinary contraction of the SQUARED operands,
  // which one split pass writes here (FP32 with a 16-byte pitch, + images when the linearisation keeps them).
  float* sq[2];
  ImgBuf sq_img[2];
  // conv layers (conv.cuh): the input in position-major layout when the first layer is a convolution, the unfolded
  // input U_l of every layer that needs one (constant for the life of the linearisation), and two scratch matrices of
  // the largest unfolded size: the unfolded tangent of the R-op and dU = cot W of the transposed sweep
  float* xn;
  std::vector<float*> U;
  std::vector<float*> RU;  // HESSIAN: the unfolded R{a_{l-1}} of every unfolded layer, kept for the transposed sweep
  float* ru;
  float* du;
  const float* pending_cur;   // phased sweep: cotangent of the first trainable layer, left by phase 0 for phase 1
  int pending_cols;
  cudaStream_t side;          // weight/bias gradients of layer l run here, concurrently with the data product that
  std::vector<cudaEvent_t> ev;  // continues the sweep on the caller's stream (ev[l]: cot[l] is ready)
  cudaEvent_t join;
  double* loss_partial;
  int loss_blocks;
  int64_t n_total;
  bool have_forward, have_gradient;
};

namespace hf {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline bool curved(int act) { return act == HF_ACT_SIGMOID || act == HF_ACT_TANH; }
// Every [batch, width] buffer the library owns is stored with its leading dimension rounded up to 4 floats
// (16 B): TMA needs 16-byte row pitches, and this is what lets the 10-class output layer of the MLP config run on
// the tensor-core tiles.  The padding columns are never read (all kernels bound their accesses by the width).
static inline int pad4(int w) { return (w + 3) & ~3; }
static inline int pad8(int w) { return (w + 7) & ~7; }  // row pitch of the BF16 image planes (16 B)

// the image of the library-owned matrix at `base`, read as rows of `width` elements (nullptr image if there is none)
static Image16 image_for(const hf_lin* lin, const float* base, int width) {
  Image16 im = {nullptr, 0, 0};
  if (!lin->use_images || !base) return im;
  auto hit = [&](const hf_lin::ImgBuf& b) {
    if (b.base != base || !b.hi) return false;
    im.hi = b.hi, im.plane = b.plane, im.ld = pad8(width);
    return true;
  };
  if (hit(lin->x_img)) return im;
  for (const auto& b : lin->imgs)
    if (hit(b)) return im;
  for (const auto& b : lin->w_img)
    if (hit(b)) return im;
  for (const auto& b : lin->v_img)
    if (hit(b)) return im;
  return im;
}
static Operand with_image(const hf_lin* lin, Operand op, int width) {
  op.img = image_for(lin, op.ptr, width);
  return op;
}

// rows of the activation layer l writes / reads: samples x positions per sample (1 for fully connected layers)
static inline int64_t rows_out(const hf_lin* lin, int l) { return lin->N * lin->net->L[l].s_out; }
static inline int64_t rows_in(const hf_lin* lin, int l) { return lin->N * lin->net->L[l].s_in; }

// The [rows_out(l), L.in] matrix layer l contracts with its weight: the unfolded input of a convolution, the input
// itself otherwise (position-major copy of x when a convolution comes first).  *ld receives its row pitch.
static const float* layer_input(const hf_lin* lin, int l, int* ld) {
  const hf::Layer& L = lin->net->L[l];
  if (L.unfold) {
    *ld = pad4(L.in);
    return lin->U[l];
  }
  if (l == 0) {
    *ld = L.kind == HF_LAYER_LINEAR ? L.in : pad4(L.in);
    return L.kind == HF_LAYER_LINEAR ? lin->x : lin->xn;
  }
  *ld = pad4(L.in);
  return lin->a[l - 1];
}

// ---- row-wise loss kernels ---------------------------------------------------------------------

struct LossArgs {
  const float* out;  // [N,C] network outputs (post final activation)
  const void* target;
  int64_t N;
  int C;
  int ld;  // row pitch of out / prob / delta (targets are dense [N,C])
  int loss, final_act;
  float scale;    // 1/n_total (ce mean), 1/(n_total*C) (mse/bce mean), 1 (sum)
  float* prob;    // may be null
  float* delta;   // may be null: dloss/dz_L (through the final activation)
  Image16 delta_img;  // optional split image of delta (operand of the gradient sweep on the pair engine)
  double* partial;
};

// one warp per row; lanes stride over the C outputs
__global__ void __launch_bounds__(256) loss_forward_kernel(LossArgs a) {
  __shared__ double sm[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double block_loss = 0.0;
  for (int64_t n = (int64_t)blockIdx.x * 8 + warp; n < a.N; n += (int64_t)gridDim.x * 8) {
    const float* z = a.out + n * a.ld;
    float row = 0.f;
    if (a.loss == HF_LOSS_SOFTMAX_CE) {
      const int64_t t = static_cast<const int64_t*>(a.target)[n];
      float mx = -INFINITY;
      for (int c = lane; c < a.C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float se = 0.f;
      for (int c = lane; c < a.C; c += 32) se += expf(z[c] - mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      const float lse = mx + logf(se);
      for (int c = lane; c < a.C; c += 32) {
        const float p = expf(z[c] - lse);
        if (a.prob) a.prob[n * a.ld + c] = p;
        if (a.delta) {
          const float d = a.scale * (p - (c == t ? 1.f : 0.f));
          a.delta[n * a.ld + c] = d;
          if (a.delta_img.hi) store_image1(a.delta_img, n, c, d);
        }
      }
      row = (t >= 0 && t < a.C) ? (lse - z[t]) : 0.f;  // every lane holds the same value
    } else {
      const float* t = static_cast<const float*>(a.target) + n * a.C;
      float acc = 0.f;
      for (int c = lane; c < a.C; c += 32) {
        const float zc = z[c], tc = t[c];
        float d;
        if (a.loss == HF_LOSS_MSE) {
          const float e = zc - tc;
          acc += e * e;
          d = 2.f * e;
        } else {  // sigmoid + binary cross entropy on logits
          const float p = 1.f / (1.f + expf(-zc));
          acc += fmaxf(zc, 0.f) - zc * tc + log1pf(expf(-fabsf(zc)));
          d = p - tc;
          if (a.prob) a.prob[n * a.ld + c] = p;
        }
        if (a.delta) {
          const float dz = a.scale * d * act_d1(a.final_act, zc);
          a.delta[n * a.ld + c] = dz;
          if (a.delta_img.hi) store_image1(a.delta_img, n, c, dz);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      row = acc;
    }
    block_loss += (double)row;
  }
  if (lane == 0) sm[warp] = block_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    a.partial[blockIdx.x] = t;
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ partial, int n, double scale, double* acc) {
  __shared__ double scratch[4 * 33];
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[0] += partial[i];
  block_sum<1>(v, scratch);
  if (threadIdx.x == 0) *acc += scale * v[0];
}

// u = H_loss * Rz, row-wise, in place (then through the final activation's derivative):
//   mse: 2*scale*Rz     softmax-ce: scale*(p.Rz - p (p^T Rz))     sigmoid-bce: scale*p(1-p)*Rz
struct HessArgs {
  float* rz;  // [N,C] in: R{output}, out: u
  const float* prob;
  const float* out;
  int64_t N;
  int C;
  int ld;
  int loss, final_act;
  float scale;
  const int32_t* skip;
  // optional: R{output} arrives as split-K partial tiles part[s][n][c] (+ bias[c]) instead of in rz
  const float* part;
  int splits;
  int64_t part_stride;
  const float* bias;
  Image16 img;  // optional: split image of u (the transposed sweep reads u as a tensor-engine operand)
};

__global__ void __launch_bounds__(256) loss_hessian_kernel(HessArgs a) {
  if (a.skip && *a.skip) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t n = (int64_t)blockIdx.x * 8 + warp; n < a.N; n += (int64_t)gridDim.x * 8) {
    float* r = a.rz + n * a.ld;
    if (a.part) {
      for (int c = lane; c < a.C; c += 32) {  // each lane only ever touches its own columns: no sync needed
        float t = a.bias ? a.bias[c] : 0.f;
        for (int z = 0; z < a.splits; ++z) t += a.part[z * a.part_stride + n * a.ld + c];
        r[c] = t;
      }
    }
    if (a.loss == HF_LOSS_SOFTMAX_CE) {
      const float* p = a.prob + n * a.ld;
      float dot = 0.f;
      for (int c = lane; c < a.C; c += 32) dot += p[c] * r[c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      for (int c = lane; c < a.C; c += 32) r[c] = a.scale * p[c] * (r[c] - dot);
    } else if (a.loss == HF_LOSS_MSE) {
      for (int c = lane; c < a.C; c += 32) {
        float v = 2.f * a.scale * r[c];
        if (a.final_act != HF_ACT_NONE) v *= act_d1(a.final_act, a.out[n * a.ld + c]);
        r[c] = v;
      }
    } else {
      const float* p = a.prob + n * a.ld;
      for (int c = lane; c < a.C; c += 32) r[c] = a.scale * p[c] * (1.f - p[c]) * r[c];
    }
    if (a.img.hi)
      for (int c = lane; c < a.C; c += 32) store_image1(a.img, n, c, r[c]);  // each lane re-reads its own columns
  }
}

// ---- helpers -----------------------------------------------------------------------------------

static inline const float* weight_ptr(const Layer& l, const float* theta) {
  return l.w_off >= 0 ? theta + l.w_off : l.w_frozen;
}
static inline const float* bias_ptr(const Layer& l, const float* theta) {
  if (!l.has_bias) return nullptr;
  return l.b_off >= 0 ? theta + l.b_off : l.b_frozen;
}

// weight / direction operand of layer l as the kernels should read it: in place, or the 16-byte-pitched copy
static inline Operand w_operand(const hf_lin* lin, int l, const float* theta, bool k_contig) {
  const Layer& L = lin->net->L[l];
  const float* p = lin->wpad[l] ? lin->wpad[l] : weight_ptr(L, theta);
  const int ld = lin->wpad[l] ? pad4(L.in) : L.in;
  Operand op = k_contig ? Operand{p, ld, 1} : Operand{p, 1, ld};
  if (lin->use_images && lin->w_img[l].hi) op.img = Image16{lin->w_img[l].hi, lin->w_img[l].plane, pad8(L.in)};
  return op;
}
static inline Operand v_operand(const hf_lin* lin, int l, const float* v, bool k_contig) {
  const Layer& L = lin->net->L[l];
  const float* p = lin->vpad[l] ? lin->vpad[l] : v + L.w_off;
  const int ld = lin->vpad[l] ? pad4(L.in) : L.in;
  Operand op = k_contig ? Operand{p, ld, 1} : Operand{p, 1, ld};
  if (lin->use_images && lin->v_img[l].hi) op.img = Image16{lin->v_img[l].hi, lin->v_img[l].plane, pad8(L.in)};
  return op;
}

struct SplitPlan {
  int splits;
  int k_per_split;
};

// split the batch dimension of a weight-gradient contraction so the grid fills the machine
static SplitPlan plan_split(int M, int N, int64_t K, bool tensor_tiles) {
  constexpr int kUnit = 32;  // K granularity both engines accept (tcgen05 k-block = 32, SIMT k-tile = 16)
  int64_t tiles, want;
  if (tensor_tiles) {  // 128x128 tiles, one CTA per SM: aim for just under one wave
    tiles = (int64_t)((M + 127) / 128) * ((N + 127) / 128);
    want = sm_count() / tiles;
  } else {
    const TileChoice t = choose_tile(M, N);
    tiles = (int64_t)((M + t.bm - 1) / t.bm) * ((N + t.bn - 1) / t.bn);
    want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
  }
  const int64_t ktiles = (K + kUnit - 1) / kUnit;
  int64_t max_split = ktiles / 4;  // at least 128 samples per split
  if (max_split < 1) max_split = 1;
  if (want > max_split) want = max_split;
  // The tensor core adds each 8-wide product group into the FP32 accumulator with truncation, a bias that grows
  // linearly with the number of accumulation steps (measured against float64: 3e-5 relative after 3000 samples,
  // ~1e-4 after 30 000).  Batch contractions are therefore cut into partial sums of at most 1024 samples, which
  // the reduction kernel adds in properly rounded FP32.
  if (tensor_tiles && want < (ktiles + 31) / 32) want = (ktiles + 31) / 32;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  const int64_t kt_per = (ktiles + want - 1) / want;
  SplitPlan p;
  p.k_per_split = (int)(kt_per * kUnit);
  p.splits = (int)((ktiles + kt_per - 1) / kt_per);
  return p;
}

// The same for the 256x256 CTA-pair tiles: the split count that minimises waves x k-blocks, subject to the 1024-sample
// accumulation rule above.  `us` receives the modelled time (gemm_tc.cuh: tc2_estimate).
static SplitPlan plan_split_pair(int M, int N, int64_t K, double* us) {
  const int64_t ktiles = (K + 31) / 32;
  const int64_t lo = std::max<int64_t>(1, (ktiles + 31) / 32), hi = std::max<int64_t>(lo, std::min<int64_t>(64, ktiles / 4));
  SplitPlan best = {1, (int)(ktiles * 32)};
  double best_us = 1e30;
  for (int64_t want = lo; want <= hi; ++want) {
    const int64_t kt_per = (ktiles + want - 1) / want;
    const int splits = (int)((ktiles + kt_per - 1) / kt_per);
    const double t = tc2_estimate(M, N, (int)K, 1, splits).us_pair + 0.05 * splits;  // the reduction reads every partial
    if (t < best_us) best_us = t, best = SplitPlan{splits, (int)(kt_per * 32)};
  }
  if (us) *us = best_us;
  return best;
}

// would a weight-gradient contraction [out,in] over `batch` samples run on the tensor-core tiles?
static bool weight_on_tensor(const hf_net* net, int out, int in, int64_t batch, int square) {
  return net->engine == 1 && !square && in % 4 == 0 && (int64_t)out * in * batch >= kTcMinWork;
}

static int colsum_plan(int64_t rows) {
  int64_t s = (rows + 511) / 512;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return (int)s;
}

static Operand op_kc(const float* p, int64_t ld) { return Operand{p, ld, 1}; }   // [MN,K], K contiguous
static Operand op_mnc(const float* p, int64_t ld) { return Operand{p, 1, ld}; }  // [K,MN], MN contiguous

static GemmArgs blank_gemm() {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.alpha = 1.f;
  g.split_k = 1;
  return g;
}

// should this contraction (operand images present) run on the CTA-pair tiles rather than the 128x128 tiles?
static bool prefer_pair(const GemmArgs& g) {
  if (tc2_mode() == 0 || !tc2_supported(g)) return false;
  if (tc2_mode() == 2) return true;
  if ((int64_t)g.M * g.N * g.K < kTcMinWork || g.N < 96 || g.M < 192) return false;  // narrow tiles are mostly padding
  const Tc2Choice c = tc2_estimate(g.M, g.N, g.K, g.n_pairs, std::max(1, g.split_k));
  return c.us_pair < 0.9 * c.us_single;
}

// Dispatch one contraction: CTA-pair tensor tiles on pre-split images where they exist and pay, 128x128 tensor tiles
// where the shape meets the TMA rules, FP32 SIMT tiles otherwise.  Whichever engine runs, the split image of C is
// written when the caller asks for one (g.c_img): by the tensor engines' epilogues, or by a split pass after the
// SIMT kernel.
static int run_gemm(const hf_net* net, const GemmArgs& g, cudaStream_t stream, bool* on_tensor = nullptr, int pair_hint = -1) {
  const bool pair = net->engine == 1 && (pair_hint < 0 ? prefer_pair(g) : pair_hint == 1);
  const bool tc = net->engine == 1 && (pair || tc_supported(g));
  if (on_tensor) *on_tensor = tc;
  if (pair) return launch_gemm_tc2(g, stream);
  if (tc) return launch_gemm_tc(g, stream);
  int rc = launch_gemm_simt(g, stream);
  if (rc || !g.c_img.hi || g.split_k > 1) return rc;
  SplitTable t;
  t.count = 1, t.skip = g.skip;
  t.seg[0] = SplitSegment{g.C, g.M, g.N, g.ldc, nullptr, 0, g.c_img};
  return launch_split(t, stream);
}

// Gradient slices of one layer: out_W[out,in] (+)= scale * sum_pairs A_s^T B_s over the batch (split-K partials),
// out_b[out] (+)= scale * column sums of d.  The column sums either come for free from the tensor-core kernel that
// produced d (`col_tiles` partial rows already in lin->colbuf) or from one colsum launch; ONE launch then reduces
// both partial sets in fixed order (deterministic) into the flat vector.
static int layer_gradient(hf_lin* lin, const Layer& L, int64_t rows, int M, int N, int n_pairs, const Operand* A, const Operand* B, int square,
                          float* out_w, const float* d, int ld_d, int col_tiles, float* colbuf, float* out_b, float scale,
                          int accumulate, const int32_t* skip, cudaStream_t stream, bool main_scratch = false) {
  int splits_w = 0, splits_b = 0;
  float* const partial = main_scratch ? lin->partial_main : lin->partial;
  if (out_w) {
    GemmArgs g = blank_gemm();
    g.M = M, g.N = N, g.K = (int)rows, g.n_pairs = n_pairs;
    for (int s = 0; s < n_pairs; ++s) g.A[s] = A[s], g.B[s] = B[s];
    int gemm_square = square;  // (the bias slice below still needs `square` for its column sums)
    if (square && lin->sq[0] && n_pairs == 1 && A[0].s_mn == 1 && B[0].s_mn == 1 && (int64_t)M * N * rows >= kTcMinWork) {
      // Fisher diagonal as a tensor-core contraction: square the two operands once (one launch), contract as usual
      SplitTable t;
      t.count = 2, t.skip = skip;
      Operand sqop[2];
      for (int w = 0; w < 2; ++w) {
        const Operand& src = w ? B[0] : A[0];
        const int cols = w ? N : M;
        const Image16 im = lin->sq_img[w].hi ? Image16{lin->sq_img[w].hi, lin->sq_img[w].plane, pad8(cols)} : Image16{nullptr, 0, 0};
        t.seg[w] = SplitSegment{src.ptr, rows, cols, src.s_k, lin->sq[w], pad4(cols), im, 1};
        sqop[w] = Operand{lin->sq[w], 1, pad4(cols), im};
      }
      int rcs = launch_split(t, stream);
      if (rcs) return rcs;
      g.A[0] = sqop[0], g.B[0] = sqop[1];
      gemm_square = 0;
    }
    g.square = gemm_square;
    SplitPlan sp = plan_split(M, N, rows, weight_on_tensor(lin->net, M, N, rows, gemm_square));
    int pair_hint = 0;
    if (lin->net->engine == 1 && tc2_mode() != 0 && tc2_supported(g)) {
      // both engines can take it: compare the modelled times of their own best split
      double us_pair = 0.0;
      const SplitPlan sp2 = plan_split_pair(M, N, rows, &us_pair);
      const double us_single = tc2_estimate(M, N, (int)rows, 1, sp.splits).us_single + 0.05 * sp.splits;
      if (tc2_mode() == 2 || us_pair < 0.9 * us_single) sp = sp2, pair_hint = 1;
    }
    HF_REQUIRE((size_t)sp.splits * M * N <= (main_scratch ? lin->partial_main_floats : lin->partial_floats), HF_ERR_WORKSPACE,
               "split-K scratch too small");
    g.C = partial, g.ldc = N;
    g.epi = EPI_STORE;
    g.split_k = sp.splits, g.k_per_split = sp.k_per_split;
    g.skip = skip;
    int rc = run_gemm(lin->net, g, stream, nullptr, pair_hint);
    if (rc) return rc;
    splits_w = sp.splits;
  }
  if (out_b) {
    if (col_tiles > 0) {
      splits_b = col_tiles;
    } else {
      splits_b = colsum_plan(rows);
      const int rows_per = (int)((rows + splits_b - 1) / splits_b);
      HF_REQUIRE((size_t)splits_b <= lin->col_rows, HF_ERR_WORKSPACE, "column-sum scratch too small");
      colsum_kernel<<<dim3((M + 31) / 32, splits_b), dim3(32, 8), 0, stream>>>(d, rows, M, ld_d, rows_per, square,
                                                                                colbuf, skip);
      HF_LAUNCH_CHECK();
    }
  }
  const int64_t count_w = out_w ? (int64_t)M * N : 0, count_b = out_b ? M : 0;
  if (count_w + count_b == 0) return HF_OK;
  int64_t blocks = (count_w / 4 + count_b + 255) / 256 + 1;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  const int taps = L.unfold ? L.geom.kh * L.geom.kw : 0;  // unfolded operands are tap-major: write back channel-major
  reduce_partials2_kernel<<<(unsigned)blocks, 256, 0, stream>>>(partial, splits_w, count_w, out_w, colbuf, splits_b,
                                                               count_b, out_b, scale, accumulate, skip, L.geom.cin, taps);
  HF_LAUNCH_CHECK();
  return HF_OK;
}

// Empirical-Fisher slices of a layer applied at S > 1 positions per sample (convolutions): conv.cuh, conv_fisher_*.
// cot = per-sample output cotangents [N*S, M], U = the layer's (unfolded) input [N*S, K].
static int conv_fisher_gradient(hf_lin* lin, const Layer& L, int M, int K, const float* cot, int ld_c, const float* U, int ld_u,
                                float* out_w, float* colbuf, float* out_b, float scale, int accumulate, const int32_t* skip,
                                cudaStream_t stream) {
  const int S = L.s_out;
  const int64_t n = lin->N;
  int groups_w = 0, groups_b = 0;
  if (out_w) {
    const int64_t tiles = (int64_t)((M + 63) / 64) * ((K + 63) / 64);
    int64_t want = std::max<int64_t>(1, (2 * sm_count() + tiles - 1) / tiles);          // ~two CTAs per SM
    want = std::min<int64_t>(want, (int64_t)(lin->partial_floats / ((size_t)M * K)));   // what the split-K scratch holds
    want = std::max<int64_t>(1, std::min<int64_t>(want, n));
    const int per = (int)((n + want - 1) / want);
    groups_w = (int)((n + per - 1) / per);
    HF_REQUIRE((size_t)groups_w * M * K <= lin->partial_floats, HF_ERR_WORKSPACE, "Fisher scratch too small");
    conv_fisher_w_kernel<<<dim3((K + 63) / 64, (M + 63) / 64, groups_w), 256, 0, stream>>>(cot, ld_c, U, ld_u, n, S, M, K, per, lin->partial,
                                                                                           skip);
    HF_LAUNCH_CHECK();
  }
  if (out_b) {
    int64_t want = std::min<int64_t>(std::min<int64_t>((int64_t)lin->col_rows, n), std::max<int64_t>(1, 4 * sm_count() / ((M + 31) / 32)));
    const int per = (int)((n + want - 1) / want);
    groups_b = (int)((n + per - 1) / per);
    conv_fisher_b_kernel<<<dim3((M + 31) / 32, groups_b), dim3(32, 8), 0, stream>>>(cot, ld_c, n, S, M, per, colbuf, skip);
    HF_LAUNCH_CHECK();
  }
  const int64_t count_w = out_w ? (int64_t)M * K : 0, count_b = out_b ? M : 0;
  if (count_w + count_b == 0) return HF_OK;
  int64_t blocks = (count_w / 4 + count_b + 255) / 256 + 1;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  const int taps = L.unfold ? L.geom.kh * L.geom.kw : 0;
  reduce_partials2_kernel<<<(unsigned)blocks, 256, 0, stream>>>(lin->partial, groups_w, count_w, out_w, colbuf, groups_b, count_b, out_b,
                                                               scale, accumulate, skip, L.geom.cin, taps);
  HF_LAUNCH_CHECK();
  return HF_OK;
}

static float loss_scale(const hf_net* net, int64_t n_total) {
  if (net->reduction == HF_RED_SUM) return 1.f;
  if (net->loss == HF_LOSS_SOFTMAX_CE) return (float)(1.0 / (double)n_total);
  return (float)(1.0 / ((double)n_total * (double)net->classes));
}

// Operand forms of the direction v for layers [l_begin, l_end): the 16-byte-pitched FP32 copy where the flat slice
// cannot be addressed by TMA in place, and the split-precision image for the pair engine -- all layers in one launch.
static int prepare_direction(hf_lin* lin, const float* v, int l_begin, int l_end, const int32_t* skip, cudaStream_t stream) {
  const hf_net* net = lin->net;
  SplitTable t;
  t.count = 0, t.skip = skip;
  for (int l = l_begin; l < l_end; ++l) {
    const Layer& L = net->L[l];
    if (L.w_off < 0) continue;
    hf_lin::ImgBuf& vi = lin->v_img[l];
    const bool img = lin->use_images && vi.hi;
    if (!lin->vpad[l] && !img) continue;
    if (img) vi.base = lin->vpad[l] ? lin->vpad[l] : v + L.w_off;
    t.seg[t.count++] = SplitSegment{v + L.w_off, L.out, L.in, L.in, lin->vpad[l], pad4(L.in),
                                    img ? Image16{vi.hi, vi.plane, pad8(L.in)} : Image16{nullptr, 0, 0}, 0,
                                    L.geom.cin, L.unfold ? L.geom.kh * L.geom.kw : 0};
    if (t.count == kMaxSplitSegments) {
      int rc = launch_split(t, stream);
      if (rc) return rc;
      t.count = 0;
    }
  }
  return launch_split(t, stream);
}

// R-op forward: R{output} = J v into lin->buf[which]; returns the buffer index holding it
// (with `below_head` the last layer is left to the fused head and *below_head receives R{a_{L-2}})
static int rop_forward(hf_lin* lin, const float* theta, const float* v, bool hessian, const int32_t* skip,
                       cudaStream_t stream, int* out_buf, const float** below_head = nullptr) {
  const hf_net* net = lin->net;
  const int nl = (int)net->L.size();
  const int l_end = nl - (below_head ? 1 : 0);
  const float* cur = nullptr;
  int which = 0;
  int rcp = prepare_direction(lin, v, net->first_trainable, l_end, skip, stream);
  if (rcp) return rcp;
  for (int l = net->first_trainable; l < l_end; ++l) {
    const Layer& L = net->L[l];
    const int ld_out = pad4(L.out);
    if (L.kind == HF_LAYER_AVGPOOL) {
      // the tangent of a global average pool is the pool of the tangent
      if (cur) {
        const bool keep_ra = hessian && l < nl - 1;
        float* dst = keep_ra ? lin->ra[l] : lin->buf[which];
        avgpool_kernel<<<conv_blocks(lin->N * (int64_t)L.out), 256, 0, stream>>>(cur, ld_out, dst, lin->N, L.s_in, L.out,
                                                                                  image_for(lin, dst, L.out), skip);
        HF_LAUNCH_CHECK();
        cur = dst;
        if (!keep_ra) which ^= 1;
      }
      continue;
    }
    int ld_in = 0;
    const float* a_in = layer_input(lin, l, &ld_in);
    GemmArgs g = blank_gemm();
    g.M = (int)rows_out(lin, l), g.N = L.out, g.K = L.in;
    int np = 0;
    if (L.w_off >= 0) {
      g.A[np] = with_image(lin, op_kc(a_in, ld_in), L.in), g.B[np] = v_operand(lin, l, v, true);
      ++np;
    }
    if (cur) {
      const float* cur_mat = cur;
      if (L.unfold) {  // the tangent of the unfolded input is the unfolded tangent (kept per layer for Hessian products)
        float* ru = (hessian && lin->RU[l]) ? lin->RU[l] : lin->ru;
        im2col_kernel<<<conv_blocks(rows_out(lin, l) * (pad4(L.in) / 4)), 256, 0, stream>>>(cur, pad4(L.geom.cin), ru, pad4(L.in), lin->N,
                                                                                           L.geom, image_for(lin, ru, L.in), skip);
        HF_LAUNCH_CHECK();
        cur_mat = ru;
      }
      g.A[np] = with_image(lin, op_kc(cur_mat, pad4(L.in)), L.in), g.B[np] = w_operand(lin, l, theta, true);
      ++np;
    }
    float* dst = (hessian && l < nl - 1) ? lin->ra[l] : lin->buf[which];
    lin->fwd_splits = 0;
    lin->loss_fused = false;
    if (l == nl - 1 && np > 0 && lin->partial_fwd && L.act == HF_ACT_NONE && net->engine == 1) {
      // Narrow output layer (10 classes): a 128x128 tensor tile per 128 rows would leave most SMs idle, so the
      // contraction is split over K instead and the loss-Hessian kernel sums the partial tiles (+ bias tangent).
      const int row_tiles = (int)((lin->N + 127) / 128);
      const int kb = (L.in + 31) / 32;
      int splits = std::min(std::min(sm_count() / row_tiles, kb / 4), lin->fwd_splits_max);
      GemmArgs t = g;
      t.n_pairs = np, t.C = lin->partial_fwd, t.ldc = ld_out, t.epi = EPI_STORE;
      if (splits >= 2 && tc_supported(t)) {
        const int kb_per = (kb + splits - 1) / splits;
        t.k_per_split = kb_per * 32, t.split_k = (kb + kb_per - 1) / kb_per;
        t.skip = skip;
        int rc = launch_gemm_tc(t, stream);
        if (rc) return rc;
        lin->fwd_splits = t.split_k;
        lin->fwd_bias = (L.has_bias && L.b_off >= 0) ? v + L.b_off : nullptr;
        cur = dst;
        which ^= 1;
        continue;
      }
    }
    if (np == 0) {
      // a frozen layer fed by a zero tangent contributes only its (frozen) nothing: R{z} = 0
      HF_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * rows_out(lin, l) * ld_out, stream));
      const Image16 im = image_for(lin, dst, L.out);
      if (im.hi) {
        HF_CUDA(cudaMemsetAsync(im.hi, 0, sizeof(uint16_t) * rows_out(lin, l) * im.ld, stream));
        HF_CUDA(cudaMemsetAsync(im.hi + im.plane, 0, sizeof(uint16_t) * rows_out(lin, l) * im.ld, stream));
      }
    } else {
      g.n_pairs = np;
      g.C = dst, g.ldc = ld_out;
      g.epi = EPI_BIAS_DACT, g.act = L.act;
      g.bias = (L.has_bias && L.b_off >= 0) ? v + L.b_off : nullptr;
      g.aux = lin->a[l], g.ldaux = ld_out;
      if (l == nl - 1 && L.act == HF_ACT_NONE && (net->loss == HF_LOSS_SIGMOID_BCE || net->loss == HF_LOSS_MSE)) {
        // A diagonal loss Hessian rides in the epilogue of the last R-op contraction instead of a pass of its own:
        // sigmoid-BCE  u = scale p(1-p) Rz  is "act' of a sigmoid" on the stored probabilities, MSE  u = 2 scale Rz.
        g.alpha = loss_scale(net, lin->n_total) * (net->loss == HF_LOSS_MSE ? 2.f : 1.f);
        if (net->loss == HF_LOSS_SIGMOID_BCE) g.act = HF_ACT_SIGMOID, g.aux = lin->prob;
        lin->loss_fused = true;
      }
      g.C2 = (hessian && curved(L.act)) ? lin->rz[l] : nullptr;
      g.c_img = image_for(lin, dst, L.out);  // the next layer's R-op (or the transposed sweep) reads it as an operand
      g.skip = skip;
      int rc = run_gemm(net, g, stream);
      if (rc) return rc;
    }
    cur = dst;
    if (!(hessian && l < nl - 1)) which ^= 1;
  }
  // cur is the last layer's output tangent and lives in a ping-pong buffer
  *out_buf = (cur == lin->buf[0]) ? 0 : 1;
  if (below_head) *below_head = cur;
  return HF_OK;
}

static int apply_loss_hessian(hf_lin* lin, float* rz, const int32_t* skip, cudaStream_t stream) {
  if (lin->loss_fused) return HF_OK;  // done in the epilogue of the last R-op contraction
  const hf_net* net = lin->net;
  HessArgs h;
  h.rz = rz, h.prob = lin->prob, h.out = lin->a.back();
  h.N = lin->N, h.C = net->classes, h.ld = pad4(net->classes), h.loss = net->loss, h.final_act = net->L.back().act;
  h.scale = loss_scale(net, lin->n_total);
  h.skip = skip;
  h.img = image_for(lin, rz, net->classes);
  h.part = lin->fwd_splits > 0 ? lin->partial_fwd : nullptr;
  h.splits = lin->fwd_splits, h.part_stride = lin->N * (int64_t)pad4(net->classes), h.bias = lin->fwd_bias;
  int64_t blocks = (lin->N + 7) / 8;
  if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
  loss_hessian_kernel<<<(unsigned)blocks, 256, 0, stream>>>(h);
  HF_LAUNCH_CHECK();
  return HF_OK;
}

// Fused output head (head.cuh): R-op through the last layer, loss Hessian, transposed product down to cot[L-2] and the
// per-CTA partials of the head's own gradient slices, in one launch.  `ra` = R{a_{L-2}} from rop_forward.
static int ggn_head(hf_lin* lin, const float* theta, const float* v, const float* ra, const int32_t* skip,
                    cudaStream_t stream) {
  const hf_net* net = lin->net;
  const int nl = (int)net->L.size();
  const Layer& L = net->L[nl - 1];
  const Layer& Lp = net->L[nl - 2];
  HeadArgs h;
  h.a = lin->a[nl - 2], h.ra = ra;
  h.W = theta + L.w_off, h.V = v + L.w_off;
  h.vb = (L.has_bias && L.b_off >= 0) ? v + L.b_off : nullptr;
  h.prob = net->loss == HF_LOSS_MSE ? nullptr : lin->prob;
  h.cot = lin->cot[nl - 2];
  h.partW = lin->head_partW;
  h.partB = (L.has_bias && L.b_off >= 0) ? lin->head_partB : nullptr;
  h.partCol = (Lp.has_bias && Lp.b_off >= 0) ? lin->colbuf[nl - 2] : nullptr;
  h.N = lin->N, h.D = L.in, h.Dp = lin->head.Dp, h.C = L.out, h.ld = pad4(L.in), h.ldc = pad4(L.out);
  h.loss = net->loss, h.act_prev = Lp.act;
  h.scale = loss_scale(net, lin->n_total);
  h.rows_per_cta = lin->head.rows_per_cta;
  h.skip = skip;
  int rc = launch_head(h, lin->head, stream);
  const Image16 img = image_for(lin, h.cot, L.in);
  if (rc || !img.hi) return rc;
  SplitTable t;  // the head kernel stores FP32 only: give the pair engine its operand forms of cot[L-2]
  t.count = 1, t.skip = skip;
  t.seg[0] = SplitSegment{h.cot, lin->N, L.in, pad4(L.in), nullptr, 0, img};
  return launch_split(t, stream);
}

enum BackMode { BACK_GRADIENT, BACK_GGN, BACK_FISHER, BACK_HESSIAN };

// Transposed sweep from the output signal `top` ([N,C]) down to the first trainable layer.
//   GRADIENT: out (+)= J^T top; with HF_LIN_HESSIAN also stores delta_l and dL/da_l
//   GGN:      out (+)= J^T top
//   FISHER:   out (+)= scale * sum_n (per-sample gradient)^2
//   HESSIAN:  top = R{delta_L}; adds the delta_l^T R{a_{l-1}} and delta_l V_l terms and the act'' term
// phase -1: the whole sweep.  phase 0: everything except the parameter gradient of the first trainable layer (whose
// cotangent is left in lin->pending_cur).  phase 1: only that gradient.  The split lets a data-parallel caller start
// the all-reduce of the upper layers' slices while the (largest, last) first-layer gradient is still being formed.
static int backward_sweep(hf_lin* lin, const float* theta, const float* v, const float* top, float* out, int accumulate,
                          BackMode mode, const int32_t* skip, cudaStream_t stream, int phase = -1, bool head = false) {
  const hf_net* net = lin->net;
  const int nl = (int)net->L.size();
  const bool keep = mode == BACK_GRADIENT && (lin->flags & HF_LIN_HESSIAN);
  const int square = mode == BACK_FISHER;
  float scale = 1.f;
  if (mode == BACK_FISHER && net->reduction == HF_RED_MEAN) scale = (float)lin->n_total;
  // The sweep is a chain of data products (cot[l] -> cot[l-1]) on the caller's stream; the parameter gradients of
  // layer l only need cot[l], so they run on the side stream while the chain moves on.  Every cotangent has its
  // own buffer, so there is no write-after-read hazard between the two streams.
  // (the Fisher sweep runs once per step and shares one pair of squared-operand buffers between its layers: no fork)
  const bool fork = lin->side != nullptr && phase != 1 && mode != BACK_FISHER;
  cudaStream_t gstream = fork ? lin->side : stream;
  const float* cur = top;
  int cur_col_tiles = 0;  // > 0: the kernel that produced `cur` also left its column sums in colbuf[l]
  int l_start = nl - 1;
  if (phase == 1) l_start = net->first_trainable, cur = lin->pending_cur, cur_col_tiles = lin->pending_cols;
  if (fork) HF_CUDA(cudaEventRecord(lin->ev[nl - 1], stream));
  for (int l = l_start; l >= net->first_trainable; --l) {
    if (phase == 0 && l == net->first_trainable) {
      lin->pending_cur = cur, lin->pending_cols = cur_col_tiles;
      break;
    }
    if (head && l == nl - 1) {
      // the fused head already produced cot[l-1] (+ its column sums) and the partials of this layer's own slices
      const Layer& Lh = net->L[l];
      const bool has_b = Lh.has_bias && Lh.b_off >= 0;
      if (fork) HF_CUDA(cudaStreamWaitEvent(gstream, lin->ev[l], 0));
      const int64_t count_w = (int64_t)Lh.out * Lh.in, count_b = has_b ? Lh.out : 0;
      int64_t blocks = (count_w + count_b + 31) / 32;
      if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
      reduce_tall_kernel<<<(unsigned)blocks, 256, 0, gstream>>>(lin->head_partW, lin->head.ctas, count_w, out + Lh.w_off,
                                                               lin->head_partB, lin->head.ctas, count_b,
                                                               has_b ? out + Lh.b_off : nullptr, scale, accumulate, skip);
      HF_LAUNCH_CHECK();
      if (fork) HF_CUDA(cudaEventRecord(lin->ev[l - 1], stream));
      const Layer& Lp = net->L[l - 1];
      cur = lin->cot[l - 1];
      cur_col_tiles = (Lp.has_bias && Lp.b_off >= 0) ? lin->head.ctas : 0;
      continue;
    }
    const Layer& L = net->L[l];
    const int ld_out = pad4(L.out);
    if (L.kind == HF_LAYER_AVGPOOL) {
      // transposed global average pool: every position of the map receives cot / (h*w), through the activation
      // derivative of the layer below (which the kernel producing cot[l-1] always applies)
      if (l - 1 < net->first_trainable) break;  // nothing trainable below
      const Layer& Lp = net->L[l - 1];
      float* dst = keep ? lin->delta[l - 1] : lin->cot[l - 1];
      const bool second = mode == BACK_HESSIAN && curved(Lp.act);
      unpool_kernel<<<conv_blocks(rows_out(lin, l - 1) * (int64_t)L.out), 256, 0, stream>>>(
          cur, ld_out, lin->a[l - 1], Lp.act, dst, lin->N, L.s_in, L.out, image_for(lin, dst, L.out), skip,
          (keep && curved(Lp.act)) ? lin->ga[l - 1] : nullptr, second ? lin->ga[l - 1] : nullptr, second ? lin->rz[l - 1] : nullptr);
      HF_LAUNCH_CHECK();
      if (fork) HF_CUDA(cudaEventRecord(lin->ev[l - 1], stream));
      cur = dst, cur_col_tiles = 0;
      continue;
    }
    int ld_in = 0;
    const float* a_in = layer_input(lin, l, &ld_in);
    const int64_t rows = rows_out(lin, l);
    {
      Operand A[2], B[2];
      int np = 0;
      A[np] = with_image(lin, op_mnc(cur, ld_out), L.out), B[np] = with_image(lin, op_mnc(a_in, ld_in), L.in), ++np;
      if (mode == BACK_HESSIAN && l > net->first_trainable) {
        // delta_l^T R{input of the layer}: the kept unfolded tangent for an unfolded convolution
        A[np] = op_mnc(lin->delta[l], ld_out), B[np] = op_mnc(L.unfold ? lin->RU[l] : lin->ra[l - 1], ld_in), ++np;
      }
      const bool has_b = L.has_bias && L.b_off >= 0;
      // The first trainable layer's gradient is the end of the chain: the caller's stream has nothing left to do, so
      // it runs there (own split-K scratch) while the side stream finishes the upper layers' reductions.
      const bool on_main = fork && l == net->first_trainable && lin->partial_main != nullptr;
      if (fork && !on_main) HF_CUDA(cudaStreamWaitEvent(gstream, lin->ev[l], 0));
      int rc;
      if (square && L.s_out > 1)  // a layer applied at several positions per sample: the square sits outside their sum
        rc = conv_fisher_gradient(lin, L, L.out, L.in, cur, ld_out, a_in, ld_in, L.w_off >= 0 ? out + L.w_off : nullptr, lin->colbuf[l],
                                  has_b ? out + L.b_off : nullptr, scale, accumulate, skip, gstream);
      else
        rc = layer_gradient(lin, L, rows, L.out, L.in, np, A, B, square, L.w_off >= 0 ? out + L.w_off : nullptr, cur, ld_out,
                            cur_col_tiles, lin->colbuf[l], has_b ? out + L.b_off : nullptr, scale, accumulate, skip,
                            on_main ? stream : gstream, on_main);
      if (rc) return rc;
    }
    cur_col_tiles = 0;
    if (l > net->first_trainable) {
      const Layer& Lp = net->L[l - 1];
      GemmArgs g = blank_gemm();
      g.M = (int)rows, g.N = L.in, g.K = L.out;
      int np = 0;
      g.A[np] = with_image(lin, op_kc(cur, ld_out), L.out), g.B[np] = w_operand(lin, l, theta, false), ++np;
      if (L.unfold) {
        // convolution: dU = cot W is the cotangent of the UNFOLDED input; fold it back onto the input map (gather
        // form) and apply the activation derivative of the layer below there
        if (mode == BACK_HESSIAN && L.w_off >= 0) {
          g.A[np] = op_kc(lin->delta[l], ld_out), g.B[np] = v_operand(lin, l, v, false), ++np;
        }
        g.n_pairs = np;
        g.C = lin->du, g.ldc = pad4(L.in);
        g.epi = EPI_STORE;
        g.skip = skip;
        int rc = run_gemm(net, g, stream);
        if (rc) return rc;
        float* dst = keep ? lin->delta[l - 1] : lin->cot[l - 1];
        const bool second = mode == BACK_HESSIAN && curved(Lp.act);
        fold_kernel<<<conv_blocks(rows_in(lin, l) * ((L.geom.cin + 3) / 4)), 256, 0, stream>>>(
            lin->du, pad4(L.in), lin->a[l - 1], pad4(L.geom.cin), Lp.act, dst, lin->N, L.geom, image_for(lin, dst, L.geom.cin), skip,
            (keep && curved(Lp.act)) ? lin->ga[l - 1] : nullptr, second ? lin->ga[l - 1] : nullptr, second ? lin->rz[l - 1] : nullptr);
        HF_LAUNCH_CHECK();
        if (fork) HF_CUDA(cudaEventRecord(lin->ev[l - 1], stream));
        cur = dst, cur_col_tiles = 0;
        continue;
      }
      const int ld_in = pad4(L.in);  // (= the pitch of a[l-1] and cot[l-1])
      if (mode == BACK_HESSIAN && L.w_off >= 0) {
        g.A[np] = op_kc(lin->delta[l], ld_out), g.B[np] = v_operand(lin, l, v, false), ++np;
      }
      g.n_pairs = np;
      float* dst = keep ? lin->delta[l - 1] : lin->cot[l - 1];
      g.C = dst, g.ldc = ld_in;
      g.c_img = image_for(lin, dst, L.in);  // operand of the next product down the sweep (kept current in every mode)
      g.act = Lp.act;
      g.aux = lin->a[l - 1], g.ldaux = ld_in;
      if (mode == BACK_HESSIAN) {
        g.epi = EPI_DACT_H;
        if (curved(Lp.act)) g.h_ga = lin->ga[l - 1], g.h_rz = lin->rz[l - 1];
      } else {
        g.epi = EPI_DACT;
        g.C2 = (keep && curved(Lp.act)) ? lin->ga[l - 1] : nullptr;
      }
      g.skip = skip;
      const int row_tiles = (int)((rows + 127) / 128);
      const bool want_cols = !square && Lp.has_bias && Lp.b_off >= 0 && (size_t)row_tiles <= lin->col_rows;
      g.colpart = want_cols ? lin->colbuf[l - 1] : nullptr;
      bool on_tensor = false;
      int rc = run_gemm(net, g, stream, &on_tensor);
      if (rc) return rc;
      if (fork) HF_CUDA(cudaEventRecord(lin->ev[l - 1], stream));
      cur_col_tiles = (want_cols && on_tensor) ? row_tiles : 0;
      cur = dst;
    }
  }
  if (fork) {
    HF_CUDA(cudaEventRecord(lin->join, gstream));
    HF_CUDA(cudaStreamWaitEvent(stream, lin->join, 0));
  }
  return HF_OK;
}

}  // namespace hf

using namespace hf;

extern "C" {

int hf_net_create(const hf_layer_desc* layers, int32_t n_layers, int32_t loss, int32_t reduction, int64_t n_params,
                  hf_net_t** out) {
  HF_REQUIRE(layers && out && n_layers >= 1, HF_ERR_INVALID, "hf_net_create: need at least one layer");
  HF_REQUIRE(loss >= HF_LOSS_MSE && loss <= HF_LOSS_SIGMOID_BCE, HF_ERR_INVALID, "hf_net_create: unknown loss %d", loss);
  HF_REQUIRE(reduction == HF_RED_MEAN || reduction == HF_RED_SUM, HF_ERR_INVALID, "hf_net_create: unknown reduction");
  hf_net* net = new (std::nothrow) hf_net();
  HF_REQUIRE(net, HF_ERR_INVALID, "hf_net_create: out of host memory");
  net->loss = loss, net->reduction = reduction, net->P = n_params;
  net->first_trainable = -1, net->max_width = 0, net->engine = 0, net->has_relu = false;
  net->has_conv = false, net->max_s = 1, net->max_act = 0;
  for (int i = 0; i < n_layers; ++i) {
    const hf_layer_desc& d = layers[i];
    bool ok = d.in_features > 0 && d.out_features > 0 && d.act >= HF_ACT_NONE && d.act <= HF_ACT_TANH;
    ok = ok && (i == 0 || d.kind != HF_LAYER_LINEAR || d.in_features == layers[i - 1].out_features);
    ok = ok && (d.kind == HF_LAYER_AVGPOOL ||
                (d.w_offset >= 0 ? d.w_offset + (int64_t)d.in_features * d.out_features <= n_params : d.d_w_frozen != nullptr));
    if (d.has_bias) ok = ok && (d.b_offset >= 0 ? d.b_offset + d.out_features <= n_params : d.d_b_frozen != nullptr);
    if (!ok) {
      delete net;
      HF_REQUIRE(false, HF_ERR_INVALID, "hf_net_create: layer %d is inconsistent", i);
    }
    Layer l{d.in_features, d.out_features, d.act, d.has_bias, d.w_offset, d.has_bias ? d.b_offset : -1, d.d_w_frozen,
            d.d_b_frozen};
    l.kind = d.kind, l.s_in = l.s_out = 1, l.unfold = false;
    l.geom = ConvGeom{d.c_in, d.h_in, d.w_in, d.k_h, d.k_w, d.stride, d.pad, d.h_out, d.w_out};
    if (l.kind != HF_LAYER_LINEAR) {
      const ConvGeom& g = l.geom;
      bool cok = g.cin > 0 && g.hin > 0 && g.win > 0;
      if (l.kind == HF_LAYER_CONV2D) {
        cok = cok && g.kh > 0 && g.kw > 0 && g.stride > 0 && g.pad >= 0 && l.in == g.cin * g.kh * g.kw &&
              g.hout == (g.hin + 2 * g.pad - g.kh) / g.stride + 1 && g.wout == (g.win + 2 * g.pad - g.kw) / g.stride + 1 && g.hout > 0 && g.wout > 0;
        l.s_in = g.hin * g.win, l.s_out = g.hout * g.wout;
        l.unfold = !(g.kh == 1 && g.kw == 1 && g.stride == 1 && g.pad == 0);
      } else if (l.kind == HF_LAYER_AVGPOOL) {
        cok = cok && l.in == g.cin && l.out == g.cin && !l.has_bias && l.w_off < 0 && l.act == HF_ACT_NONE && i > 0;
        l.s_in = g.hin * g.win, l.s_out = 1;
      } else {
        cok = false;
      }
      // chaining: the feature map this layer reads is the one the previous layer wrote
      if (i > 0) cok = cok && g.cin == net->L[i - 1].out && l.s_in == net->L[i - 1].s_out;
      if (!cok) {
        delete net;
        HF_REQUIRE(false, HF_ERR_INVALID, "hf_net_create: layer %d: inconsistent convolution / pooling geometry", i);
      }
      net->has_conv = true;
    } else if (i > 0 && net->L[i - 1].s_out != 1) {
      delete net;
      HF_REQUIRE(false, HF_ERR_UNSUPPORTED, "hf_net_create: layer %d: a fully connected layer needs one row per sample (pool first)", i);
    }
    if (net->first_trainable < 0 && (l.w_off >= 0 || l.b_off >= 0)) net->first_trainable = i;
    if (l.act == HF_ACT_RELU) net->has_relu = true;
    net->max_width = std::max(net->max_width, std::max(l.in, pad4(l.out)));
    net->max_s = std::max(net->max_s, std::max(l.s_in, l.s_out));
    net->max_act = std::max(net->max_act, std::max((int64_t)l.s_out * pad4(l.out), (int64_t)l.s_out * pad4(l.in)));
    if (i == 0) net->max_act = std::max(net->max_act, (int64_t)l.s_in * pad4(l.kind == HF_LAYER_LINEAR ? l.in : l.geom.cin));
    net->L.push_back(l);
  }
  net->classes = net->L.back().out;
  if (net->L.back().s_out != 1) {
    delete net;
    HF_REQUIRE(false, HF_ERR_UNSUPPORTED, "hf_net_create: the last layer must produce one row per sample (add the average pool)");
  }
  if (net->first_trainable < 0) {
    delete net;
    HF_REQUIRE(false, HF_ERR_INVALID, "hf_net_create: no trainable parameter");
  }
  if (net->L.back().act != HF_ACT_NONE && loss != HF_LOSS_MSE) {
    delete net;
    HF_REQUIRE(false, HF_ERR_UNSUPPORTED, "hf_net_create: softmax-ce / sigmoid-bce expect raw logits from the last layer");
  }
  *out = net;
  return HF_OK;
}

void hf_net_destroy(hf_net_t* net) { delete net; }

int hf_net_set_engine(hf_net_t* net, int32_t engine) {
  HF_REQUIRE(net && (engine == 0 || engine == 1), HF_ERR_INVALID, "hf_net_set_engine: engine must be 0 or 1");
  net->engine = engine;
  return HF_OK;
}

static void release_async(hf_lin* lin) {
  for (cudaEvent_t e : lin->ev)
    if (e) cudaEventDestroy(e);
  lin->ev.clear();
  if (lin->join) cudaEventDestroy(lin->join), lin->join = nullptr;
  if (lin->side) cudaStreamDestroy(lin->side), lin->side = nullptr;
}

// the largest split count any engine's plan may ask for on a weight-gradient contraction (sizes the partial-tile scratch)
static int max_weight_splits(const hf_net* net, int out, int in, int64_t N, bool images) {
  int splits = plan_split(out, in, N, false).splits;
  if (weight_on_tensor(net, out, in, N, 0)) splits = std::max(splits, plan_split(out, in, N, true).splits);
  if (images) splits = std::max(splits, plan_split_pair(out, in, N, nullptr).splits);
  return splits;
}

// Does this net at this chunk size have a contraction the CTA-pair engine wins?  Then the linearisation keeps the
// split-precision images of its operands (twice the activation memory) and the epilogues emit them.
static bool wants_images(const hf_net* net, int64_t N, int flags) {
  if (net->engine != 1 || tc2_mode() == 0 || (flags & HF_LIN_LOSS_ONLY)) return false;
  if (tc2_mode() == 2) return true;
  for (int l = net->first_trainable; l < (int)net->L.size(); ++l) {
    const Layer& L = net->L[l];
    if (L.kind == HF_LAYER_AVGPOOL) continue;
    const int64_t rows = N * L.s_out;
    if (rows * L.out * L.in < kTcMinWork || L.out < 96 || rows < 192) continue;
    const Tc2Choice c = tc2_estimate((int)rows, L.out, L.in, 1, 1);
    if (c.us_pair < 0.8 * c.us_single) return true;
  }
  return false;
}

// one pass over the carve-up: either measures (base == nullptr) or assigns pointers
static size_t carve(const hf_net* net, int64_t N, int flags, char* base, hf_lin* lin) {
  size_t off = 0;
  auto take = [&](size_t bytes) -> char* {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const int nl = (int)net->L.size();
  const bool hess = flags & HF_LIN_HESSIAN, loss_only = flags & HF_LIN_LOSS_ONLY;
  const bool images = wants_images(net, N, flags);
  if (lin) lin->a.assign(nl, nullptr), lin->delta.assign(nl, nullptr), lin->ga.assign(nl, nullptr),
      lin->ra.assign(nl, nullptr), lin->rz.assign(nl, nullptr);
  // conv layers: the largest unfolded input (floats per sample), the position-major copy of the inputs
  int64_t max_unfold = 0, max_unfold16 = 0;
  for (int l = 0; l < nl; ++l)
    if (net->L[l].unfold) {
      max_unfold = std::max(max_unfold, (int64_t)net->L[l].s_out * pad4(net->L[l].in));
      max_unfold16 = std::max(max_unfold16, (int64_t)net->L[l].s_out * pad8(net->L[l].in));
    }
  float* xn = net->L[0].kind != HF_LAYER_LINEAR ? (float*)take(sizeof(float) * N * net->L[0].s_in * pad4(net->L[0].geom.cin)) : nullptr;
  float* ru = max_unfold ? (float*)take(sizeof(float) * N * max_unfold) : nullptr;
  float* du = (max_unfold && !loss_only) ? (float*)take(sizeof(float) * N * max_unfold) : nullptr;
  if (lin) lin->xn = xn, lin->ru = ru, lin->du = du, lin->U.assign(nl, nullptr), lin->RU.assign(nl, nullptr);
  if (loss_only) {
    // activations are not kept: alternate between two buffers; unfolded inputs share one scratch matrix
    float* b0 = (float*)take(sizeof(float) * N * net->max_act);
    float* b1 = (float*)take(sizeof(float) * N * net->max_act);
    if (lin)
      for (int l = 0; l < nl; ++l) lin->a[l] = (l & 1) ? b1 : b0, lin->U[l] = net->L[l].unfold ? ru : nullptr;
  } else {
    for (int l = 0; l < nl; ++l) {
      float* p = (float*)take(sizeof(float) * N * net->L[l].s_out * pad4(net->L[l].out));
      float* u = net->L[l].unfold ? (float*)take(sizeof(float) * N * net->L[l].s_out * pad4(net->L[l].in)) : nullptr;
      if (lin) lin->a[l] = p, lin->U[l] = u;
    }
  }
  float* prob = nullptr;
  float* dL = nullptr;
  if (!loss_only) {
    if (net->loss != HF_LOSS_MSE) prob = (float*)take(sizeof(float) * N * pad4(net->classes));
    dL = (float*)take(sizeof(float) * N * pad4(net->classes));
  }
  float* b0 = loss_only ? nullptr : (float*)take(sizeof(float) * N * net->max_act);
  float* b1 = loss_only ? nullptr : (float*)take(sizeof(float) * N * net->max_act);
  // split-K scratch: the largest weight-gradient partial set, the widest column sum
  size_t pf = 0;
  if (!loss_only)
    for (int l = net->first_trainable; l < nl; ++l) {
      const Layer& L = net->L[l];
      if (L.w_off >= 0) pf = std::max(pf, (size_t)max_weight_splits(net, L.out, L.in, N * L.s_out, images) * L.out * L.in);
      pf = std::max(pf, (size_t)colsum_plan(N * L.s_out) * L.out);
    }
  float* part = pf ? (float*)take(sizeof(float) * pf) : nullptr;
  size_t pf_main = 0;
  if (!loss_only && net->L[net->first_trainable].w_off >= 0 && net->first_trainable < nl - 1) {
    const Layer& L = net->L[net->first_trainable];
    pf_main = (size_t)max_weight_splits(net, L.out, L.in, N * L.s_out, images) * L.out * L.in;
  }
  float* part_main = pf_main ? (float*)take(sizeof(float) * pf_main) : nullptr;
  const int fwd_max = (!loss_only && net->classes <= 32 && !net->has_conv) ? 16 : 0;
  float* pfwd = fwd_max ? (float*)take(sizeof(float) * fwd_max * N * pad4(net->classes)) : nullptr;
  if (lin) lin->partial_fwd = pfwd, lin->fwd_splits_max = fwd_max, lin->fwd_splits = 0, lin->fwd_bias = nullptr;
  if (lin) lin->wpad.assign(nl, nullptr), lin->vpad.assign(nl, nullptr);
  for (int l = 0; l < nl; ++l) {
    const Layer& L = net->L[l];
    if (L.kind == HF_LAYER_AVGPOOL) continue;
    // Unfolded convolutions always contract with a re-ordered (tap-major) copy of their weight, in every engine and
    // also on the loss-only path.  Otherwise a copy is only needed by the tensor engines, when TMA cannot address the
    // flat slice in place: that needs a 16-byte row pitch AND a 16-byte aligned slice start (the flat offset of a
    // layer depends on the sizes of all layers before it: after a 30- or 250-wide layer every later slice is misaligned)
    bool need = L.unfold;
    if (!need && !loss_only && net->engine == 1 && l >= net->first_trainable) {
      const bool aligned = L.w_off >= 0 ? L.w_off % 4 == 0 : (reinterpret_cast<uintptr_t>(L.w_frozen) & 15u) == 0;
      need = !(L.in % 4 == 0 && aligned);
    }
    if (!need) continue;
    float* wp = (float*)take(sizeof(float) * L.out * pad4(L.in));
    float* vp = (!loss_only && L.w_off >= 0 && l >= net->first_trainable) ? (float*)take(sizeof(float) * L.out * pad4(L.in)) : nullptr;
    if (lin) lin->wpad[l] = wp, lin->vpad[l] = vp;
  }
  // fused output head for GGN products: narrow trainable last layer without activation on top of a trainable stack
  bool head_ok = false;
  HeadPlan hp = {};
  if (!loss_only && nl >= 2 && net->first_trainable < nl - 1 && !net->has_conv) {
    const Layer& Lh = net->L[nl - 1];
    head_ok = Lh.act == HF_ACT_NONE && Lh.w_off >= 0 && head_shape_ok(N, Lh.in, Lh.out, sm_count());
    if (head_ok) hp = head_plan(N, Lh.in, Lh.out, sm_count());
  }
  float* hpw = head_ok ? (float*)take(sizeof(float) * (size_t)hp.ctas * net->L[nl - 1].out * net->L[nl - 1].in) : nullptr;
  float* hpb = head_ok ? (float*)take(sizeof(float) * (size_t)hp.ctas * net->L[nl - 1].out) : nullptr;
  if (lin) lin->head_ok = head_ok, lin->head = hp, lin->head_partW = hpw, lin->head_partB = hpb;
  const int64_t max_rows = N * net->max_s;
  const size_t col_rows = (size_t)std::max<int64_t>(std::max<int64_t>(colsum_plan(max_rows), (max_rows + 127) / 128), head_ok ? hp.ctas : 0);
  if (lin) lin->cot.assign(nl, nullptr), lin->colbuf.assign(nl, nullptr), lin->col_rows = col_rows;
  if (!loss_only)
    for (int l = net->first_trainable; l < nl; ++l) {
      float* cb = (float*)take(sizeof(float) * col_rows * net->L[l].out);
      float* ct = l < nl - 1 ? (float*)take(sizeof(float) * N * net->L[l].s_out * pad4(net->L[l].out)) : nullptr;
      if (lin) lin->colbuf[l] = cb, lin->cot[l] = ct;
    }
  float* sq0 = nullptr;
  float* sq1 = nullptr;
  const bool have_sq = !loss_only && net->engine == 1 && !net->has_conv;  // (the Fisher diagonal of a convolution is not a single contraction)
  if (have_sq) {
    sq0 = (float*)take(sizeof(float) * N * pad4(net->max_width));
    sq1 = (float*)take(sizeof(float) * N * pad4(net->max_width));
  }
  if (lin) lin->sq[0] = sq0, lin->sq[1] = sq1, lin->sq_img[0] = lin->sq_img[1] = {nullptr, nullptr, 0};
  // split-precision images: two BF16 planes per matrix, row pitch pad8(width)
  if (lin) lin->use_images = images, lin->imgs.clear(), lin->x_img = {nullptr, nullptr, 0},
           lin->w_img.assign(nl, {nullptr, nullptr, 0}), lin->v_img.assign(nl, {nullptr, nullptr, 0});
  if (images) {
    auto take_img = [&](const float* fp32, int64_t rows, int width) {
      const size_t plane = align_up((size_t)rows * pad8(width) * 2, 256);
      uint16_t* hi = (uint16_t*)take(2 * plane);
      return hf_lin::ImgBuf{fp32, hi, (int64_t)(plane / 2)};
    };
    if (net->L[0].kind == HF_LAYER_LINEAR) {
      const hf_lin::ImgBuf xi = take_img(nullptr, N, net->L[0].in);  // base = the caller's inputs, known at forward time
      if (lin) lin->x_img = xi;
    } else {
      const hf_lin::ImgBuf xi = take_img(xn, N * net->L[0].s_in, net->L[0].geom.cin);
      if (lin) lin->imgs.push_back(xi);
    }
    int64_t max_act16 = 0;  // largest image plane of any ping-pong matrix, in elements per sample
    for (int l = 0; l < nl; ++l) max_act16 = std::max(max_act16, (int64_t)net->L[l].s_out * pad8(net->L[l].out));
    for (int l = 0; l < nl - 1; ++l) {
      const hf_lin::ImgBuf b = take_img(lin ? lin->a[l] : nullptr, N * net->L[l].s_out, net->L[l].out);
      if (lin) lin->imgs.push_back(b);
    }
    for (int l = 0; l < nl; ++l)
      if (net->L[l].unfold) {
        const hf_lin::ImgBuf b = take_img(lin ? lin->U[l] : nullptr, N * net->L[l].s_out, net->L[l].in);
        if (lin) lin->imgs.push_back(b);
      }
    if (max_unfold) {  // (decided by sizes, never by pointers: the measuring pass of this function has none)
      const hf_lin::ImgBuf b = take_img(ru, N, (int)max_unfold16);
      if (lin) lin->imgs.push_back(b);
    }
    for (int i = 0; i < 2; ++i) {
      const hf_lin::ImgBuf b = take_img(i ? b1 : b0, N, (int)max_act16);
      if (lin) lin->imgs.push_back(b);
    }
    const hf_lin::ImgBuf dl = take_img(dL, N, net->classes);
    if (lin) lin->imgs.push_back(dl);
    for (int i = 0; i < 2 && have_sq; ++i) {
      const hf_lin::ImgBuf b = take_img(i ? sq1 : sq0, N, net->max_width);
      if (lin) lin->sq_img[i] = b;
    }
    for (int l = net->first_trainable; l < nl; ++l) {
      const Layer& L = net->L[l];
      if (l < nl - 1) {
        const hf_lin::ImgBuf b = take_img(lin ? lin->cot[l] : nullptr, N * L.s_out, L.out);
        if (lin) lin->imgs.push_back(b);
      }
      if (L.kind == HF_LAYER_AVGPOOL) continue;
      const hf_lin::ImgBuf wi = take_img(nullptr, L.out, L.in);
      hf_lin::ImgBuf vi = {nullptr, nullptr, 0};
      if (L.w_off >= 0) vi = take_img(nullptr, L.out, L.in);
      if (lin) lin->w_img[l] = wi, lin->v_img[l] = vi;
    }
  }
  int64_t lb = (N + 7) / 8;
  if (lb > 1024) lb = 1024;
  double* lp = (double*)take(sizeof(double) * lb);
  if (hess && !loss_only) {
    for (int l = net->first_trainable; l < nl; ++l) {
      const size_t bytes = sizeof(float) * N * net->L[l].s_out * pad4(net->L[l].out);
      if (net->L[l].unfold && l > net->first_trainable) {
        float* ruk = (float*)take(sizeof(float) * N * net->L[l].s_out * pad4(net->L[l].in));
        if (lin) lin->RU[l] = ruk;
      }
      float* d = (l < nl - 1) ? (float*)take(bytes) : nullptr;  // the last layer's delta is deltaL
      float* r = (l < nl - 1) ? (float*)take(bytes) : nullptr;
      float* g = nullptr;
      float* z = nullptr;
      if (l < nl - 1 && curved(net->L[l].act)) g = (float*)take(bytes), z = (float*)take(bytes);
      if (lin) lin->delta[l] = d ? d : dL, lin->ra[l] = r, lin->ga[l] = g, lin->rz[l] = z;
    }
  }
  if (lin) {
    lin->prob = prob, lin->deltaL = dL, lin->buf[0] = b0, lin->buf[1] = b1;
    lin->partial = part, lin->partial_floats = pf, lin->partial_main = part_main, lin->partial_main_floats = pf_main;
    lin->loss_partial = lp, lin->loss_blocks = (int)lb;
  }
  return off;
}

size_t hf_lin_workspace_bytes(const hf_net_t* net, int64_t batch, int32_t flags) {
  if (!net || batch <= 0) return 0;
  return carve(net, batch, flags, nullptr, nullptr);
}

int hf_lin_create(const hf_net_t* net, int64_t batch, int32_t flags, void* d_workspace, size_t workspace_bytes,
                  hf_lin_t** out) {
  HF_REQUIRE(net && out && batch > 0 && d_workspace, HF_ERR_INVALID, "hf_lin_create: bad arguments");
  HF_REQUIRE(batch < (1ll << 31), HF_ERR_UNSUPPORTED, "hf_lin_create: chunk of %lld samples is too large", (long long)batch);
  HF_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255u) == 0, HF_ERR_WORKSPACE, "hf_lin_create: workspace must be 256-byte aligned");
  if ((flags & HF_LIN_HESSIAN) && net->L.back().act != HF_ACT_NONE)
    HF_REQUIRE(false, HF_ERR_UNSUPPORTED, "Hessian products with an activation after the last layer are not supported");
  const size_t need = carve(net, batch, flags, nullptr, nullptr);
  HF_REQUIRE(workspace_bytes >= need, HF_ERR_WORKSPACE, "hf_lin_create: workspace has %zu bytes, need %zu", workspace_bytes, need);
  hf_lin* lin = new (std::nothrow) hf_lin();
  HF_REQUIRE(lin, HF_ERR_INVALID, "hf_lin_create: out of host memory");
  lin->net = net, lin->N = batch, lin->flags = flags, lin->x = nullptr, lin->n_total = batch;
  lin->have_forward = lin->have_gradient = false, lin->loss_fused = false;
  carve(net, batch, flags, static_cast<char*>(d_workspace), lin);
  lin->side = nullptr, lin->join = nullptr, lin->pending_cur = nullptr, lin->pending_cols = 0;
  if (!(flags & HF_LIN_LOSS_ONLY) && !getenv("HF_SINGLE_STREAM")) {
    // side stream + events for the forked gradient work; failure to create them only disables the overlap
    const int nl = (int)net->L.size();
    bool ok = cudaStreamCreateWithFlags(&lin->side, cudaStreamNonBlocking) == cudaSuccess;
    lin->ev.assign(nl, nullptr);
    for (int l = 0; ok && l < nl; ++l) ok = cudaEventCreateWithFlags(&lin->ev[l], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&lin->join, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      release_async(lin);
      (void)cudaGetLastError();
    }
  }
  *out = lin;
  return HF_OK;
}

void hf_lin_destroy(hf_lin_t* lin) {
  if (!lin) return;
  release_async(lin);
  delete lin;
}

const float* hf_lin_logits(const hf_lin_t* lin) { return lin ? lin->a.back() : nullptr; }

int hf_lin_forward(hf_lin_t* lin, const float* d_theta, const float* d_x, const void* d_targets, int64_t n_total,
                   double* d_loss_acc, void* stream_) {
  HF_REQUIRE(lin && d_x && d_targets && n_total > 0, HF_ERR_INVALID, "hf_lin_forward: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  const hf_net* net = lin->net;
  HF_REQUIRE(d_theta || net->P == 0, HF_ERR_INVALID, "hf_lin_forward: theta is null");
  const int nl = (int)net->L.size();
  lin->x = d_x, lin->n_total = n_total;
  if (net->L[0].kind != HF_LAYER_LINEAR) {  // inputs arrive NCHW: position-major copy (+ image) for the tile engines
    const Layer& L0 = net->L[0];
    nchw_to_nhwc_kernel<<<conv_blocks(lin->N * L0.s_in * (int64_t)L0.geom.cin), 256, 0, stream>>>(
        d_x, lin->xn, lin->N, L0.geom.cin, L0.s_in, pad4(L0.geom.cin), image_for(lin, lin->xn, L0.geom.cin));
    HF_LAUNCH_CHECK();
  }
  const Image16 none = {nullptr, 0, 0};
  {
    // Operand forms of the weights (constant for the life of this linearisation), in one launch: the tap-major copies
    // unfolded convolutions contract with, the 16-byte-pitched FP32 copies of slices TMA cannot address in place, and
    // (pair engine) the split-precision images.
    SplitTable t;
    t.count = 0, t.skip = nullptr;
    int rc = HF_OK;
    for (int l = 0; l < nl && !rc; ++l) {
      const Layer& L = net->L[l];
      if (L.kind == HF_LAYER_AVGPOOL) continue;
      const bool img = lin->use_images && lin->w_img[l].hi;
      if (!lin->wpad[l] && !img) continue;
      if (img) lin->w_img[l].base = lin->wpad[l] ? lin->wpad[l] : weight_ptr(L, d_theta);
      t.seg[t.count++] = SplitSegment{weight_ptr(L, d_theta), L.out, L.in, L.in, lin->wpad[l], pad4(L.in),
                                      img ? Image16{lin->w_img[l].hi, lin->w_img[l].plane, pad8(L.in)} : none, 0,
                                      L.geom.cin, L.unfold ? L.geom.kh * L.geom.kw : 0};
      if (t.count == kMaxSplitSegments) rc = launch_split(t, stream), t.count = 0;
    }
    if (!rc) rc = launch_split(t, stream);
    if (rc) return rc;
  }
  for (int l = 0; l < nl; ++l) {
    const Layer& L = net->L[l];
    if (L.kind == HF_LAYER_AVGPOOL) {
      avgpool_kernel<<<conv_blocks(lin->N * (int64_t)L.out), 256, 0, stream>>>(lin->a[l - 1], pad4(L.out), lin->a[l], lin->N, L.s_in, L.out,
                                                                                Image16{nullptr, 0, 0}, nullptr);
      HF_LAUNCH_CHECK();
      continue;
    }
    if (L.unfold) {  // the unfolded input (+ its image) stays resident: every product of the solve contracts with it
      const float* src = l == 0 ? lin->xn : lin->a[l - 1];
      im2col_kernel<<<conv_blocks(rows_out(lin, l) * (pad4(L.in) / 4)), 256, 0, stream>>>(src, pad4(L.geom.cin), lin->U[l], pad4(L.in), lin->N,
                                                                                         L.geom, image_for(lin, lin->U[l], L.in), nullptr);
      HF_LAUNCH_CHECK();
    }
    int ld_in = 0;
    const float* a_in = layer_input(lin, l, &ld_in);
    GemmArgs g = blank_gemm();
    g.M = (int)rows_out(lin, l), g.N = L.out, g.K = L.in, g.n_pairs = 1;
    g.A[0] = op_kc(a_in, ld_in);
    g.B[0] = L.unfold ? op_kc(lin->wpad[l], pad4(L.in)) : op_kc(weight_ptr(L, d_theta), L.in);  // (tap-major copy for unfolded inputs)
    g.C = lin->a[l], g.ldc = pad4(L.out);
    g.epi = EPI_BIAS_ACT, g.act = L.act, g.bias = bias_ptr(L, d_theta);
    // The linearisation point fixes the ReLU masks for the whole solve, so it is evaluated in plain FP32: the
    // ~5e-6 error of the split-precision tensor tiles would flip a few dozen of the 4M masks of the MLP config (each flip moves a
    // gradient row by ~1/sqrt(N)).  Loss-only evaluations (line search, backtracking) have no masks to fix and
    // stay on the tensor-core tiles.
    // Nets without ReLU have no masks to fix: their linearisation point tolerates the ~5e-6 of the tensor tiles, which
    // only moves it as a slightly different parameter vector would.
    const bool exact = !(lin->flags & HF_LIN_LOSS_ONLY) && net->has_relu;
    int rc = exact ? launch_gemm_simt(g, stream) : run_gemm(net, g, stream);
    if (rc) return rc;
  }
  {
    // (pair engine) the split-precision images of the inputs and the activations, constant for the life of this
    // linearisation, in one launch
    SplitTable t;
    t.count = 0, t.skip = nullptr;
    auto push = [&](const SplitSegment& sg) -> int {
      t.seg[t.count++] = sg;
      if (t.count < kMaxSplitSegments) return HF_OK;
      int rc = launch_split(t, stream);
      t.count = 0;
      return rc;
    };
    int rc = HF_OK;
    if (lin->use_images) {
      if (net->L[0].kind == HF_LAYER_LINEAR) {
        lin->x_img.base = d_x;
        rc = push(SplitSegment{d_x, lin->N, net->L[0].in, net->L[0].in, nullptr, 0, image_for(lin, d_x, net->L[0].in)});
      }
      for (int l = 0; l < nl - 1 && !rc; ++l)
        rc = push(SplitSegment{lin->a[l], rows_out(lin, l), net->L[l].out, pad4(net->L[l].out), nullptr, 0,
                               image_for(lin, lin->a[l], net->L[l].out)});
    }
    if (!rc) rc = launch_split(t, stream);
    if (rc) return rc;
  }
  LossArgs a;
  a.out = lin->a.back(), a.target = d_targets, a.N = lin->N, a.C = net->classes, a.ld = pad4(net->classes);
  a.loss = net->loss, a.final_act = net->L.back().act;
  a.scale = loss_scale(net, n_total);
  a.prob = lin->prob, a.delta = lin->deltaL, a.partial = lin->loss_partial;
  a.delta_img = image_for(lin, lin->deltaL, net->classes);
  int64_t blocks = (lin->N + 7) / 8;
  if (blocks > lin->loss_blocks) blocks = lin->loss_blocks;
  loss_forward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  HF_LAUNCH_CHECK();
  if (d_loss_acc) {
    loss_finalize_kernel<<<1, 256, 0, stream>>>(lin->loss_partial, (int)blocks, (double)a.scale, d_loss_acc);
    HF_LAUNCH_CHECK();
  }
  lin->have_forward = true, lin->have_gradient = false;
  return HF_OK;
}

static int check_ready(const hf_lin* lin, const char* who, bool need_grad) {
  HF_REQUIRE(lin, HF_ERR_INVALID, "%s: null linearisation", who);
  HF_REQUIRE(!(lin->flags & HF_LIN_LOSS_ONLY), HF_ERR_INVALID, "%s: linearisation was created loss-only", who);
  HF_REQUIRE(lin->have_forward, HF_ERR_INVALID, "%s: call hf_lin_forward first", who);
  HF_REQUIRE(!need_grad || lin->have_gradient, HF_ERR_INVALID, "%s: call hf_lin_gradient first", who);
  return HF_OK;
}

int hf_lin_gradient(hf_lin_t* lin, const float* d_theta, float* d_grad, int32_t accumulate, void* stream) {
  int rc = check_ready(lin, "hf_lin_gradient", false);
  if (rc) return rc;
  HF_REQUIRE(d_grad, HF_ERR_INVALID, "hf_lin_gradient: null output");
  rc = backward_sweep(lin, d_theta, nullptr, lin->deltaL, d_grad, accumulate, BACK_GRADIENT, nullptr,
                      (cudaStream_t)stream);
  if (rc == HF_OK) lin->have_gradient = true;
  return rc;
}

int hf_ggn_matvec(hf_lin_t* lin, const float* d_theta, const float* d_v, float* d_out, int32_t accumulate,
                  const int32_t* d_skip, void* stream_) {
  int rc = check_ready(lin, "hf_ggn_matvec", false);
  if (rc) return rc;
  HF_REQUIRE(d_v && d_out, HF_ERR_INVALID, "hf_ggn_matvec: null vector");
  cudaStream_t stream = (cudaStream_t)stream_;
  int top = 0;
  if (lin->head_ok) {
    const float* ra = nullptr;
    rc = rop_forward(lin, d_theta, d_v, false, d_skip, stream, &top, &ra);
    if (rc) return rc;
    rc = ggn_head(lin, d_theta, d_v, ra, d_skip, stream);
    if (rc) return rc;
    return backward_sweep(lin, d_theta, d_v, nullptr, d_out, accumulate, BACK_GGN, d_skip, stream, -1, true);
  }
  rc = rop_forward(lin, d_theta, d_v, false, d_skip, stream, &top);
  if (rc) return rc;
  rc = apply_loss_hessian(lin, lin->buf[top], d_skip, stream);
  if (rc) return rc;
  return backward_sweep(lin, d_theta, d_v, lin->buf[top], d_out, accumulate, BACK_GGN, d_skip, stream);
}

int hf_matvec_phase(hf_lin_t* lin, int32_t kind, const float* d_theta, const float* d_v, float* d_out, int32_t accumulate,
                    const int32_t* d_skip, void* stream_, int32_t phase) {
  int rc = check_ready(lin, "hf_matvec_phase", kind == 1);
  if (rc) return rc;
  HF_REQUIRE((kind == 0 || kind == 1) && (phase == 0 || phase == 1), HF_ERR_INVALID, "hf_matvec_phase: bad kind/phase");
  HF_REQUIRE(kind == 0 || (lin->flags & HF_LIN_HESSIAN), HF_ERR_INVALID, "hf_matvec_phase: linearisation lacks HF_LIN_HESSIAN");
  HF_REQUIRE(d_v && d_out, HF_ERR_INVALID, "hf_matvec_phase: null vector");
  cudaStream_t stream = (cudaStream_t)stream_;
  const BackMode mode = kind == 1 ? BACK_HESSIAN : BACK_GGN;
  if (phase == 0) {
    int top = 0;
    if (kind == 0 && lin->head_ok) {
      const float* ra = nullptr;
      rc = rop_forward(lin, d_theta, d_v, false, d_skip, stream, &top, &ra);
      if (rc) return rc;
      rc = ggn_head(lin, d_theta, d_v, ra, d_skip, stream);
      if (rc) return rc;
      return backward_sweep(lin, d_theta, d_v, nullptr, d_out, accumulate, mode, d_skip, stream, 0, true);
    }
    rc = rop_forward(lin, d_theta, d_v, kind == 1, d_skip, stream, &top);
    if (rc) return rc;
    rc = apply_loss_hessian(lin, lin->buf[top], d_skip, stream);
    if (rc) return rc;
    return backward_sweep(lin, d_theta, d_v, lin->buf[top], d_out, accumulate, mode, d_skip, stream, 0);
  }
  HF_REQUIRE(lin->pending_cur, HF_ERR_INVALID, "hf_matvec_phase: phase 1 without phase 0");
  return backward_sweep(lin, d_theta, d_v, nullptr, d_out, accumulate, mode, d_skip, stream, 1);
}

int hf_net_first_layer_span(const hf_net_t* net, int64_t* offset, int64_t* count) {
  HF_REQUIRE(net && offset && count, HF_ERR_INVALID, "hf_net_first_layer_span: null argument");
  const Layer& L = net->L[net->first_trainable];
  int64_t lo = INT64_MAX, hi = 0;
  if (L.w_off >= 0) lo = std::min(lo, L.w_off), hi = std::max(hi, L.w_off + (int64_t)L.in * L.out);
  if (L.has_bias && L.b_off >= 0) lo = std::min(lo, L.b_off), hi = std::max(hi, L.b_off + (int64_t)L.out);
  int64_t sum = (L.w_off >= 0 ? (int64_t)L.in * L.out : 0) + ((L.has_bias && L.b_off >= 0) ? L.out : 0);
  *offset = lo, *count = (hi - lo == sum) ? sum : 0;  // 0: the layer's slices are not adjacent in the flat vector
  return HF_OK;
}

int hf_hessian_matvec(hf_lin_t* lin, const float* d_theta, const float* d_v, float* d_out, int32_t accumulate,
                      const int32_t* d_skip, void* stream_) {
  int rc = check_ready(lin, "hf_hessian_matvec", true);
  if (rc) return rc;
  HF_REQUIRE(lin->flags & HF_LIN_HESSIAN, HF_ERR_INVALID, "hf_hessian_matvec: linearisation lacks HF_LIN_HESSIAN");
  HF_REQUIRE(d_v && d_out, HF_ERR_INVALID, "hf_hessian_matvec: null vector");
  cudaStream_t stream = (cudaStream_t)stream_;
  int top = 0;
  rc = rop_forward(lin, d_theta, d_v, true, d_skip, stream, &top);
  if (rc) return rc;
  rc = apply_loss_hessian(lin, lin->buf[top], d_skip, stream);
  if (rc) return rc;
  return backward_sweep(lin, d_theta, d_v, lin->buf[top], d_out, accumulate, BACK_HESSIAN, d_skip, stream);
}

int hf_fisher_diag(hf_lin_t* lin, const float* d_theta, float* d_out, int32_t accumulate, void* stream) {
  int rc = check_ready(lin, "hf_fisher_diag", false);
  if (rc) return rc;
  // (for a convolution the square sits outside the sum over positions: (d^2)^T (a^2) is NOT the diagonal, SURVEY.md
  // section 7, hard part 7; those layers go through conv_fisher_gradient)
  HF_REQUIRE(d_out, HF_ERR_INVALID, "hf_fisher_diag: null output");
  return backward_sweep(lin, d_theta, nullptr, lin->deltaL, d_out, accumulate, BACK_FISHER, nullptr,
                        (cudaStream_t)stream);
}

int hf_debug_tc_trace(void* d_buf) { return set_tc_trace(d_buf); }
int hf_debug_tc2_trace(void* d_buf) { return set_tc2_trace(d_buf); }
int hf_debug_tc_trace_iters(void* d_buf) { return set_tc_trace_iters(d_buf); }

size_t hf_contract_workspace_bytes(int64_t M, int64_t N, int64_t K, int32_t n_pairs) {
  // engine 2: two BF16 planes per operand; either orientation of [MN, K] fits in max(rows * pad8(cols))
  auto planes = [&](int64_t mn) {
    return 2 * std::max(align_up((size_t)mn * pad8((int)K) * 2, 256), align_up((size_t)K * pad8((int)mn) * 2, 256));
  };
  return (size_t)n_pairs * (planes(M) + planes(N));
}

int hf_contract(int32_t engine, int64_t M, int64_t N, int64_t K, int32_t n_pairs, const hf_operand* A,
                const hf_operand* B, float* d_C, int64_t ldc, void* d_workspace, size_t workspace_bytes, void* stream) {
  HF_REQUIRE(A && B && d_C && n_pairs >= 1 && n_pairs <= 2, HF_ERR_INVALID, "hf_contract: bad arguments");
  HF_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), HF_ERR_INVALID, "hf_contract: bad shape");
  GemmArgs g = blank_gemm();
  g.M = (int)M, g.N = (int)N, g.K = (int)K, g.n_pairs = n_pairs;
  for (int s = 0; s < n_pairs; ++s) {
    g.A[s] = Operand{A[s].d_ptr, A[s].stride_mn, A[s].stride_k};
    g.B[s] = Operand{B[s].d_ptr, B[s].stride_mn, B[s].stride_k};
  }
  g.C = d_C, g.ldc = ldc, g.epi = EPI_STORE;
  if (engine == 2 || engine == 3) {  // 3 (profiling): the images are still in the workspace from an identical engine-2 call
    // pre-split pair engine: the operand images are built in the caller's workspace first (the products keep them
    // resident instead: hf_lin_forward / the producing epilogues write them)
    SplitTable t;
    t.count = 0, t.skip = nullptr;
    size_t off = 0;
    char* base = static_cast<char*>(d_workspace);
    for (int s = 0; s < n_pairs; ++s)
      for (int which = 0; which < 2; ++which) {
        Operand& op = which ? g.B[s] : g.A[s];
        const int64_t mn = which ? N : M;
        const bool kc = op.s_k == 1;
        const int64_t rows = kc ? mn : K, cols = kc ? K : mn, ld = kc ? op.s_mn : op.s_k;
        const int64_t ld16 = pad8((int)cols);
        const size_t plane_bytes = align_up((size_t)rows * ld16 * 2, 256);
        HF_REQUIRE(base && off + 2 * plane_bytes <= workspace_bytes, HF_ERR_WORKSPACE,
                   "hf_contract: engine 2 needs a workspace of at least %zu bytes for the operand images", hf_contract_workspace_bytes(M, N, K, n_pairs));
        op.img.hi = reinterpret_cast<uint16_t*>(base + off), op.img.plane = (int64_t)(plane_bytes / 2), op.img.ld = ld16;
        off += 2 * plane_bytes;
        t.seg[t.count++] = SplitSegment{op.ptr, rows, (int)cols, ld, nullptr, 0, op.img};
      }
    int rc = engine == 2 ? launch_split(t, (cudaStream_t)stream) : HF_OK;
    if (rc) return rc;
    HF_REQUIRE(tc2_supported(g), HF_ERR_UNSUPPORTED, "hf_contract: shape/alignment not supported by the pre-split tcgen05 engine");
    return launch_gemm_tc2(g, (cudaStream_t)stream);
  }
  if (engine == 1) {
    HF_REQUIRE(tc_supported(g), HF_ERR_UNSUPPORTED, "hf_contract: shape/alignment not supported by the tcgen05 engine");
    return launch_gemm_tc(g, (cudaStream_t)stream);
  }
  HF_REQUIRE(engine == 0, HF_ERR_INVALID, "hf_contract: unknown engine %d (0 = SIMT, 1 = tcgen05 128x128, 2 = tcgen05 pre-split pairs)", engine);
  return launch_gemm_simt(g, (cudaStream_t)stream);
}

}  // extern "C"
