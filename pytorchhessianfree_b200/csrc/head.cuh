// Fused output head of a GGN product for a narrow last layer (C <= 32 outputs, e.g. 10 classes).
//
// For the last affine layer z = W a + b (no activation) with the loss on top, one pass over the rows does what is
// otherwise five launches (R-op GEMM with a 10-wide output, loss-Hessian kernel, transposed data GEMM with K = 10,
// weight-gradient GEMM with M = 10, bias column sums) -- each of which costs a full tensor-tile kernel's fixed
// overhead for a few MFLOP:
//     Rz[n,:]  = W Ra[n,:] + V a[n,:] + vb                      (R-op, reference: backpack rop / optimizer.py:457-462)
//     u[n,:]   = H_loss(n) Rz[n,:]                              (mse 2s, softmax s(diag p - p p^T), bce s p(1-p))
//     cot[n,:] = (W^T u[n,:]) * act'(a[n,:])                    (cotangent handed to the layer below)
//     G_W     += u[n,:] a[n,:]^T,  G_b += u[n,:],  colsum += cot[n,:]   (per-CTA partials, reduced in fixed order)
// The work is memory bound (a, Ra in, cot out: 12 D bytes per row); W and V live in shared memory, each warp carries
// two rows at a time and each lane four columns (128-bit shared and global accesses), and all sums have a fixed order
// (deterministic).
#pragma once
#include "common.cuh"

namespace hf {

struct HeadArgs {
  const float* a;     // [N, ld]  activations feeding the head layer
  const float* ra;    // [N, ld]  their tangent, or null (zero)
  const float* W;     // [C, D]
  const float* V;     // [C, D]   tangent of W
  const float* vb;    // [C]      tangent of the bias, or null
  const float* prob;  // [N, ldc] softmax / sigmoid probabilities (null for mse)
  float* cot;         // [N, ld]  out
  float* partW;       // [ctas][C*D] out
  float* partB;       // [ctas][C]   out, or null
  float* partCol;     // [ctas][D]   out, or null: column sums of cot (bias gradient of the layer below)
  int64_t N;
  int D, Dp, C, ld, ldc;
  int loss, act_prev;
  float scale;
  int rows_per_cta;   // multiple of 32
  const int32_t* skip;
};

constexpr int kHeadThreads = 512;
constexpr int kHeadWarps = kHeadThreads / 32;
constexpr int kHeadRows = 32;                      // rows per step of a CTA
constexpr int kHeadRQ = kHeadRows / kHeadWarps;    // rows a warp carries at a time

__device__ __forceinline__ float head_act_d1(int act, float s) {
  if (act == HF_ACT_RELU) return s > 0.f ? 1.f : 0.f;
  if (act == HF_ACT_SIGMOID) return s * (1.f - s);
  if (act == HF_ACT_TANH) return 1.f - s * s;
  return 1.f;
}

template <int CP>
__global__ void __launch_bounds__(kHeadThreads, 1) ggn_head_kernel(HeadArgs h) {
  if (h.skip && *h.skip) return;
  constexpr int RQ = kHeadRQ;
  extern __shared__ __align__(16) float head_smem[];
  const int Dp = h.Dp, D = h.D, C = h.C;
  float* Ws = head_smem;                 // [CP][Dp], zero padded
  float* Vs = Ws + CP * Dp;              // [CP][Dp]
  float* as = Vs + CP * Dp;              // [32][Dp]  the 32 rows of `a` of the current step
  float* us = as + kHeadRows * Dp;       // [32][CP]  their u = H_loss Rz
  float* colw = us + kHeadRows * CP;     // [warps][Dp] per-warp column sums of cot
  float* Gs = colw + kHeadWarps * Dp;    // [CP][Dp]  running head gradient (only when a CTA takes more than one step)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // stage W and V (zero padded to [CP][Dp]); four rows of loads are in flight per thread
  if (CP != C || Dp != D) {
    for (int i = tid; i < 2 * CP * Dp; i += kHeadThreads) Ws[i] = 0.f;
    __syncthreads();
  }
  for (int d = tid; d < D; d += kHeadThreads)
    for (int c0 = 0; c0 < C; c0 += 4) {
      float w[4], v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k] = c0 + k < C ? __ldg(h.W + (int64_t)(c0 + k) * D + d) : 0.f;
        v[k] = c0 + k < C ? __ldg(h.V + (int64_t)(c0 + k) * D + d) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c0 + k < C) Ws[(c0 + k) * Dp + d] = w[k], Vs[(c0 + k) * Dp + d] = v[k];
    }
  for (int i = tid; i < kHeadWarps * Dp; i += kHeadThreads) colw[i] = 0.f;
  float vb[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) vb[c] = (h.vb && c < C) ? __ldg(h.vb + c) : 0.f;
  __syncthreads();

  const int64_t n_cta = (int64_t)blockIdx.x * h.rows_per_cta;
  const int rows_cta = (int)min((int64_t)h.rows_per_cta, h.N - n_cta);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float bsum = 0.f;  // thread c < C: running sum of u[:, c]
  for (int s0 = 0; s0 < rows_cta; s0 += kHeadRows) {
    const int64_t n0 = n_cta + s0;
    const int rows = min(kHeadRows, rows_cta - s0);
    const int r0 = warp * RQ;
    // ---- R-op: RQ rows per warp, each lane four consecutive columns per 128-column chunk (128-bit loads); the
    // rows of `a` are parked in shared memory on the way, and the next chunk's global loads are in flight while the
    // current one is multiplied ----
    float acc[RQ][CP];
#pragma unroll
    for (int q = 0; q < RQ; ++q)
#pragma unroll
      for (int c = 0; c < CP; ++c) acc[q][c] = 0.f;
    float4 a4[RQ], r4[RQ], an[RQ], rn[RQ];
    auto fetch = [&](int d, float4 (&ao)[RQ], float4 (&ro)[RQ]) {
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        const bool ok = r0 + q < rows && d < D;  // d is a multiple of 4 and ld = pad4(D), so the quad stays in the row
        const int64_t at = (n0 + r0 + q) * h.ld + d;
        ao[q] = ok ? __ldg(reinterpret_cast<const float4*>(h.a + at)) : z4;
        ro[q] = (ok && h.ra) ? __ldg(reinterpret_cast<const float4*>(h.ra + at)) : z4;
        if (D & 3) {  // the pitch padding of a row is never written by its producer: mask it
          if (d + 1 >= D) ao[q].y = 0.f, ro[q].y = 0.f;
          if (d + 2 >= D) ao[q].z = 0.f, ro[q].z = 0.f;
          if (d + 3 >= D) ao[q].w = 0.f, ro[q].w = 0.f;
        }
      }
    };
    fetch(4 * lane, a4, r4);
#pragma unroll 1
    for (int d0 = 0; d0 < Dp; d0 += 128) {
      const int d = d0 + 4 * lane;
      if (d0 + 128 < Dp) fetch(d + 128, an, rn);
#pragma unroll
      for (int q = 0; q < RQ; ++q) *reinterpret_cast<float4*>(as + (r0 + q) * Dp + d) = a4[q];
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + c * Dp + d);
        const float4 vv = *reinterpret_cast<const float4*>(Vs + c * Dp + d);
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
          float t = acc[q][c];
          t = fmaf(r4[q].x, w.x, t), t = fmaf(r4[q].y, w.y, t), t = fmaf(r4[q].z, w.z, t), t = fmaf(r4[q].w, w.w, t);
          t = fmaf(a4[q].x, vv.x, t), t = fmaf(a4[q].y, vv.y, t), t = fmaf(a4[q].z, vv.z, t), t = fmaf(a4[q].w, vv.w, t);
          acc[q][c] = t;
        }
      }
#pragma unroll
      for (int q = 0; q < RQ; ++q) a4[q] = an[q], r4[q] = rn[q];
    }
#pragma unroll
    for (int q = 0; q < RQ; ++q)
#pragma unroll
      for (int c = 0; c < CP; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q][c] += __shfl_xor_sync(0xffffffffu, acc[q][c], o);
    // ---- loss Hessian, every lane redundantly (the values stay in registers for the transposed product) ----
#pragma unroll
    for (int q = 0; q < RQ; ++q) {
      const bool row_ok = r0 + q < rows;
      const int64_t n = n0 + r0 + q;
      float pr[CP];
#pragma unroll
      for (int c = 0; c < CP; c += 4) {  // ldc = pad4(C): whole quads, 16-byte aligned
        const float4 p4 = (row_ok && c < C && h.prob) ? __ldg(reinterpret_cast<const float4*>(h.prob + n * h.ldc + c)) : z4;
        pr[c] = p4.x, pr[c + 1] = p4.y, pr[c + 2] = p4.z, pr[c + 3] = p4.w;
      }
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float rz = acc[q][c] + vb[c];
        if (c >= C) pr[c] = 0.f;
        if (h.loss == HF_LOSS_SOFTMAX_CE) {
          dot = fmaf(pr[c], rz, dot);
          acc[q][c] = rz;
        } else if (h.loss == HF_LOSS_MSE) {
          acc[q][c] = 2.f * h.scale * rz;
        } else {
          acc[q][c] = h.scale * pr[c] * (1.f - pr[c]) * rz;
        }
      }
      if (h.loss == HF_LOSS_SOFTMAX_CE) {
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[q][c] = h.scale * pr[c] * (acc[q][c] - dot);
      }
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (!row_ok || c >= C) acc[q][c] = 0.f;
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < CP; c += 4)
          *reinterpret_cast<float4*>(us + (r0 + q) * CP + c) = make_float4(acc[q][c], acc[q][c + 1], acc[q][c + 2], acc[q][c + 3]);
      }
    }
    // ---- transposed product: cot = (u W) * act'(a) ----
#pragma unroll 1
    for (int d0 = 0; d0 < Dp; d0 += 128) {
      const int d = d0 + 4 * lane;
      float4 cv[RQ];
#pragma unroll
      for (int q = 0; q < RQ; ++q) cv[q] = z4;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + c * Dp + d);
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
          cv[q].x = fmaf(acc[q][c], w.x, cv[q].x), cv[q].y = fmaf(acc[q][c], w.y, cv[q].y);
          cv[q].z = fmaf(acc[q][c], w.z, cv[q].z), cv[q].w = fmaf(acc[q][c], w.w, cv[q].w);
        }
      }
      float4 csum = *reinterpret_cast<const float4*>(colw + warp * Dp + d);  // this warp's row, this lane's columns
#pragma unroll
      for (int q = 0; q < RQ; ++q) {
        if (r0 + q < rows && d < D) {
          const float4 s4 = *reinterpret_cast<const float4*>(as + (r0 + q) * Dp + d);
          float4 val;
          val.x = cv[q].x * head_act_d1(h.act_prev, s4.x), val.y = cv[q].y * head_act_d1(h.act_prev, s4.y);
          val.z = cv[q].z * head_act_d1(h.act_prev, s4.z), val.w = cv[q].w * head_act_d1(h.act_prev, s4.w);
          *reinterpret_cast<float4*>(h.cot + (n0 + r0 + q) * h.ld + d) = val;
          csum.x += val.x, csum.y += val.y, csum.z += val.z, csum.w += val.w;
        }
      }
      *reinterpret_cast<float4*>(colw + warp * Dp + d) = csum;
    }
    __syncthreads();
    // ---- weight gradient of the head: G[c][d] += sum_r u[r][c] a[r][d], threads over columns ----
    const bool last = s0 + kHeadRows >= rows_cta;
    for (int d = tid; d < D; d += kHeadThreads) {
      float g[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) g[c] = s0 > 0 ? Gs[c * Dp + d] : 0.f;
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float a_rd = as[r * Dp + d];
#pragma unroll
        for (int c = 0; c < CP; c += 4) {
          const float4 u4 = *reinterpret_cast<const float4*>(us + r * CP + c);
          g[c] = fmaf(u4.x, a_rd, g[c]), g[c + 1] = fmaf(u4.y, a_rd, g[c + 1]);
          g[c + 2] = fmaf(u4.z, a_rd, g[c + 2]), g[c + 3] = fmaf(u4.w, a_rd, g[c + 3]);
        }
      }
      if (last) {
#pragma unroll
        for (int c = 0; c < CP; ++c)
          if (c < C) h.partW[((int64_t)blockIdx.x * C + c) * D + d] = g[c];
      } else {
#pragma unroll
        for (int c = 0; c < CP; ++c) Gs[c * Dp + d] = g[c];
      }
    }
    if (tid < C)
      for (int r = 0; r < rows; ++r) bsum += us[r * CP + tid];
    __syncthreads();  // as / us are rewritten by the next step
  }
  if (h.partB && tid < C) h.partB[(int64_t)blockIdx.x * C + tid] = bsum;
  if (h.partCol)
    for (int d = tid; d < D; d += kHeadThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kHeadWarps; ++w) s += colw[w * Dp + d];
      h.partCol[(int64_t)blockIdx.x * D + d] = s;
    }
}

struct HeadPlan {
  int ctas, rows_per_cta, CP, Dp;
  size_t smem;
};

// rows are dealt out in blocks of 32 (16 warps x 2 rows); at most one CTA per SM so the partial sets stay small
inline HeadPlan head_plan(int64_t N, int D, int C, int sms) {
  HeadPlan p;
  const int64_t steps = (N + 32LL * sms - 1) / (32LL * sms);
  p.rows_per_cta = (int)(32 * steps);
  p.ctas = (int)((N + p.rows_per_cta - 1) / p.rows_per_cta);
  p.CP = (C + 3) / 4 * 4;
  p.Dp = (D + 127) / 128 * 128;
  p.smem = sizeof(float) * ((size_t)2 * p.CP * p.Dp + (size_t)kHeadRows * p.Dp + (size_t)kHeadRows * p.CP +
                            (size_t)kHeadWarps * p.Dp + (p.rows_per_cta > kHeadRows ? (size_t)p.CP * p.Dp : 0));
  return p;
}

inline bool head_shape_ok(int64_t N, int D, int C, int sms) {
  if (C < 1 || C > 32 || D < 1 || N < 1) return false;
  const HeadPlan p = head_plan(N, D, C, sms);
  return p.smem <= 200 * 1024;
}

template <int CP>
inline int launch_head_cp(const HeadArgs& h, const HeadPlan& p, cudaStream_t stream) {
  static bool seen[64] = {};
  if (first_use_on_device(seen))  // the opt-in is per device
    HF_CUDA(cudaFuncSetAttribute(ggn_head_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ggn_head_kernel<CP><<<p.ctas, kHeadThreads, p.smem, stream>>>(h);
  HF_LAUNCH_CHECK();
  return HF_OK;
}

inline int launch_head(const HeadArgs& h, const HeadPlan& p, cudaStream_t stream) {
  switch (p.CP) {
    case 4: return launch_head_cp<4>(h, p, stream);
    case 8: return launch_head_cp<8>(h, p, stream);
    case 12: return launch_head_cp<12>(h, p, stream);
    case 16: return launch_head_cp<16>(h, p, stream);
    case 20: return launch_head_cp<20>(h, p, stream);
    case 24: return launch_head_cp<24>(h, p, stream);
    case 28: return launch_head_cp<28>(h, p, stream);
    default: return launch_head_cp<32>(h, p, stream);
  }
}

}  // namespace hf
