// Fused preconditioned-CG vector pass for sm_100a.
//
// Replaces the ~25 element-wise/reduction launches and >=4 host syncs per iteration of the
// reference loop body (hessianfree/cg.py:205-224, _terminate_cg :80-118, _postprocess_pAp :123-147),
// the damping add (optimizer.py:266) and the diagonal-preconditioner apply (preconditioners.py:125)
// by ONE cooperative persistent launch.  Each CTA owns a contiguous slice of the P-vector and keeps
// p and Ap (later y) in shared memory across the two grid-wide reductions (pAp -> alpha, r.y -> beta),
// so HBM sees exactly: read Bp,p,x,r,b,minv; write x,r,p = 9 streams (36 P bytes in FP32).
//
// Reductions are deterministic (fixed per-thread order -> warp butterfly -> per-CTA partial ->
// fixed-order grid sum), so every CTA, every run and every data-parallel rank derives bit-identical
// alpha/beta and the split form (ALPHA launch, user M(r), BETA launch) reproduces the fused form
// bit for bit (reference tests/test_cg.py:217-218 demands M=None == M=identity exactly).
#include <math.h>

#include "common.cuh"

namespace hf {

struct PcgState {
  hf_pcg_status st;
  int32_t first_beta;  // split-mode init done, p = -y still pending
  int32_t martens;
  int64_t max_iter;
  int32_t* progress;  // optional host-mapped {iter, reason} pair (hf_pcg_set_progress), written after every update
  // grid all-reduce slots: [0] p.Ap  [1] r.r,(r-b).x,r.y  [2] r.y of a BETA-only launch  [3] start-up sums
  ReduceSlot slots[4][kMaxCtas];
  double m_iters[2];  // really max_iter + 2 entries
};

constexpr int kThreads = 1024;

// Publish {iter, reason} to pinned host memory so the host can follow the solve without a copy or a sync.  iter is
// written (and made visible system-wide) before reason: a host that reads reason != 0 first and iter second sees the
// final iteration count.
__device__ __forceinline__ void publish_progress(const PcgState* s, int iter, int reason) {
  volatile int32_t* m = s->progress;
  if (!m) return;
  m[0] = iter;
  __threadfence_system();
  m[1] = reason;
}

// optional phase trace (tools/pcg_trace.py): per CTA, %globaltimer at the phase boundaries of pcg_iter_kernel
__device__ unsigned long long* g_pcg_trace = nullptr;
__device__ __forceinline__ void trace_mark(int slot) {
  if (g_pcg_trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_pcg_trace[blockIdx.x * 8 + slot] = t;
  }
}
constexpr size_t kResidentBytes = 200 * 1024;  // dynamic smem budget for the resident p / Ap slices

template <typename T>
struct alignas(16) Vec {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

template <typename T>
__device__ __forceinline__ Vec<T> load_vec(const T* __restrict__ base, int64_t g, int64_t P) {
  Vec<T> out;
  const int64_t i = g * Vec<T>::N;
  if (i + Vec<T>::N <= P) {
    out = *reinterpret_cast<const Vec<T>*>(base + i);
  } else {
#pragma unroll
    for (int e = 0; e < Vec<T>::N; ++e) out.v[e] = (i + e < P) ? base[i + e] : T(0);
  }
  return out;
}

template <typename T>
__device__ __forceinline__ void store_vec(T* __restrict__ base, int64_t g, int64_t P, const Vec<T>& val) {
  const int64_t i = g * Vec<T>::N;
  if (i + Vec<T>::N <= P) {
    *reinterpret_cast<Vec<T>*>(base + i) = val;
  } else {
#pragma unroll
    for (int e = 0; e < Vec<T>::N; ++e)
      if (i + e < P) base[i + e] = val.v[e];
  }
}

__device__ __forceinline__ void l2_prefetch(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// rounding-faithful a + s*b (two roundings, like the reference's out-of-place `x + alpha * p`)
__device__ __forceinline__ float mul_add(float s, float b, float a) { return __fadd_rn(a, __fmul_rn(s, b)); }
__device__ __forceinline__ double mul_add(double s, double b, double a) { return __dadd_rn(a, __dmul_rn(s, b)); }
__device__ __forceinline__ float fma_acc(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_acc(double a, double b, double c) { return __fma_rn(a, b, c); }

template <typename T>
struct IterArgs {
  int64_t P;
  int64_t groups_per_cta;
  PcgState* state;
  // distinct vectors (y_ext may equal r, but then neither is written): lets loads of the next group be issued
  // before the stores of the current one
  const T* __restrict__ Bp;
  const T* __restrict__ b;
  const T* __restrict__ minv;
  const T* __restrict__ y_ext;
  T* __restrict__ x;
  T* __restrict__ r;
  T* __restrict__ p;
  T* __restrict__ snapshot;
  float* __restrict__ p_lo;
  double lambda;
  int phase;
  int resident;
};

// Termination tests in the reference's order (cg.py:96-115); arithmetic in the solve dtype T where the
// reference's is.  Returns the hf_cg_reason (0 = keep going).
template <typename T>
__device__ __forceinline__ int terminate_cg(PcgState* s, int iter, double rr, double rbx, bool writer, double* m_out,
                                            double* rnorm_out) {
  const T rnorm = (T)sqrt(rr);
  const T m_new = (T)(0.5 * rbx);
  *m_out = (double)m_new;
  *rnorm_out = (double)rnorm;
  int reason = HF_CG_RUNNING;
  if (s->martens) {
    if (writer) s->m_iters[iter] = (double)m_new;
    const int k = max(10, iter / 10);
    if (k < iter) {
      const T num = m_new - (T)__ldcg(&s->m_iters[iter - k]);
      const T den = m_new - (T)__ldcg(&s->m_iters[0]);
      if (num / den < (T)5e-4) reason = HF_CG_MARTENS;
    }
  }
  if (reason == HF_CG_RUNNING) {
    if (iter >= s->max_iter)
      reason = HF_CG_MAXITER;
    else if (isnan(rnorm))
      reason = HF_CG_DIVERGED;
    else if ((double)rnorm < s->st.res_bound)
      reason = HF_CG_TOL;
  }
  return reason;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) pcg_iter_kernel(IterArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[4 * 33];
  using V = Vec<T>;
  constexpr int N = V::N;
  PcgState* s = a.state;
  if (s->st.reason != HF_CG_RUNNING) return;  // solver already terminated: leave everything untouched

  const unsigned n_ctas = gridDim.x;
  const int64_t total_groups = (a.P + N - 1) / N;
  const int64_t g0 = (int64_t)blockIdx.x * a.groups_per_cta;
  const int64_t g1 = min(total_groups, g0 + a.groups_per_cta);
  const int nloc = g1 > g0 ? (int)(g1 - g0) : 0;
  V* sp = reinterpret_cast<V*>(smem_raw);  // resident p slice
  V* sq = sp + a.groups_per_cta;           // resident Ap slice, later y slice
  const bool writer = blockIdx.x == 0 && threadIdx.x == 0;
  const T lam = (T)a.lambda;
  const double ry_old = s->st.ry;
  const int iter = s->st.iter + 1;
  const bool first_beta = s->first_beta != 0;
  const unsigned gen = (unsigned)iter + (first_beta ? 0u : 1u);  // fresh for every use of a slot array
  double ry_new = 0.0;

  trace_mark(0);
  if (a.phase & HF_PCG_ALPHA) {
    // ---- phase 1: Ap = Bp + lambda p, partial p.Ap ------------------------------------------------
    T acc = T(0);
#pragma unroll 2
    for (int j = threadIdx.x; j < nloc; j += kThreads) {
      const V p4 = load_vec(a.p, g0 + j, a.P);
      V q4 = load_vec(a.Bp, g0 + j, a.P);
      // Pull the operands of phase 2 into L2 now: HBM then streams all six inputs back to back while the grid-wide
      // reduction below is in flight, instead of idling until alpha is known.
      l2_prefetch(a.x + (g0 + j) * N);
      l2_prefetch(a.r + (g0 + j) * N);
      l2_prefetch(a.b + (g0 + j) * N);
      if (a.minv) l2_prefetch(a.minv + (g0 + j) * N);
#pragma unroll
      for (int e = 0; e < N; ++e) {
        q4.v[e] = mul_add(lam, p4.v[e], q4.v[e]);
        acc = fma_acc(p4.v[e], q4.v[e], acc);
      }
      if (a.resident) {
        sp[j] = p4;
        sq[j] = q4;
      }
    }
    double red1[1] = {(double)acc};
    trace_mark(1);
    grid_allreduce<1>(s->slots[0], n_ctas, gen, red1, scratch);
    trace_mark(2);
    const double pAp = red1[0];
    const double alpha = ry_old / pAp;
    const T al = (T)alpha;

    // ---- phase 2: x += alpha p, r += alpha Ap, y = M r, partial r.r, (r-b).x, r.y ------------------
    const bool want_y = (a.phase & HF_PCG_BETA) != 0;
    T acc_rr = T(0), acc_m = T(0), acc_ry = T(0);
#pragma unroll 2
    for (int j = threadIdx.x; j < nloc; j += kThreads) {
      V p4, q4;
      if (a.resident) {
        p4 = sp[j];
        q4 = sq[j];
      } else {
        p4 = load_vec(a.p, g0 + j, a.P);
        q4 = load_vec(a.Bp, g0 + j, a.P);
#pragma unroll
        for (int e = 0; e < N; ++e) q4.v[e] = mul_add(lam, p4.v[e], q4.v[e]);
      }
      V x4 = load_vec((const T*)a.x, g0 + j, a.P);
      V r4 = load_vec((const T*)a.r, g0 + j, a.P);
      const V b4 = load_vec(a.b, g0 + j, a.P);
      V y4;
      if (want_y && a.minv) y4 = load_vec(a.minv, g0 + j, a.P);
#pragma unroll
      for (int e = 0; e < N; ++e) {
        x4.v[e] = mul_add(al, p4.v[e], x4.v[e]);
        r4.v[e] = mul_add(al, q4.v[e], r4.v[e]);
        acc_rr = fma_acc(r4.v[e], r4.v[e], acc_rr);
        acc_m = fma_acc(r4.v[e] - b4.v[e], x4.v[e], acc_m);
      }
      store_vec(a.x, g0 + j, a.P, x4);
      store_vec(a.r, g0 + j, a.P, r4);
      if (a.snapshot) store_vec(a.snapshot, g0 + j, a.P, x4);
      if (want_y) {
#pragma unroll
        for (int e = 0; e < N; ++e) {
          y4.v[e] = a.minv ? y4.v[e] * r4.v[e] : r4.v[e];
          acc_ry = fma_acc(r4.v[e], y4.v[e], acc_ry);
        }
        if (a.resident) sq[j] = y4;
      }
    }
    double red3[3] = {(double)acc_rr, (double)acc_m, (double)acc_ry};
    trace_mark(3);
    grid_allreduce<3>(s->slots[1], n_ctas, gen, red3, scratch);
    trace_mark(4);
    ry_new = red3[2];
    double m_new, rnorm;
    const int reason = terminate_cg<T>(s, iter, red3[0], red3[1], writer, &m_new, &rnorm);
    if (writer) {
      s->st.iter = iter;
      s->st.pAp = pAp;
      s->st.alpha = alpha;
      s->st.rnorm = rnorm;
      s->st.m = m_new;
      if (!(pAp > 0.0) && s->st.nonpos_iter == 0) {
        s->st.nonpos_iter = iter;
        s->st.nonpos_pAp = pAp;
      }
      s->st.reason = reason;
      publish_progress(s, iter, reason);
    }
    if (reason != HF_CG_RUNNING || !want_y) return;
  } else {
    // ---- split form, BETA launch: y supplied by the caller, r.y over the same thread mapping ------
    T acc_ry = T(0);
    for (int j = threadIdx.x; j < nloc; j += kThreads) {
      const V r4 = load_vec((const T*)a.r, g0 + j, a.P);
      V y4;
      if (a.y_ext)
        y4 = load_vec(a.y_ext, g0 + j, a.P);
      else if (a.minv)
        y4 = load_vec(a.minv, g0 + j, a.P);
#pragma unroll
      for (int e = 0; e < N; ++e) {
        if (!a.y_ext) y4.v[e] = a.minv ? y4.v[e] * r4.v[e] : r4.v[e];
        acc_ry = fma_acc(r4.v[e], y4.v[e], acc_ry);
      }
      if (a.resident) sq[j] = y4;
    }
    double red1[1] = {(double)acc_ry};
    // same per-thread order, same block and grid summation order as the fused launch -> same bits
    grid_allreduce<1>(s->slots[2], n_ctas, gen, red1, scratch);
    ry_new = red1[0];
  }

  // ---- phase 3: beta = ry'/ry, p = -y + beta p -----------------------------------------------------
  const double beta = first_beta ? 0.0 : ry_new / ry_old;
  const T be = (T)beta;
  const bool p_resident = a.resident && (a.phase & HF_PCG_ALPHA);
  for (int j = threadIdx.x; j < nloc; j += kThreads) {
    V y4, p4;
    if (a.resident) {
      y4 = sq[j];
    } else {
      const V r4 = load_vec((const T*)a.r, g0 + j, a.P);
      if (a.y_ext)
        y4 = load_vec(a.y_ext, g0 + j, a.P);
      else if (a.minv)
        y4 = load_vec(a.minv, g0 + j, a.P);
#pragma unroll
      for (int e = 0; e < N; ++e)
        if (!a.y_ext) y4.v[e] = a.minv ? y4.v[e] * r4.v[e] : r4.v[e];
    }
    if (first_beta) {
#pragma unroll
      for (int e = 0; e < N; ++e) p4.v[e] = -y4.v[e];
    } else {
      p4 = p_resident ? sp[j] : load_vec((const T*)a.p, g0 + j, a.P);
#pragma unroll
      for (int e = 0; e < N; ++e) p4.v[e] = mul_add(be, p4.v[e], -y4.v[e]);
    }
    store_vec(a.p, g0 + j, a.P, p4);
    if (a.p_lo) {
      if constexpr (sizeof(T) == 4) {
        Vec<float> lo;
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const float hi = __uint_as_float(__float_as_uint((float)p4.v[e]) & 0xffffe000u);
          lo.v[e] = (float)p4.v[e] - hi;
        }
        store_vec(a.p_lo, g0 + j, a.P, lo);
      }
    }
  }
  trace_mark(5);
  if (writer) {
    s->st.ry = ry_new;
    s->st.beta = beta;
    s->first_beta = 0;
  }
}

template <typename T>
struct InitArgs {
  int64_t P;
  int64_t groups_per_cta;
  PcgState* state;
  const T* Bx0;
  const T* x0;
  const T* b;
  const T* minv;
  T* x;
  T* r;
  T* p;
  double lambda, tol, atol;
  int64_t max_iter;
  int martens;
  int split;
};

// cg.py:186-192: x = x0, r = A(x0) - b, m_0 = 0.5 (r-b).x0, y = M r, ry = r.y, p = -y; res_bound (:75-76).
template <typename T>
__global__ void __launch_bounds__(kThreads, 1) pcg_init_kernel(InitArgs<T> a) {
  __shared__ double scratch[4 * 33];
  using V = Vec<T>;
  constexpr int N = V::N;
  PcgState* s = a.state;
  const unsigned n_ctas = gridDim.x;
  const int64_t total_groups = (a.P + N - 1) / N;
  const int64_t g0 = (int64_t)blockIdx.x * a.groups_per_cta;
  const int64_t g1 = min(total_groups, g0 + a.groups_per_cta);
  const int nloc = g1 > g0 ? (int)(g1 - g0) : 0;
  const T lam = (T)a.lambda;
  T acc_bb = T(0), acc_m = T(0), acc_ry = T(0);
  for (int j = threadIdx.x; j < nloc; j += kThreads) {
    const V b4 = load_vec(a.b, g0 + j, a.P);
    V x4, r4, y4;
    if (a.x0) {
      x4 = load_vec(a.x0, g0 + j, a.P);
      r4 = load_vec(a.Bx0, g0 + j, a.P);
    }
    if (!a.split && a.minv) y4 = load_vec(a.minv, g0 + j, a.P);
#pragma unroll
    for (int e = 0; e < N; ++e) {
      if (a.x0) {
        r4.v[e] = mul_add(lam, x4.v[e], r4.v[e]) - b4.v[e];
      } else {
        x4.v[e] = T(0);
        r4.v[e] = T(0) - b4.v[e];
      }
      acc_bb = fma_acc(b4.v[e], b4.v[e], acc_bb);
      acc_m = fma_acc(r4.v[e] - b4.v[e], x4.v[e], acc_m);
    }
    store_vec(a.x, g0 + j, a.P, x4);
    store_vec(a.r, g0 + j, a.P, r4);
    if (!a.split) {
      V p4;
#pragma unroll
      for (int e = 0; e < N; ++e) {
        y4.v[e] = a.minv ? y4.v[e] * r4.v[e] : r4.v[e];
        acc_ry = fma_acc(r4.v[e], y4.v[e], acc_ry);
        p4.v[e] = -y4.v[e];
      }
      store_vec(a.p, g0 + j, a.P, p4);
    }
  }
  double red[3] = {(double)acc_bb, (double)acc_m, (double)acc_ry};
  grid_allreduce<3>(s->slots[3], n_ctas, 1u, red, scratch);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const T bnorm = (T)sqrt(red[0]);
    double bound = a.tol * (double)bnorm;
    if (a.atol >= 0.0) bound = fmax(bound, a.atol);
    const T m0 = (T)(0.5 * red[1]);
    s->st.iter = 0;
    s->st.reason = HF_CG_RUNNING;
    s->st.nonpos_iter = 0;
    s->st.nonpos_pAp = 0.0;
    s->st.ry = red[2];
    s->st.pAp = 0.0;
    s->st.alpha = 0.0;
    s->st.beta = 0.0;
    s->st.rnorm = 0.0;
    s->st.m = (double)m0;
    s->st.res_bound = bound;
    s->first_beta = a.split ? 1 : 0;
    s->martens = a.martens;
    s->max_iter = a.max_iter;
    s->m_iters[0] = (double)m0;
    publish_progress(s, 0, HF_CG_RUNNING);
  }
}

template <typename T>
__global__ void precond_power_kernel(int64_t P, const T* __restrict__ d, T damping, T neg_exponent, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = pow(d[i] + damping, neg_exponent);
}

template <typename T>
__global__ void axpy_out_kernel(int64_t P, const T* __restrict__ a, T alpha, const T* __restrict__ b, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = mul_add(alpha, b[i], a[i]);
}

struct Geometry {
  unsigned n_ctas;
  int64_t groups_per_cta;
  int resident;
  size_t smem;
};

static Geometry pcg_geometry(int64_t P, size_t elem) {
  const int64_t per_vec = 16 / (int64_t)elem;
  const int64_t groups = (P + per_vec - 1) / per_vec;
  int64_t n = (groups + 255) / 256;
  const int64_t cap = sm_count() < kMaxCtas ? sm_count() : kMaxCtas;
  if (n > cap) n = cap;
  if (n < 1) n = 1;
  Geometry g;
  g.n_ctas = (unsigned)n;
  g.groups_per_cta = (groups + n - 1) / n;
  const size_t need = (size_t)g.groups_per_cta * 32;
  g.resident = need <= kResidentBytes ? 1 : 0;
  g.smem = g.resident ? need : 0;
  return g;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
static int launch_iter(int64_t P, void* d_state, int phase, const void* Bp, const void* b, const void* minv,
                       const void* y_ext, double lambda, void* x, void* r, void* p, void* snapshot, void* p_lo,
                       cudaStream_t stream) {
  const Geometry g = pcg_geometry(P, sizeof(T));
  static bool seen[64] = {};
  if (first_use_on_device(seen))  // the opt-in is per device
    HF_CUDA(cudaFuncSetAttribute(pcg_iter_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResidentBytes));
  IterArgs<T> a;
  a.P = P;
  a.groups_per_cta = g.groups_per_cta;
  a.state = static_cast<PcgState*>(d_state);
  a.Bp = static_cast<const T*>(Bp);
  a.b = static_cast<const T*>(b);
  a.minv = static_cast<const T*>(minv);
  a.y_ext = static_cast<const T*>(y_ext);
  a.x = static_cast<T*>(x);
  a.r = static_cast<T*>(r);
  a.p = static_cast<T*>(p);
  a.snapshot = static_cast<T*>(snapshot);
  a.p_lo = static_cast<float*>(p_lo);
  a.lambda = lambda;
  a.phase = phase;
  a.resident = g.resident;
  void* params[] = {&a};
  HF_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_iter_kernel<T>, dim3(g.n_ctas), dim3(kThreads), params, g.smem,
                                      stream));
  note_launch();
  return HF_OK;
}

template <typename T>
static int launch_init(int64_t P, void* d_state, const void* Bx0, const void* x0, const void* b, const void* minv,
                       double lambda, double tol, double atol, int64_t max_iter, int martens, int split, void* x,
                       void* r, void* p, cudaStream_t stream) {
  const Geometry g = pcg_geometry(P, sizeof(T));
  InitArgs<T> a;
  a.P = P;
  a.groups_per_cta = g.groups_per_cta;
  a.state = static_cast<PcgState*>(d_state);
  a.Bx0 = static_cast<const T*>(Bx0);
  a.x0 = static_cast<const T*>(x0);
  a.b = static_cast<const T*>(b);
  a.minv = static_cast<const T*>(minv);
  a.x = static_cast<T*>(x);
  a.r = static_cast<T*>(r);
  a.p = static_cast<T*>(p);
  a.lambda = lambda;
  a.tol = tol;
  a.atol = atol;
  a.max_iter = max_iter;
  a.martens = martens;
  a.split = split;
  void* params[] = {&a};
  HF_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_init_kernel<T>, dim3(g.n_ctas), dim3(kThreads), params, 0,
                                      stream));
  note_launch();
  return HF_OK;
}

}  // namespace hf

using namespace hf;

extern "C" {

size_t hf_pcg_state_bytes(int64_t max_iter) {
  if (max_iter < 0) max_iter = 0;
  return offsetof(PcgState, m_iters) + sizeof(double) * (size_t)(max_iter + 2);
}

size_t hf_pcg_m_iters_offset(void) { return offsetof(PcgState, m_iters); }

int hf_pcg_set_progress(void* d_state, int32_t* mapped_host_pair, void* stream) {
  HF_REQUIRE(d_state, HF_ERR_INVALID, "hf_pcg_set_progress: null state");
  // pageable source: the runtime stages the 8 bytes before returning, so the local may go out of scope
  HF_CUDA(cudaMemcpyAsync(static_cast<char*>(d_state) + offsetof(PcgState, progress), &mapped_host_pair,
                          sizeof(mapped_host_pair), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return HF_OK;
}

int hf_debug_pcg_trace(void* d_buf) {
  unsigned long long* p = static_cast<unsigned long long*>(d_buf);
  HF_CUDA(cudaMemcpyToSymbol(g_pcg_trace, &p, sizeof(p)));
  return HF_OK;
}

int hf_pcg_init(int dtype, int64_t P, void* d_state, size_t state_bytes, const void* d_Bx0, const void* d_x0,
                const void* d_b, const void* d_minv, double lambda, double tol, double atol, int64_t max_iter,
                int martens, int split, void* d_x, void* d_r, void* d_p, void* stream) {
  HF_REQUIRE(dtype == HF_F32 || dtype == HF_F64, HF_ERR_INVALID, "hf_pcg_init: dtype must be HF_F32 or HF_F64");
  HF_REQUIRE(P > 0 && max_iter >= 1, HF_ERR_INVALID, "hf_pcg_init: need P > 0 and max_iter >= 1");
  HF_REQUIRE(d_state && d_b && d_x && d_r && d_p, HF_ERR_INVALID, "hf_pcg_init: null pointer");
  HF_REQUIRE((d_x0 == nullptr) == (d_Bx0 == nullptr), HF_ERR_INVALID, "hf_pcg_init: x0 and Bx0 go together");
  HF_REQUIRE(state_bytes >= hf_pcg_state_bytes(max_iter), HF_ERR_WORKSPACE, "hf_pcg_init: state block too small");
  HF_REQUIRE(aligned16(d_state) && aligned16(d_b) && aligned16(d_x) && aligned16(d_r) && aligned16(d_p) &&
                 aligned16(d_Bx0) && aligned16(d_x0) && aligned16(d_minv),
             HF_ERR_INVALID, "hf_pcg_init: all vectors must be 16-byte aligned");
  // the generation words of the reduce slots must start from zero; everything else is written by the kernel
  HF_CUDA(cudaMemsetAsync(static_cast<char*>(d_state) + offsetof(PcgState, slots), 0, sizeof(ReduceSlot) * 4 * kMaxCtas,
                          (cudaStream_t)stream));
  if (dtype == HF_F32)
    return launch_init<float>(P, d_state, d_Bx0, d_x0, d_b, d_minv, lambda, tol, atol, max_iter, martens, split, d_x,
                              d_r, d_p, (cudaStream_t)stream);
  return launch_init<double>(P, d_state, d_Bx0, d_x0, d_b, d_minv, lambda, tol, atol, max_iter, martens, split, d_x,
                             d_r, d_p, (cudaStream_t)stream);
}

int hf_pcg_iter(int dtype, int64_t P, void* d_state, int phase, const void* d_Bp, const void* d_b,
                const void* d_minv, const void* d_y_ext, double lambda, void* d_x, void* d_r, void* d_p,
                void* d_snapshot, void* d_p_lo, void* stream) {
  HF_REQUIRE(dtype == HF_F32 || dtype == HF_F64, HF_ERR_INVALID, "hf_pcg_iter: dtype must be HF_F32 or HF_F64");
  HF_REQUIRE(phase >= HF_PCG_ALPHA && phase <= HF_PCG_FUSED, HF_ERR_INVALID, "hf_pcg_iter: bad phase %d", phase);
  HF_REQUIRE(P > 0 && d_state && d_x && d_r && d_p && d_b, HF_ERR_INVALID, "hf_pcg_iter: null pointer");
  HF_REQUIRE(!(phase & HF_PCG_ALPHA) || d_Bp, HF_ERR_INVALID, "hf_pcg_iter: the ALPHA phase needs Bp");
  HF_REQUIRE(!(phase == HF_PCG_FUSED && d_y_ext), HF_ERR_INVALID, "hf_pcg_iter: y_ext only with HF_PCG_BETA");
  HF_REQUIRE(!(d_p_lo && dtype != HF_F32), HF_ERR_INVALID, "hf_pcg_iter: p_lo is FP32 only");
  HF_REQUIRE(aligned16(d_Bp) && aligned16(d_b) && aligned16(d_minv) && aligned16(d_y_ext) && aligned16(d_x) &&
                 aligned16(d_r) && aligned16(d_p) && aligned16(d_snapshot) && aligned16(d_p_lo),
             HF_ERR_INVALID, "hf_pcg_iter: all vectors must be 16-byte aligned");
  if (dtype == HF_F32)
    return launch_iter<float>(P, d_state, phase, d_Bp, d_b, d_minv, d_y_ext, lambda, d_x, d_r, d_p, d_snapshot,
                              d_p_lo, (cudaStream_t)stream);
  return launch_iter<double>(P, d_state, phase, d_Bp, d_b, d_minv, d_y_ext, lambda, d_x, d_r, d_p, d_snapshot, d_p_lo,
                             (cudaStream_t)stream);
}

int hf_precond_power(int dtype, int64_t P, const void* d_diag, double damping, double exponent, void* d_out,
                     void* stream) {
  HF_REQUIRE(P > 0 && d_diag && d_out, HF_ERR_INVALID, "hf_precond_power: bad arguments");
  const int threads = 256;
  int64_t blocks = (P + threads - 1) / threads;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  if (dtype == HF_F32)
    precond_power_kernel<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        P, (const float*)d_diag, (float)damping, (float)-exponent, (float*)d_out);
  else if (dtype == HF_F64)
    precond_power_kernel<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        P, (const double*)d_diag, damping, -exponent, (double*)d_out);
  else
    HF_REQUIRE(false, HF_ERR_INVALID, "hf_precond_power: bad dtype");
  HF_LAUNCH_CHECK();
  return HF_OK;
}

int hf_axpy_out(int dtype, int64_t P, const void* d_a, double alpha, const void* d_b, void* d_out, void* stream) {
  HF_REQUIRE(P > 0 && d_a && d_b && d_out, HF_ERR_INVALID, "hf_axpy_out: bad arguments");
  const int threads = 256;
  int64_t blocks = (P + threads - 1) / threads;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  if (dtype == HF_F32)
    axpy_out_kernel<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(P, (const float*)d_a, (float)alpha,
                                                                                  (const float*)d_b, (float*)d_out);
  else if (dtype == HF_F64)
    axpy_out_kernel<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(P, (const double*)d_a, alpha,
                                                                                   (const double*)d_b, (double*)d_out);
  else
    HF_REQUIRE(false, HF_ERR_INVALID, "hf_axpy_out: bad dtype");
  HF_LAUNCH_CHECK();
  return HF_OK;
}

}  // extern "C"
