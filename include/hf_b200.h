/*
 * hf_b200.h -- C ABI of the B200-native Hessian-free inner solve.
 *
 * Drop-in boundary for the one hot path of ltatzel/PyTorchHessianFree that this
 * repository accelerates: the preconditioned-CG Newton-step solve and the
 * curvature-matrix-vector products it calls.  The reference has no FFI; what
 * is replaced are its three nested Python seams:
 *
 *   hf_pcg_*              <- hessianfree/cg.py:9-231 (cg, _terminate_cg, _postprocess_pAp),
 *                            the damping add at optimizer.py:266 and the diagonal
 *                            preconditioner apply at preconditioners.py:125
 *   hf_net_* / hf_lin_*   <- the autograd graph captured by the closures at
 *                            optimizer.py:241-247 and the forward/grad at :223,:231-234
 *   hf_ggn_matvec         <- HessianFree._Gv, optimizer.py:457-462 (BackPACK ggnvp)
 *   hf_hessian_matvec     <- HessianFree._Hv, optimizer.py:450-455 (BackPACK hvp)
 *   hf_fisher_diag        <- diag_EF_backpack, preconditioners.py:11-60 (BackPACK SumGradSquared)
 *   hf_precond_power      <- diag_to_preconditioner, preconditioners.py:108-127
 *
 * Conventions
 *   - every pointer named d_* is a raw CUDA device pointer borrowed from the caller
 *     (torch.Tensor.data_ptr()); the library never allocates, frees or retains device
 *     memory: scratch comes from caller-provided workspaces whose size it reports;
 *   - every call that launches work takes the cudaStream_t (as void*) to launch on and
 *     returns immediately (no host synchronisation inside the library);
 *   - every call returns 0 on success, a negative hf_status otherwise;
 *     hf_last_error_string() describes the last failure on the calling thread;
 *     nothing throws across this boundary;
 *   - all vectors on the path are contiguous and use the reference's flat layout: the
 *     trainable parameters in param_groups[0]["params"] order, each flattened row-major
 *     (optimizer.py:122, utils.py:41-76); for nn.Linear: weight[out,in] then bias[out];
 *   - handles are not thread-safe (the reference is single-threaded).
 */
#ifndef HF_B200_H
#define HF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HF_ABI_VERSION 2

enum hf_status {
  HF_OK = 0,
  HF_ERR_INVALID = -1,     /* bad argument                       -> ValueError   */
  HF_ERR_UNSUPPORTED = -2, /* shape/op the kernels do not cover   -> NotImplementedError */
  HF_ERR_WORKSPACE = -3,   /* workspace too small / misaligned    -> ValueError   */
  HF_ERR_CUDA = -4         /* CUDA runtime error                  -> RuntimeError */
};

enum hf_dtype { HF_F32 = 0, HF_F64 = 1 };
enum hf_act { HF_ACT_NONE = 0, HF_ACT_RELU = 1, HF_ACT_SIGMOID = 2, HF_ACT_TANH = 3 };
enum hf_loss { HF_LOSS_MSE = 0, HF_LOSS_SOFTMAX_CE = 1, HF_LOSS_SIGMOID_BCE = 2 };
enum hf_reduction { HF_RED_MEAN = 0, HF_RED_SUM = 1 };

/* reason codes written to the solver state; the strings are the reference's (cg.py:103-115) */
enum hf_cg_reason {
  HF_CG_RUNNING = 0,
  HF_CG_MARTENS = 1,   /* "Convergence (Martens)"    */
  HF_CG_MAXITER = 2,   /* "Number of iterations"     */
  HF_CG_DIVERGED = 3,  /* "Divergence"               */
  HF_CG_TOL = 4        /* "Convergence (tolerances)" */
};

int hf_abi_version(void);
const char* hf_last_error_string(void);
/* number of SMs / co-resident CTAs the fused solver kernel will use on the current device */
int hf_device_sm_count(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long hf_debug_launch_count(void);
/* profiling aid: if d_buf (>= 8 * 256 uint64) is non-NULL, hf_pcg_iter records %globaltimer at its phase boundaries
 * per CTA: [cta][0] start, [1] partial p.Ap ready, [2] alpha known, [3] x/r updated, [4] beta known, [5] p written */
int hf_debug_pcg_trace(void* d_buf);
/* same for the tcgen05 contraction kernel: [cta][0] entry, [1] prologue done (barriers, TMEM), [2] first stage landed,
 * [3] accumulator complete, [4] epilogue done, [5] TMEM -> shared done.  Consecutive launches after the call write
 * consecutive blocks of 1024 CTA records (d_buf >= launches * 1024 * 8 uint64), so gaps between kernels can be read
 * off too; NULL switches the trace off. */
int hf_debug_tc_trace(void* d_buf);
/* the same for the CTA-pair kernel on pre-split operands (gemm_tc2.cu), one block of 4096 CTA records that every launch
 * overwrites: [cta][0] entry, [1] prologue done, [2] first stage landed, [3] accumulator complete, [4] first 128 columns
 * staged, [5] first 128 columns stored, [6] all stored, [7] pair released (d_buf >= 4096 * 8 uint64) */
int hf_debug_tc2_trace(void* d_buf);
/* per-k-block pipeline trace of CTA (0,0,0) of the same kernel, first 64 k-blocks: [it][0] TMA issued, [1] raw tiles seen
 * by the splitters, [2] split done, [3] MMAs issued, [4] producer's slot wait done (d_buf >= 64 * 8 uint64) */
int hf_debug_tc_trace_iters(void* d_buf);

/* ------------------------------------------------------------------------------------------
 * Fused PCG vector pass  (cg.py:186-224 + optimizer.py:266 + preconditioners.py:125)
 * ------------------------------------------------------------------------------------------
 * One cooperative, persistent launch per CG iteration does
 *     Ap = Bp + lambda p;  pAp = p.Ap;  alpha = ry/pAp;  x += alpha p;  r += alpha Ap;
 *     ||r||, m = 0.5 (r-b).x, termination tests in the reference's order;
 *     y = minv * r (or r);  ry' = r.y;  beta = ry'/ry;  p = -y + beta p
 * in a single HBM pass (36 P bytes with preconditioner, FP32): each CTA keeps its slice of p and
 * Ap/y in shared memory across the two grid-wide reductions.  Scalars live in the device-side
 * state block, so no host synchronisation is needed between iterations; once a termination test
 * fires, later launches return immediately and leave x, r, p and the state untouched.
 *
 * An arbitrary Python preconditioner (seam B2 of SURVEY.md) is served by the split form:
 * hf_pcg_iter(..., phase = HF_PCG_ALPHA) -> caller computes y = M(r) -> hf_pcg_iter(..., HF_PCG_BETA).
 * Both forms share one kernel and one reduction order, so M=None and M=identity give
 * bit-identical iterates (reference tests/test_cg.py:217-218).
 */
enum hf_pcg_phase { HF_PCG_ALPHA = 1, HF_PCG_BETA = 2, HF_PCG_FUSED = 3 };

/* bytes of the device-side solver state for at most max_iter iterations */
size_t hf_pcg_state_bytes(int64_t max_iter);
/* byte offset, inside the state block, of the double array m_0..m_iter (cg.py:189, :97): the values of the
 * quadratic 0.5 x^T A x - b^T x, rounded to the solve dtype like the reference's               */
size_t hf_pcg_m_iters_offset(void);
/* Optional progress mirror: `mapped_host_pair` points at two int32 in pinned (device-mapped) host memory; after every
 * hf_pcg_init / hf_pcg_iter that changes the status the kernel stores {iter, reason} there (iter first), so the host
 * can follow the solve by reading memory -- no copy, no event, no synchronisation (the reference instead synchronises
 * >= 4 times per iteration: cg.py:102, :110, :114, :133).  Call before hf_pcg_init on a zeroed state block; NULL
 * detaches.  The pinned pair must outlive every launch that uses the state block.                               */
int hf_pcg_set_progress(void* d_state, int32_t* mapped_host_pair, void* stream);

/* Host-visible mirror of the head of the state block (copy sizeof(hf_pcg_status) bytes D2H). */
typedef struct {
  int32_t iter;        /* iterations completed                                            */
  int32_t reason;      /* hf_cg_reason                                                    */
  int32_t nonpos_iter; /* first iteration with pAp <= 0 (0 = none), cg.py:132-143         */
  int32_t pad_;
  double nonpos_pAp;
  double ry, pAp, alpha, beta, rnorm, m, res_bound;
} hf_pcg_status;

/* Start a solve: writes r = Bx0 + lambda x0 - b, m_0 = 0.5 (r-b).x0, res_bound = max(tol ||b||, atol)
 * (atol < 0 = None), y = minv*r | r, ry = r.y, p = -y, iter = 0.   d_Bx0 may be NULL when x0 == 0
 * (then d_x is zero-filled).  With split = 1 the y/ry/p part is skipped; finish it with
 * hf_pcg_iter(HF_PCG_BETA) after computing y (beta is forced to 0 on the first call).       */
int hf_pcg_init(int dtype, int64_t P, void* d_state, size_t state_bytes, const void* d_Bx0, const void* d_x0,
                const void* d_b, const void* d_minv, double lambda, double tol, double atol, int64_t max_iter,
                int martens, int split, void* d_x, void* d_r, void* d_p, void* stream);

/* One CG iteration (see above).  d_y_ext: externally supplied y for HF_PCG_BETA (NULL = use minv / r).
 * d_snapshot: if non-NULL, the updated x is also written there (cg.py:209-210).
 * d_p_lo: if non-NULL (FP32 only), receives p - tf32_trunc(p), the low word of the split-precision
 * operand the tensor-core matvec consumes.                                                   */
int hf_pcg_iter(int dtype, int64_t P, void* d_state, int phase, const void* d_Bp, const void* d_b,
                const void* d_minv, const void* d_y_ext, double lambda, void* d_x, void* d_r, void* d_p,
                void* d_snapshot, void* d_p_lo, void* stream);

/* d_out[i] = (d_diag[i] + damping)^(-exponent)   (preconditioners.py:125, hoisted out of the loop) */
int hf_precond_power(int dtype, int64_t P, const void* d_diag, double damping, double exponent, void* d_out,
                     void* stream);

/* d_out = d_a + alpha * d_b  (candidate parameters theta + alpha*s for optimizer.py:293) */
int hf_axpy_out(int dtype, int64_t P, const void* d_a, double alpha, const void* d_b, void* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layer program ("net") and its linearisation at a parameter point ("lin")
 * ------------------------------------------------------------------------------------------ */
typedef struct hf_net hf_net_t;
typedef struct hf_lin hf_lin_t;

/* HF_LAYER_CONV2D: nn.Conv2d(c_in, out_features, (k_h, k_w), stride, pad) on an [h_in, w_in] map; in_features must be
 * c_in*k_h*k_w, the weight [out, c_in, k_h, k_w] sits at w_offset exactly as PyTorch flattens it, inputs are NCHW when it
 * is the first layer.  HF_LAYER_AVGPOOL: average over the whole [h_in, w_in] map (in_features = out_features = c_in, no
 * parameters); the layers after it see one row per sample again.  The last layer must produce one row per sample.
 * Conv nets support loss, gradient, GGN and Hessian products (examples/run_allcnnc_cifar100_deepobs.py, eval mode); the
 * Fisher diagonal of conv nets is not lowered yet (HF_ERR_UNSUPPORTED). */
enum hf_layer_kind { HF_LAYER_LINEAR = 0, HF_LAYER_CONV2D = 1, HF_LAYER_AVGPOOL = 2 };

typedef struct {
  int32_t in_features;
  int32_t out_features;
  int32_t act;          /* hf_act applied after this affine layer                         */
  int32_t has_bias;
  int64_t w_offset;     /* offset of weight[out,in] in the flat trainable vector, -1 = frozen */
  int64_t b_offset;     /* offset of bias[out], -1 = frozen or absent                     */
  const float* d_w_frozen; /* device pointer to the weight when w_offset < 0              */
  const float* d_b_frozen; /* device pointer to the bias when has_bias && b_offset < 0    */
  int32_t kind;         /* hf_layer_kind; 0 = fully connected (the fields below are ignored) */
  int32_t c_in, h_in, w_in;
  int32_t k_h, k_w, stride, pad;
  int32_t h_out, w_out;
} hf_layer_desc;

/* n_params = length P of the flat trainable vector. */
int hf_net_create(const hf_layer_desc* layers, int32_t n_layers, int32_t loss, int32_t reduction, int64_t n_params,
                  hf_net_t** out);
void hf_net_destroy(hf_net_t* net);
/* contraction engine: 0 = FP32 SIMT tiles for every layer, 1 = tcgen05 split-precision tiles wherever the layer
 * shape meets the TMA alignment rules (feature widths multiples of 4 floats), SIMT otherwise.   */
int hf_net_set_engine(hf_net_t* net, int32_t engine);

#define HF_LIN_HESSIAN 1 /* also keep what the Hessian-vector product needs (delta_l, dL/da_l, Rz_l) */
#define HF_LIN_LOSS_ONLY 2 /* forward + loss only (step selection): smallest workspace              */

size_t hf_lin_workspace_bytes(const hf_net_t* net, int64_t batch, int32_t flags);
/* Bind a linearisation of `batch` samples to a caller-owned, 256-byte aligned workspace. */
int hf_lin_create(const hf_net_t* net, int64_t batch, int32_t flags, void* d_workspace, size_t workspace_bytes,
                  hf_lin_t** out);
void hf_lin_destroy(hf_lin_t* lin);

/* Forward pass at d_theta on this chunk; stores activations; adds this chunk's share of the loss
 * (already scaled for `mean` by n_total, the sample count of the WHOLE batch over all chunks and
 * ranks, optimizer.py:678-684) to *d_loss_acc (double).  d_targets: float[batch,C] (mse, bce) or
 * int64[batch] (softmax-ce).  d_x must stay valid for the life of the linearisation.         */
int hf_lin_forward(hf_lin_t* lin, const float* d_theta, const float* d_x, const void* d_targets, int64_t n_total,
                   double* d_loss_acc, void* stream);

/* d_grad (+)= this chunk's share of the gradient (optimizer.py:231-234, :751-754). */
int hf_lin_gradient(hf_lin_t* lin, const float* d_theta, float* d_grad, int32_t accumulate, void* stream);

/* d_out (+)= J^T H_loss J v on this chunk (optimizer.py:457-462).  d_skip: optional device int; when
 * non-zero at launch time the kernels return without touching d_out (solver already terminated). */
int hf_ggn_matvec(hf_lin_t* lin, const float* d_theta, const float* d_v, float* d_out, int32_t accumulate,
                  const int32_t* d_skip, void* stream);

/* d_out (+)= (d^2 loss / d theta^2) v on this chunk (optimizer.py:450-455); needs HF_LIN_HESSIAN
 * and a prior hf_lin_gradient.                                                               */
int hf_hessian_matvec(hf_lin_t* lin, const float* d_theta, const float* d_v, float* d_out, int32_t accumulate,
                      const int32_t* d_skip, void* stream);

/* The same products in two phases, for data-parallel callers that overlap communication with the sweep:
 * phase 0 = tangent forward pass, loss Hessian and the transposed sweep except the parameter gradient of the first
 * trainable layer; phase 1 = that gradient (the largest slice, formed last).  After phase 0 every entry of d_out
 * outside the first layer's span is final.  kind: 0 = GGN, 1 = Hessian.                          */
int hf_matvec_phase(hf_lin_t* lin, int32_t kind, const float* d_theta, const float* d_v, float* d_out, int32_t accumulate,
                    const int32_t* d_skip, void* stream, int32_t phase);
/* flat range [offset, offset+count) of the first trainable layer's parameters (count = 0 if not contiguous) */
int hf_net_first_layer_span(const hf_net_t* net, int64_t* offset, int64_t* count);

/* d_out (+)= sum_n g_n^2 ("sum") or (1/n_total) sum_n g_n^2 ("mean") on this chunk
 * (preconditioners.py:11-60 incl. the rescaling at :56-58).                                  */
int hf_fisher_diag(hf_lin_t* lin, const float* d_theta, float* d_out, int32_t accumulate, void* stream);

/* Tooling (tools/scale_probe.py): barrier flavour of hf_allreduce_multimem. bit 0: the entry barrier is a relaxed
 * rendezvous; bit 1: no per-thread system fence in front of the exit barrier.  Default 3. */
void hf_debug_allreduce_variant(int32_t v);

/* ------------------------------------------------------------------------------------------
 * The exchange step of the data-parallel path (new; the reference is single-device): in-place all-reduce(sum) of
 * count floats at element `offset` of a SYMMETRIC buffer through the NVSwitch multicast address (multimem.ld_reduce /
 * multimem.st, two-shot).  d_multicast: the multicast address of the buffer; d_signal_pads: device array of `world`
 * pointers to the ranks' zero-initialised, peer-mapped signal pads (>= max_blocks * world uint32 each) -- both as
 * torch.distributed._symmetric_memory hands them out.  offset % 4 == 0, count % (4 * world) == 0.  Every rank must make
 * the same call; d_skip as in hf_ggn_matvec (the flag is replicated, so all ranks skip together).
 * ------------------------------------------------------------------------------------------ */
int hf_allreduce_multimem(void* d_multicast, void* d_signal_pads, int32_t rank, int32_t world, int64_t offset, int64_t count,
                          int32_t max_blocks, const int32_t* d_skip, void* stream);

/* device pointers into the linearisation, for tests: logits [batch,C] */
const float* hf_lin_logits(const hf_lin_t* lin);

/* ------------------------------------------------------------------------------------------
 * Stand-alone contraction used by the kernels above, exported for unit tests and profiling:
 * C[M,N] = sum_s A_s[M,K] * B_s[N,K]^T  (n_pairs in {1,2}), element strides given per operand.
 * engine: 0 = FP32 SIMT tiles, 1 = tcgen05 split-precision 128x128 tiles (requires the alignment it reports),
 * 2 = tcgen05 256x256 CTA-pair tiles on pre-split operand images; engine 2 builds the images of its operands in
 * d_workspace first (at least hf_contract_workspace_bytes bytes, 256-byte aligned); 3 = engine 2 with the images left in
 * the workspace by a previous engine-2 call with the same arguments (times the tile kernel alone).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* d_ptr;
  int64_t stride_mn; /* element stride along M (for A) or N (for B) */
  int64_t stride_k;  /* element stride along K; one of the two strides must be 1 */
} hf_operand;

size_t hf_contract_workspace_bytes(int64_t M, int64_t N, int64_t K, int32_t n_pairs);
int hf_contract(int32_t engine, int64_t M, int64_t N, int64_t K, int32_t n_pairs, const hf_operand* A,
                const hf_operand* B, float* d_C, int64_t ldc, void* d_workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HF_B200_H */
