"""Preconditioned conjugate gradients on the fused sm_100a vector kernel.

Same call signature, return values and termination behaviour as the reference
solver (``hessianfree/cg.py:9-231``): ``cg(A, b, x0, M, max_iter, tol, atol,
martens_conv_crit, store_x_at_iters, verbose) -> (x_iters, m_iters, reason)``.
``A`` and ``M`` stay arbitrary Python callables (seam B2 of SURVEY.md); what is
replaced is everything between two calls of ``A``: the ~25 element-wise and
reduction launches and the >=4 host syncs per iteration of the reference loop
body (``cg.py:205-224``) become one cooperative launch (``hf_pcg_iter``) and one
status read.

How ``M`` is applied:

* ``M is None``                         -> fused launch, y = r;
* ``M`` is a :class:`DiagonalPreconditioner` (what ``diag_to_preconditioner`` of
  this package returns)                 -> fused launch, y = minv * r inside the kernel;
* any other callable                    -> split form: ALPHA launch, ``y = M(r)`` in
  Python, BETA launch.  Same kernel, same reduction order: ``M=None`` and
  ``M=identity`` give bit-identical iterates (reference ``tests/test_cg.py:217``).
"""
import ctypes as C
import time
from math import ceil, log
from warnings import warn

import torch

from . import _lib
from ._lib import PCG_ALPHA, PCG_BETA, PCG_FUSED, PcgStatus, REASONS


class DiagonalPreconditioner:
    """``x -> (diag + damping)^(-exponent) * x`` with the power hoisted out of the loop.

    Callable like the closure the reference builds (``preconditioners.py:124-127``), so it can be
    handed to any solver; :func:`cg` recognises it and applies ``minv`` inside the fused kernel.
    """

    def __init__(self, diag_vec, damping, exponent=0.75):
        _lib.require_cuda(diag_vec, "diag_vec")
        lib = _lib.load()
        d = _lib.vec(diag_vec.detach())
        self.diag_vec, self.damping, self.exponent = d, float(damping), float(exponent)
        self.minv = torch.empty_like(d)
        _lib.check(lib.hf_precond_power(_lib.dtype_code(d), d.numel(), d.data_ptr(), self.damping, self.exponent,
                                        self.minv.data_ptr(), _lib.stream()))

    def __call__(self, x):
        return torch.mul(self.minv, x)


def cg_storing_grid(max_iter, gamma=1.3):
    """Iterations ``ceil(gamma**j) - 1`` at which snapshots of x are kept (Martens 2010, sec. 4.6;
    reference ``cg.py:152-170``).  Evaluated in float32 on an integer ``arange`` like the reference,
    so the two grids agree entry for entry."""
    if gamma < 1.0:
        raise ValueError(f"Invalid gamma = {gamma}")
    top = ceil(log(max_iter + 1) / log(gamma))
    powers = gamma ** torch.arange(top + 1)
    return sorted(set((torch.ceil(powers) - 1).int().tolist()))


class _Solver:
    """Device buffers and state block of one solve; thin wrapper over hf_pcg_init / hf_pcg_iter."""

    def __init__(self, b, max_iter):
        self.lib = _lib.load()
        self.b = b
        self.P = b.numel()
        self.code = _lib.dtype_code(b)
        self.max_iter = int(max_iter)
        nbytes = self.lib.hf_pcg_state_bytes(self.max_iter)
        self.state = torch.zeros(nbytes, dtype=torch.uint8, device=b.device)
        self.x, self.r, self.p = (torch.empty_like(b) for _ in range(3))
        self._host = torch.empty(C.sizeof(PcgStatus), dtype=torch.uint8).pin_memory()
        self._m_off = self.lib.hf_pcg_m_iters_offset()
        # {iter, reason} published by the kernels into pinned host memory: the host follows the solve by reading it
        self._progress = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.progress = self._progress.numpy()
        _lib.check(self.lib.hf_pcg_set_progress(self.state.data_ptr(), self._progress.data_ptr(), _lib.stream()))

    def wait_for(self, iteration, timeout=60.0):
        """Spin (no CUDA call) until the device has finished ``iteration`` or stopped; returns (iter, reason)."""
        t0 = None
        while True:
            reason = int(self.progress[1])  # reason first: a non-zero reason makes the iter read below final
            done = int(self.progress[0])
            if reason != 0 or done >= iteration:
                return done, reason
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout:
                torch.cuda.current_stream().synchronize()  # surfaces a CUDA error if that is why nothing moves
                if int(self.progress[1]) == 0 and int(self.progress[0]) < iteration:
                    raise RuntimeError("pcg: the device made no progress (were the iterations enqueued?)")

    def init(self, Bx0, x0, minv, lam, tol, atol, martens, split):
        _lib.check(self.lib.hf_pcg_init(
            self.code, self.P, self.state.data_ptr(), self.state.numel(), _lib.ptr(Bx0), _lib.ptr(x0),
            self.b.data_ptr(), _lib.ptr(minv), float(lam), float(tol), -1.0 if atol is None else float(atol),
            self.max_iter, int(bool(martens)), int(bool(split)), self.x.data_ptr(), self.r.data_ptr(),
            self.p.data_ptr(), _lib.stream()))

    def iterate(self, phase, Bp=None, minv=None, y_ext=None, lam=0.0, snapshot=None, p_lo=None):
        _lib.check(self.lib.hf_pcg_iter(
            self.code, self.P, self.state.data_ptr(), phase, _lib.ptr(Bp), self.b.data_ptr(), _lib.ptr(minv),
            _lib.ptr(y_ext), float(lam), self.x.data_ptr(), self.r.data_ptr(), self.p.data_ptr(), _lib.ptr(snapshot),
            _lib.ptr(p_lo), _lib.stream()))

    @property
    def reason_ptr(self):
        """Device address of the reason word: non-zero once the solve has terminated."""
        return self.state.data_ptr() + PcgStatus.reason.offset

    def status(self):
        """Blocking read of the solver status (the one host sync per iteration of the generic path)."""
        self._host.copy_(self.state[: self._host.numel()], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return PcgStatus.from_buffer_copy(self._host.numpy().tobytes())

    def m_iters(self, count):
        raw = self.state[self._m_off: self._m_off + 8 * count].view(torch.float64)
        return raw.to(self.b.dtype)


def _warn_nonpositive(pAp, it):
    # message prefix is matched by the reference's tests (tests/test_optimizer_acc.py:122)
    warn(f"Directional curvature pAp = {pAp:.3e} <= 0 detected in cg-iteration {it}. "
         "This is a violation to the assumption of positive definiteness.")


def cg(
    A,
    b,
    x0=None,
    M=None,
    max_iter=None,
    tol=1e-5,
    atol=None,
    martens_conv_crit=False,
    store_x_at_iters=[],
    verbose=False,
):
    """Minimise ``0.5 x^T A x - b^T x`` (solve ``A x = b``, A s.p.d.) by preconditioned CG.

    Args and returns as in the reference (``cg.py:21-63``):
      A: callable ``v -> A v``;  b: right-hand side (CUDA, float32/float64, 1-D);
      x0: start (zeros if None);  M: callable approximating ``A^-1`` or None;
      max_iter: iteration cap (``b.numel()`` if None);
      tol, atol: stop when ``||r|| < max(tol*||b||, atol)``;
      martens_conv_crit: also stop on Martens' relative-progress test (window
        ``k = max(10, iter/10)``, threshold 5e-4) and return the values of the quadratic;
      store_x_at_iters: iterations whose x is kept (None = the gamma=1.3 grid); the final x is always kept.
    Returns ``(x_iters, m_iters, reason)``: list of tensors/None of length ``iterations+1``, list of 0-dim
    tensors (or None), and one of "Convergence (Martens)", "Number of iterations", "Divergence",
    "Convergence (tolerances)".
    """
    _lib.require_cuda(b, "b")
    b = _lib.vec(b.detach())
    max_iter = b.numel() if max_iter is None else int(max_iter)
    if store_x_at_iters is None:
        store_x_at_iters = cg_storing_grid(max_iter)
    keep = set(store_x_at_iters)

    fused_minv = M.minv if isinstance(M, DiagonalPreconditioner) else None
    split = M is not None and fused_minv is None
    if fused_minv is not None and (fused_minv.dtype != b.dtype or fused_minv.numel() != b.numel()):
        raise ValueError("preconditioner and right-hand side disagree in dtype or length")

    s = _Solver(b, max_iter)
    if verbose:
        print("\nStarting cg...")
    if x0 is None:
        # A(0) = 0 for a linear operator: r = -b without spending a matvec (reference cg.py:188 spends one)
        s.init(None, None, fused_minv, 0.0, tol, atol, martens_conv_crit, split)
    else:
        _lib.require_cuda(x0, "x0")
        x0 = _lib.vec(x0.detach().to(b.dtype))
        s.init(_lib.vec(A(x0).detach()), x0, fused_minv, 0.0, tol, atol, martens_conv_crit, split)
    if split:
        s.iterate(PCG_BETA, y_ext=_lib.vec(M(s.r).detach()))
    x_iters = [s.x.clone() if 0 in keep else None]
    if verbose:
        print(f"Residual norm required for termination: {s.status().res_bound:.6e}")
        print(f"Starting iterations (max_iter = {max_iter})...")

    it = 0
    while True:
        it += 1
        if verbose:
            print(f"  cg-iteration {it}")
        Ap = _lib.vec(A(s.p).detach())
        snap = torch.empty_like(b) if it in keep else None
        s.iterate(PCG_ALPHA if split else PCG_FUSED, Bp=Ap, minv=fused_minv, snapshot=snap)
        x_iters.append(snap)
        st = s.status()
        if not st.pAp > 0.0:
            _warn_nonpositive(st.pAp, it)
        if st.reason:
            break
        if split:
            s.iterate(PCG_BETA, y_ext=_lib.vec(M(s.r).detach()))

    reason = REASONS[st.reason]
    if verbose:
        print(reason)
    if x_iters[-1] is None:
        x_iters[-1] = s.x
    m_iters = list(s.m_iters(it + 1).unbind()) if martens_conv_crit else None
    return x_iters, m_iters, reason


def pcg_device(matvec, b, x0=None, minv=None, damping=0.0, max_iter=250, tol=1e-5, atol=None,
               martens_conv_crit=True, store_x_at_iters=None, poll=3, verbose=False, use_graph=False, out_buffer=None):
    """The same solve with a device-resident operator and **no host synchronisation per iteration**.

    ``matvec(v, out, skip_ptr)`` enqueues ``out = B v`` (the undamped curvature product) on the current
    stream; the damping ``lambda v`` is added inside the fused kernel (reference ``optimizer.py:266``) and
    ``minv`` (or None) is the hoisted diagonal preconditioner.  Termination is decided on the device
    (same tests, same order as ``cg.py:96-115``) and published as an ``{iter, reason}`` pair in pinned host memory
    (``hf_pcg_set_progress``).  The host enqueues iteration ``i`` once it knows that iteration ``i - poll`` did not
    terminate the solve -- a plain memory read, no copy, event or synchronisation -- so the queue holds ``poll - 1``
    iterations of work beyond the one executing and never drains.  The at most ``poll - 1`` launches enqueued after
    the solver stopped are no-ops (``skip_ptr``), and the iterate/iteration count reported are exactly the ones the
    reference would stop at.  The number of enqueued iterations, ``min(max_iter, n_stop + poll - 1)``, does not
    depend on timing, so data-parallel replicas issue the same collectives.  Returns ``(x_iters, m_iters, reason)``
    like :func:`cg`.

    With ``use_graph`` the body of one iteration (every kernel of the product + the fused update) is captured once
    into a CUDA graph after the first iteration and replayed, which removes the per-launch host cost; ``matvec``
    must then be capture-safe (no allocation, no host sync), which the native operators are.  Capture failures fall
    back to plain launches.  Off by default: capture + instantiation cost about as much as the launches they save
    on a 50-iteration solve (measured), so it only pays for long solves.
    """
    _lib.require_cuda(b, "b")
    b = _lib.vec(b.detach())
    max_iter = b.numel() if max_iter is None else int(max_iter)
    keep = set(cg_storing_grid(max_iter) if store_x_at_iters is None else store_x_at_iters)
    s = _Solver(b, max_iter)
    # `out_buffer`: where the operator wants its products written (a data-parallel problem hands out a view of its
    # symmetric vector, so that the per-iteration all-reduce can run through the NVSwitch in place)
    Bp = out_buffer if out_buffer is not None and out_buffer.shape == b.shape and out_buffer.dtype == b.dtype else torch.empty_like(b)
    if x0 is None:
        s.init(None, None, minv, damping, tol, atol, martens_conv_crit, False)
    else:
        x0 = _lib.vec(x0.detach().to(b.dtype))
        matvec(x0, Bp, None)
        s.init(Bp, x0, minv, damping, tol, atol, martens_conv_crit, False)
    x_first = s.x.clone() if 0 in keep else None
    slots = {it: k for k, it in enumerate(sorted(i for i in keep if 1 <= i <= max_iter))}
    row = (b.numel() + 3) // 4 * 4  # keep every snapshot row 16-byte aligned
    snaps = torch.empty((max(1, len(slots)), row), dtype=b.dtype, device=b.device)[:, : b.numel()]

    it, depth = 0, max(1, int(poll))
    graph = None

    def one_iteration(snapshot):
        matvec(s.p, Bp, s.reason_ptr)
        s.iterate(PCG_FUSED, Bp=Bp, minv=minv, lam=damping, snapshot=snapshot)

    while it < max_iter:
        if it + 1 - depth >= 1:
            done, reason = s.wait_for(it + 1 - depth)
            if reason != 0 and done <= it + 1 - depth:
                break
        it += 1
        snap = snaps[slots[it]] if it in slots else None
        if graph is None:
            one_iteration(snap)
            if use_graph and it == 1 and max_iter > 2:  # everything lazily initialised: capture the body once
                try:
                    cand, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):  # bare capture: no gc / empty_cache / device sync
                        cand.capture_begin()
                        one_iteration(None)
                        cand.capture_end()
                    torch.cuda.current_stream().wait_stream(side)
                    graph = cand
                except Exception:  # noqa: BLE001 -- any capture problem: keep launching directly
                    graph, use_graph = None, False
                    torch.cuda.synchronize()
        else:
            graph.replay()
            if snap is not None:
                snap.copy_(s.x)  # a terminated solve leaves x untouched, so a late copy is harmless
    st = s.status()
    if st.reason == 0:
        raise RuntimeError("pcg_device: solver did not terminate (internal error)")
    if st.nonpos_iter:
        _warn_nonpositive(st.nonpos_pAp, st.nonpos_iter)
    n = st.iter
    x_iters = [x_first] + [snaps[slots[i]] if i in slots else None for i in range(1, n + 1)]
    if x_iters[-1] is None:
        x_iters[-1] = s.x
    m_iters = list(s.m_iters(n + 1).unbind()) if martens_conv_crit else None
    if verbose:
        print(f"cg: {n} iterations, {REASONS[st.reason]}")
    return x_iters, m_iters, REASONS[st.reason]
