"""CPU oracle for the Hessian-free inner solve  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, autograd) restatement of the algorithm of
ltatzel/PyTorchHessianFree on the one hot path this repository accelerates:
the preconditioned CG Newton-step solve and the curvature-matrix-vector
products it calls.  It exists so that the CUDA path can be checked against
something that (a) travels to the GPU box (``/root/reference`` does not) and
(b) was itself pinned against the unmodified reference: see
``tests/golden/make_golden.py`` (runs the reference from ``/root/reference``
with ``oracle/backpack_shim`` on the path, asserts this oracle reproduces it,
and writes the fixtures under ``tests/golden/``) and
``tests/test_oracle_golden.py`` (re-checks the oracle against those fixtures
on every CPU test run).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.  The
product package ``pytorchhessianfree_b200`` never does.

Third-party arithmetic: the reference delegates the matvecs to
``backpack-for-pytorch>=1.5.0,<2.0.0`` (reference ``setup.py:16``; no lock
file, source not vendored).  BackPACK's ``hessianfree`` helpers are published
as thin autograd recipes (R-op by the double-backward trick, L-op = vjp,
``hvp = R_op(grad)``, ``ggnvp = L_op(H_loss * R_op)``); they are restated in
``rop/lop/hvp/ggnvp`` below and anchored on the reference's own call sites
(``hessianfree/optimizer.py:450-462``) and tests (see the golden script).

Every function cites the reference file:line it follows.
"""

from __future__ import annotations

import math
import warnings
from typing import Callable, List, Optional, Sequence

import torch

# ----------------------------------------------------------------------------
# flat-vector <-> parameter-list plumbing (reference hessianfree/utils.py)
# ----------------------------------------------------------------------------


def flatten(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """``torch.nn.utils.parameters_to_vector`` (used at optimizer.py:234,455,462)."""
    return torch.cat([t.reshape(-1) for t in tensors])


def unflatten(vec: torch.Tensor, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Views of ``vec`` shaped like ``like`` (reference utils.py:41-76)."""
    out, off = [], 0
    for t in like:
        n = t.numel()
        out.append(vec[off : off + n].view_as(t))
        off += n
    if off != vec.numel():
        warnings.warn("Not all entries of `vec` have been used.")
    return out


def load_trainable(vec: torch.Tensor, params: Sequence[torch.Tensor]) -> None:
    """Write ``vec`` into the trainable parameters only (reference utils.py:8-38)."""
    off = 0
    for p in params:
        if p.requires_grad:
            n = p.numel()
            p.data = vec[off : off + n].view_as(p).data
            off += n
    if off != vec.numel():
        warnings.warn("Not all entries of `vec` have been used.")


# ----------------------------------------------------------------------------
# BackPACK hessianfree helpers, restated (call sites optimizer.py:454,461)
# ----------------------------------------------------------------------------


def lop(ys, xs, ws):
    """Vector-Jacobian product  J^T w  (BackPACK ``L_op``)."""
    return torch.autograd.grad(ys, xs, grad_outputs=ws, create_graph=True, retain_graph=True, allow_unused=True)


def rop(ys, xs, vs):
    """Jacobian-vector product  J v  by the double-backward trick (BackPACK ``R_op``).

    g(w) = J^T w is linear in the dummy w, so d<g(w), v>/dw = J v.
    """
    single = isinstance(ys, torch.Tensor)
    ys_t = (ys,) if single else tuple(ys)
    ws = [torch.zeros_like(y, requires_grad=True) for y in ys_t]
    gs = torch.autograd.grad(ys_t, xs, grad_outputs=ws, create_graph=True, retain_graph=True, allow_unused=True)
    # parameters the outputs do not depend on contribute nothing
    pairs = [(g, v) for g, v in zip(gs, vs) if g is not None]
    re = torch.autograd.grad([g for g, _ in pairs], ws, grad_outputs=[v for _, v in pairs], create_graph=True, retain_graph=True, allow_unused=True)
    re = tuple(torch.zeros_like(y) if r is None else r for r, y in zip(re, ys_t))
    return re


def hvp(f, xs, vs, grad_xs=None):
    """Hessian-vector product  (d^2 f / dx^2) v  (BackPACK ``hessian_vector_product``)."""
    if grad_xs is None:
        grad_xs = torch.autograd.grad(f, xs, create_graph=True, retain_graph=True)
    return rop(grad_xs, xs, vs)


def ggnvp(loss, outputs, plist, vlist):
    """GGN-vector product  J^T H_loss J v  (BackPACK ``ggn_vector_product_from_plist``)."""
    Jv = rop(outputs, plist, vlist)
    HJv = hvp(loss, outputs, Jv)
    return lop(outputs, plist, HJv)


def Gv(loss, outputs, params, vec):
    """Flat-in / flat-out GGN product (reference optimizer.py:457-462)."""
    res = ggnvp(loss, outputs, params, unflatten(vec, params))
    return flatten([torch.zeros_like(p) if r is None else r for r, p in zip(res, params)]).detach()


def Hv(loss, params, vec):
    """Flat-in / flat-out Hessian product (reference optimizer.py:450-455)."""
    res = hvp(loss, params, unflatten(vec, params))
    return flatten(res).detach()


# ----------------------------------------------------------------------------
# Preconditioned CG (reference hessianfree/cg.py)
# ----------------------------------------------------------------------------

REASON_MARTENS = "Convergence (Martens)"
REASON_MAXITER = "Number of iterations"
REASON_DIVERGED = "Divergence"
REASON_TOL = "Convergence (tolerances)"


def storing_grid(max_iter: int, gamma: float = 1.3) -> List[int]:
    """Iterations ceil(gamma^j)-1 at which CG keeps a snapshot (cg.py:152-170).

    The reference evaluates gamma**j on an int64 ``arange`` in float32; the
    arithmetic is reproduced with torch so that the float32 rounding matches.
    """
    if gamma < 1.0:
        raise ValueError(f"Invalid gamma = {gamma}")
    j_hi = math.ceil(math.log(max_iter + 1) / math.log(gamma))
    pw = gamma ** torch.arange(j_hi + 1)
    return sorted({int(v) for v in (torch.ceil(pw) - 1).int().tolist()})


def pcg(
    A: Callable,
    b: torch.Tensor,
    x0: Optional[torch.Tensor] = None,
    M: Optional[Callable] = None,
    max_iter: Optional[int] = None,
    tol: float = 1e-5,
    atol: Optional[float] = None,
    martens_conv_crit: bool = False,
    store_x_at_iters=(),
):
    """Martens' Algorithm-2 form of PCG; returns (x_iters, m_iters, reason).

    Conventions follow cg.py:186-231: residual r = A x - b, direction p = -y,
    m_i = 0.5 (r-b)^T x; termination tests in the order Martens, max_iter,
    NaN, tolerance (cg.py:96-115).
    """
    bound = tol * torch.linalg.norm(b).item()  # cg.py:75
    if atol is not None:
        bound = max(bound, atol)  # cg.py:76
    if max_iter is None:
        max_iter = b.numel()  # cg.py:177
    if x0 is None:
        x0 = torch.zeros_like(b)  # cg.py:178
    keep = set(storing_grid(max_iter) if store_x_at_iters is None else store_x_at_iters)

    x = x0
    xs = [x if 0 in keep else None]
    r = A(x0) - b  # cg.py:188
    ms = [0.5 * torch.dot(r - b, x0)] if martens_conv_crit else None
    y = r if M is None else M(r)
    ry = torch.dot(r, y)
    p = -y

    k = 0
    while True:
        k += 1
        Ap = A(p).detach()
        pAp = torch.dot(p, Ap)
        if not pAp > 0:  # cg.py:133-143, option "ignore": warn, keep the value
            warnings.warn(
                f"Directional curvature pAp = {pAp:.3e} <= 0 detected in cg-iteration {k}. "
                "This is a violation to the assumption of positive definiteness."
            )
        step = ry / pAp
        x = x + step * p
        stored = k in keep
        xs.append(x if stored else None)
        r = r + step * Ap

        # termination, cg.py:93-118
        rn = torch.linalg.norm(r)
        reason = None
        if martens_conv_crit:
            ms.append(0.5 * torch.dot(r - b, x))
            w = max(10, int(k / 10))
            if w < k and (ms[k] - ms[k - w]) / (ms[k] - ms[0]) < 5e-4:
                reason = REASON_MARTENS
        if reason is None:
            if k >= max_iter:
                reason = REASON_MAXITER
            elif torch.isnan(rn):
                reason = REASON_DIVERGED
            elif rn < bound:
                reason = REASON_TOL
        if reason is not None:
            break

        y = r if M is None else M(r)
        ry_next = torch.dot(r, y)
        p = -y + (ry_next / ry) * p
        ry = ry_next

    if not stored:
        xs[-1] = x  # cg.py:229-230
    return xs, ms, reason


# ----------------------------------------------------------------------------
# Empirical-Fisher diagonal and the diagonal preconditioner
# (reference hessianfree/preconditioners.py)
# ----------------------------------------------------------------------------


def ef_diag(model, loss_fn, inputs, targets, reduction: str) -> torch.Tensor:
    """sum_n g_n^2 ("sum") or (1/N) sum_n g_n^2 ("mean"), per-sample loop
    (preconditioners.py:63-105, which the reference's own test uses as the
    check for the BackPACK variant, tests/test_preconditioners.py:79-99)."""
    if reduction not in ("sum", "mean"):
        raise ValueError(f"reduction {reduction} is not supported.")
    params = [p for p in model.parameters() if p.requires_grad]
    acc = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype)
    for xi, ti in zip(inputs, targets):
        li = loss_fn(model(xi), ti)
        acc += flatten(torch.autograd.grad(li, params)) ** 2
    return acc / inputs.shape[0] if reduction == "mean" else acc


def ef_diag_layerwise(model, loss_fn, inputs, targets, reduction: str) -> torch.Tensor:
    """The same diagonal the way BackPACK's ``SumGradSquared`` forms it for ``nn.Linear`` (what
    ``diag_EF_backpack`` reads, preconditioners.py:43-58): with ``d = d loss / d z_l`` (rows are per-sample
    because the samples are independent) the per-sample gradients are ``d[n]^T a[n]`` and ``d[n]``, so
    ``sum_n g_n^2 = (d^2)^T (a^2)`` for the weight and ``sum_n d^2`` for the bias; "mean" multiplies by N
    because the loss already scaled every per-sample gradient by 1/N (preconditioners.py:56-58).  One
    backward pass instead of N: usable as the oracle at benchmark batch sizes.  Pinned against the
    per-sample loop ``ef_diag`` in tests/test_oracle_golden.py."""
    if reduction not in ("sum", "mean"):
        raise ValueError(f"reduction {reduction} is not supported.")
    linears = [m for m in model.modules() if isinstance(m, torch.nn.Linear) and any(p.requires_grad for p in m.parameters())]
    acts_in, outs = {}, {}
    def remember(mod, args, out):  # returns None: the output is not replaced
        acts_in[mod], outs[mod] = args[0].detach(), out

    hooks = [m.register_forward_hook(remember) for m in linears]
    try:
        loss = loss_fn(model(inputs), targets)
    finally:
        for h in hooks:
            h.remove()
    deltas = torch.autograd.grad(loss, [outs[m] for m in linears])
    by_param = {}
    for m, d in zip(linears, deltas):
        if m.weight.requires_grad:
            by_param[m.weight] = (d * d).t() @ (acts_in[m] * acts_in[m])
        if m.bias is not None and m.bias.requires_grad:
            by_param[m.bias] = (d * d).sum(0)
    params = [p for p in model.parameters() if p.requires_grad]
    diag = flatten([by_param[p] for p in params])
    return diag * inputs.shape[0] if reduction == "mean" else diag


def diag_precond(diag: torch.Tensor, damping: float, exponent: float = 0.75) -> Callable:
    """x -> (diag + damping)^(-exponent) * x   (preconditioners.py:108-127)."""
    return lambda x: torch.mul((diag + damping) ** -exponent, x)


# ----------------------------------------------------------------------------
# Step selection after the solve
# ----------------------------------------------------------------------------


def backtrack_all(f, steps):
    """Exhaustive snapshot search (cg_backtracking.py:6-50)."""
    vals = [float("inf") if s is None else f(s) for s in steps]
    best = int(torch.argmin(torch.tensor(vals)))
    return best, vals[best]


def backtrack_efficient(f, steps):
    """Walk back from the last snapshot while the loss improves
    (cg_backtracking.py:53-112)."""
    best, best_val = None, float("inf")
    for i in range(len(steps) - 1, -1, -1):
        if steps[i] is None:
            continue
        v = f(steps[i])
        if v < best_val:
            best, best_val = i, v
        else:
            break
    return best, best_val


def armijo(f, grad0, step, init_alpha=1.0, beta=0.8, c=1e-2, max_iter=20):
    """Back-off line search (linesearch.py:8-103); (0.0, f0) on failure."""
    if beta >= 1.0:
        raise ValueError(f"Invalid reduction factor beta = {beta}")
    if c < 0.0:
        raise ValueError(f"Invalid c = {c}")
    f0 = float(f(torch.zeros_like(step)))
    alpha = init_alpha
    fa = float(f(alpha * step))
    slope = c * torch.dot(grad0, step).item()
    if slope >= 0:
        warnings.warn(
            "`update_vec`-parameter in `simple_linesearch` is not a descent "
            f"direction. The directional derivative is {slope:.6f}."
        )
    for _ in range(max_iter):
        if float(fa) <= f0 + alpha * slope:
            return alpha, fa
        alpha *= beta
        fa = f(alpha * step)
    warnings.warn("No suitable update could be found by the line search.")
    return 0.0, f0


# ----------------------------------------------------------------------------
# The optimizer step (reference hessianfree/optimizer.py:126-363, 519-814)
# ----------------------------------------------------------------------------


class OracleHF:
    """Restatement of ``HessianFree.step`` / ``acc_step`` on CPU autograd.

    Not a ``torch.optim.Optimizer``: just enough state (damping, x0, logs) to
    reproduce the reference trajectory for parity tests.
    """

    def __init__(self, params, curvature_opt="ggn", damping=1.0, adapt_damping=True, cg_max_iter=250,
                 cg_decay_x0=0.95, use_cg_backtracking=True, lr=1.0, use_linesearch=True):
        self.all_params = list(params)
        self.params = [p for p in self.all_params if p.requires_grad]  # optimizer.py:122
        self.curvature_opt = curvature_opt
        self.damping = damping
        self.adapt_damping = adapt_damping and damping != 0.0  # optimizer.py:88-90
        self.cg_max_iter = cg_max_iter
        self.cg_decay_x0 = cg_decay_x0
        self.use_cg_backtracking = use_cg_backtracking
        self.lr = lr
        self.use_linesearch = use_linesearch
        self.x0 = None
        self.log = {k: [] for k in ("init_losses", "dampings", "cg_reasons", "num_cg_iters", "best_cg_iters",
                                    "learning_rates", "final_losses", "m_iters")}

    # -- step ---------------------------------------------------------------
    def step(self, forward, grad=None, mvp=None, M_func=None):
        ctx = torch.no_grad() if (grad is not None and mvp is not None) else torch.enable_grad()
        with ctx:
            loss, outputs = forward()  # optimizer.py:218-223
        self.log["init_losses"].append(loss.item())
        if grad is None:
            g = torch.autograd.grad(loss, self.params, create_graph=True, retain_graph=True)
            grad = flatten(g).detach()  # optimizer.py:230-234
        if mvp is None:
            if self.curvature_opt == "hessian":
                mvp = lambda v: Hv(loss, self.params, v)  # noqa: E731
            else:
                mvp = lambda v: Gv(loss, outputs, self.params, v)  # noqa: E731
        lam = self.damping
        self.log["dampings"].append(lam)
        xs, ms, reason = pcg(
            lambda v: mvp(v) + lam * v,  # optimizer.py:266
            -grad, x0=self.x0, M=M_func, max_iter=self.cg_max_iter, martens_conv_crit=True,
            store_x_at_iters=None if self.use_cg_backtracking else [0],
        )
        self.log["cg_reasons"].append(reason)
        self.log["num_cg_iters"].append(len(xs) - 1)
        self.log["m_iters"].append([float(m) for m in ms])
        self.x0 = self.cg_decay_x0 * xs[-1]  # optimizer.py:281
        base = flatten(self.params).detach()

        @torch.no_grad()
        def tfunc(s):  # optimizer.py:290-294
            load_trainable(base + s, self.all_params)
            return forward()[0].item()

        if self.adapt_damping:  # optimizer.py:300-306, 464-506
            f_start = tfunc(xs[0])  # loss at the CG start point, not at the current parameters
            f_end = tfunc(xs[-1])
            rho = (f_end - f_start) / (ms[-1] - ms[0])
            if rho < 0.25:
                self.damping *= 3 / 2
            elif rho > 0.75:
                self.damping *= 2 / 3
            if rho < 0:
                warnings.warn("The reduction ratio `rho` is negative. This might result in a bad "
                              "cg-initialization in the next step.")
        step_vec = xs[-1]
        if self.use_cg_backtracking:  # optimizer.py:311-318
            best, _ = backtrack_efficient(tfunc, xs)
            self.log["best_cg_iters"].append(best)
            step_vec = xs[best]
        lr, final = self.lr, None
        if self.use_linesearch:  # optimizer.py:333-339
            lr, final = armijo(tfunc, grad, step_vec, init_alpha=lr)
        self.log["learning_rates"].append(lr)
        load_trainable(base + lr * step_vec, self.all_params)  # optimizer.py:349-350
        self.log["final_losses"].append(final)
        return final

    # -- acc_step -------------------------------------------------------------
    @staticmethod
    def accumulate(model, loss_fn, datalist, with_grad, init, per_chunk, reduction):
        """N-weighted chunk accumulation (optimizer.py:608-684)."""
        if reduction not in ("mean", "sum"):
            raise ValueError(f"Invalid reduction {reduction}")
        total, n_all = init, 0
        for X, T in datalist:
            n = T.shape[0]
            n_all += n
            with torch.enable_grad() if with_grad else torch.no_grad():
                out = model(X)
                loss = loss_fn(out, T)
            q = per_chunk(loss, out)
            total = total + (n * q if reduction == "mean" else q)
        return total / n_all if reduction == "mean" else total

    def acc_loss(self, model, loss_fn, datalist, reduction):
        return self.accumulate(model, loss_fn, datalist, False, 0.0, lambda l, o: l.detach(), reduction)

    def acc_grad(self, model, loss_fn, datalist, reduction):
        z = torch.zeros_like(flatten(self.params))
        return self.accumulate(model, loss_fn, datalist, True, z,
                               lambda l, o: flatten(torch.autograd.grad(l, self.params)).detach(), reduction)

    def acc_mvp(self, model, loss_fn, datalist, reduction, v):
        z = torch.zeros_like(flatten(self.params))
        if self.curvature_opt == "hessian":
            f = lambda l, o: Hv(l, self.params, v)  # noqa: E731
        else:
            f = lambda l, o: Gv(l, o, self.params, v)  # noqa: E731
        return self.accumulate(model, loss_fn, datalist, True, z, f, reduction)

    def acc_step(self, model, loss_fn, loss_datalist, grad_datalist=None, mvp_datalist=None, M_func=None,
                 reduction="mean"):
        grad_datalist = loss_datalist if grad_datalist is None else grad_datalist
        mvp_datalist = loss_datalist if mvp_datalist is None else mvp_datalist
        return self.step(
            forward=lambda: (self.acc_loss(model, loss_fn, loss_datalist, reduction), None),
            grad=self.acc_grad(model, loss_fn, grad_datalist, reduction),
            mvp=lambda v: self.acc_mvp(model, loss_fn, mvp_datalist, reduction, v),
            M_func=M_func,
        )


# ----------------------------------------------------------------------------
# Explicit-matrix known answers (not in the reference)
# ----------------------------------------------------------------------------


def explicit_ggn(model, loss_fn, inputs, targets) -> torch.Tensor:
    """Dense  J^T H_loss J  over the trainable parameters of a tiny net (float64 advised)."""
    params = [p for p in model.parameters() if p.requires_grad]
    sizes = [p.numel() for p in params]

    def out_of(flat):
        parts = unflatten(flat, params)
        names = [n for n, p in model.named_parameters() if p.requires_grad]
        return torch.func.functional_call(model, dict(zip(names, parts)), (inputs,)).reshape(-1)

    theta = flatten([p.detach() for p in params])
    J = torch.autograd.functional.jacobian(out_of, theta)  # [N*C, P]
    z = model(inputs).detach()
    H = torch.autograd.functional.hessian(lambda zz: loss_fn(zz.view_as(z), targets), z.reshape(-1))
    assert J.shape[1] == sum(sizes)
    return J.T @ H @ J


def explicit_hessian(model, loss_fn, inputs, targets) -> torch.Tensor:
    """Dense Hessian of the loss over the trainable parameters of a tiny net."""
    params = [p for p in model.parameters() if p.requires_grad]
    names = [n for n, p in model.named_parameters() if p.requires_grad]

    def loss_of(flat):
        parts = unflatten(flat, params)
        return loss_fn(torch.func.functional_call(model, dict(zip(names, parts)), (inputs,)), targets)

    theta = flatten([p.detach() for p in params])
    return torch.autograd.functional.hessian(loss_of, theta)
