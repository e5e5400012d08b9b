"""Empirical-Fisher diagonal and the diagonal preconditioner (Martens 2010, sec. 4.7; reference
``hessianfree/preconditioners.py``), on the sm_100a kernels.

``diag_EF_backpack`` and ``diag_EF_autograd`` keep the reference's names and results
(``sum_n g_n^2`` for "sum", ``(1/N) sum_n g_n^2`` for "mean"); both run the same device contraction
``(delta^2)^T (a^2)`` per Linear layer instead of BackPACK's ``SumGradSquared`` or a per-sample loop.
"""
import torch

from . import _lib
from .cg import DiagonalPreconditioner
from .lowering import lower_module
from .native import NativeNet


def _diag_EF(model, loss_function, inputs, targets, reduction, engine="simt"):
    if reduction not in ["sum", "mean"]:
        raise ValueError(f"reduction {reduction} is not supported.")
    params = [p for p in model.parameters() if p.requires_grad]
    prog = lower_module(model, loss_function, params, input_shape=tuple(inputs.shape[1:]) if inputs.dim() == 4 else None)
    if prog.reduction != reduction:
        raise ValueError(f"the loss function reduces by {prog.reduction!r} but reduction={reduction!r} was given")
    _lib.require_cuda(inputs, "inputs")
    theta = torch.cat([p.detach().reshape(-1) for p in params]).to(torch.float32)
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    lin = net.linearize(inputs, targets)
    lin.forward(theta, lin.n, None)
    out = torch.empty_like(theta)
    lin.fisher(theta, out)
    return out


def diag_EF_backpack(model, loss_function, inputs, targets, reduction):
    """Diagonal of the empirical Fisher (reference ``preconditioners.py:11-60``)."""
    return _diag_EF(model, loss_function, inputs, targets, reduction)


def diag_EF_autograd(model, loss_function, inputs, targets, reduction):
    """Same quantity; the reference's slow per-sample variant (``preconditioners.py:63-105``) needs no
    separate implementation here."""
    return _diag_EF(model, loss_function, inputs, targets, reduction)


def diag_to_preconditioner(diag_vec, damping, exponent=0.75):
    """``x -> (diag_vec + damping)^(-exponent) * x`` (reference ``preconditioners.py:108-127``).  The
    returned object is callable and is recognised by :func:`pytorchhessianfree_b200.cg.cg`."""
    return DiagonalPreconditioner(diag_vec, damping, exponent)


def diag_EF_preconditioner(model, loss_function, inputs, targets, reduction, damping, exponent=None,
                           use_backpack=True):
    """Fisher diagonal -> preconditioner (reference ``preconditioners.py:130-159``); ``use_backpack`` is
    accepted for signature compatibility and has no effect."""
    d = _diag_EF(model, loss_function, inputs, targets, reduction)
    return diag_to_preconditioner(d, damping) if exponent is None else diag_to_preconditioner(d, damping, exponent)
