"""Stand-in for ``backpack-for-pytorch`` (absent here: no wheel, no network)  --  TEST INFRASTRUCTURE ONLY.

Lets the UNMODIFIED reference ``/root/reference/hessianfree`` import and run in
this container so that golden vectors can be minted from the reference's own
code path (``tests/golden/make_golden.py``).  The four ``backpack.hessianfree``
helpers are BackPACK's published autograd recipes, restated in
``oracle/hf_oracle.py``; ``SumGradSquared`` is provided for ``nn.Linear`` only
(sum over the batch of squared per-sample gradients, with whatever 1/N factor
the loss reduction put into the back-propagated signal -- the convention the
reference corrects for at ``preconditioners.py:56-58``).
"""
import contextlib

import torch

_ACTIVE = []


def extend(module, **_):
    if getattr(module, "_shim_extended", False):
        return module
    for sub in module.modules():
        if isinstance(sub, torch.nn.Linear) and not getattr(sub, "_shim_hooked", False):
            sub.register_forward_hook(_remember_input)
            sub.register_full_backward_hook(_sum_grad_squared)
            sub._shim_hooked = True
    module._shim_extended = True
    return module


def _remember_input(mod, args, out):
    mod._shim_in = args[0].detach()


def _sum_grad_squared(mod, grad_in, grad_out):
    if not _ACTIVE:
        return
    go = grad_out[0].detach()
    a = mod._shim_in
    go2 = go.reshape(-1, go.shape[-1]) ** 2 if go.dim() > 1 else go[None] ** 2
    a2 = a.reshape(-1, a.shape[-1]) ** 2 if a.dim() > 1 else a[None] ** 2
    if mod.weight.requires_grad:
        mod.weight.sum_grad_squared = go2.T @ a2
    if mod.bias is not None and mod.bias.requires_grad:
        mod.bias.sum_grad_squared = go2.sum(0)


@contextlib.contextmanager
def backpack(*exts, **_):
    _ACTIVE.append(exts)
    try:
        yield
    finally:
        _ACTIVE.pop()
