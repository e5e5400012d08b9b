"""Flat-vector <-> parameter plumbing (reference ``hessianfree/utils.py``).

The flat layout is the ABI of every vector on the hot path: the trainable parameters in
``param_groups[0]["params"]`` order, each flattened row-major.
"""
from warnings import warn

import torch


def _check_vec(vec):
    if not isinstance(vec, torch.Tensor):
        raise TypeError(f"`vec` should be a torch.Tensor, not {type(vec)}.")


def vector_to_trainparams(vec, parameters):
    """Point every *trainable* parameter at its slice of ``vec`` (reference ``utils.py:8-38``);
    parameters with ``requires_grad == False`` are skipped and consume no entries."""
    _check_vec(vec)
    used = 0
    for prm in parameters:
        if not prm.requires_grad:
            continue
        n = prm.numel()
        prm.data = vec[used: used + n].view_as(prm).data
        used += n
    if used != len(vec):
        warn("Not all entries of `vec` have been used.")


def vector_to_parameter_list(vec, parameters):
    """Views of ``vec`` shaped like ``parameters``, which are left untouched (reference ``utils.py:41-76``)."""
    _check_vec(vec)
    views, used = [], 0
    for prm in parameters:
        n = prm.numel()
        views.append(vec[used: used + n].view_as(prm).data)
        used += n
    if used != len(vec):
        warn("Not all entries of `vec` have been used.")
    return views
