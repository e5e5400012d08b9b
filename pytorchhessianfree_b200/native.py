"""Python handles for the layer program (``hf_net``) and its linearisations (``hf_lin``).

A :class:`NativeNet` is the lowered form of the user's model + loss: the thing the reference keeps
implicitly as an autograd graph inside the closures at ``optimizer.py:241-247``.  A
:class:`Linearization` is that net evaluated on one chunk of data at one parameter point, with the
activations resident in HBM; every curvature product of the CG solve re-uses them (the reference
re-runs the forward pass per chunk per iteration, ``optimizer.py:805-814``).
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib
from ._lib import ACT, LIN_HESSIAN, LIN_LOSS_ONLY, LOSS, REDUCTION, LayerDesc


@dataclass
class LayerSpec:
    """One affine layer followed by an activation; offsets index the flat trainable vector."""

    in_features: int
    out_features: int
    act: str = "none"
    has_bias: bool = True
    w_offset: int = -1
    b_offset: int = -1
    w_frozen: Optional[torch.Tensor] = None
    b_frozen: Optional[torch.Tensor] = None
    # "conv2d": in_features = c_in*k_h*k_w and geom = (c_in, h_in, w_in, k_h, k_w, stride, pad, h_out, w_out);
    # "avgpool": average over the whole (h_in, w_in) map, geom = (c, h_in, w_in, 0, 0, 0, 0, 1, 1), no parameters
    kind: str = "linear"
    geom: tuple = ()

    def signature(self):
        return (self.in_features, self.out_features, self.act, self.has_bias, self.w_offset, self.b_offset,
                _lib.ptr(self.w_frozen), _lib.ptr(self.b_frozen), self.kind, tuple(self.geom))


ENGINES = {"simt": 0, "tc": 1}


class NativeNet:
    """Owner of an ``hf_net`` handle."""

    def __init__(self, layers: List[LayerSpec], loss: str, reduction: str, n_params: int, engine: str = "simt"):
        if loss not in LOSS:
            raise NotImplementedError(f"loss {loss!r} is not lowered; supported: {sorted(LOSS)}")
        if reduction not in REDUCTION:
            raise ValueError(f"Invalid reduction {reduction}")
        self.lib = _lib.load()
        self.layers, self.loss, self.reduction, self.n_params = layers, loss, reduction, int(n_params)
        self._keep = []  # frozen tensors must outlive the handle
        arr = (LayerDesc * len(layers))()
        for d, l in zip(arr, layers):
            d.in_features, d.out_features, d.act, d.has_bias = l.in_features, l.out_features, ACT[l.act], int(l.has_bias)
            d.w_offset, d.b_offset = l.w_offset, l.b_offset if l.has_bias else -1
            d.kind = _lib.LAYER_KIND[l.kind]
            if l.kind != "linear":
                (d.c_in, d.h_in, d.w_in, d.k_h, d.k_w, d.stride, d.pad, d.h_out, d.w_out) = l.geom
            for name, t in (("d_w_frozen", l.w_frozen), ("d_b_frozen", l.b_frozen)):
                if t is not None:
                    _lib.require_cuda(t, "frozen parameter")
                    t = t.detach().to(torch.float32).contiguous()
                    self._keep.append(t)
                    setattr(d, name, t.data_ptr())
        h = C.c_void_p()
        _lib.check(self.lib.hf_net_create(arr, len(layers), LOSS[loss], REDUCTION[reduction], self.n_params, C.byref(h)))
        self.handle = h
        self.engine = engine
        _lib.check(self.lib.hf_net_set_engine(self.handle, ENGINES[engine]))
        self.in_features, self.classes = layers[0].in_features, layers[-1].out_features
        # a convolution first: inputs are [batch, c_in, h_in, w_in] (NCHW, as the user's model takes them)
        self.input_shape = tuple(layers[0].geom[:3]) if layers[0].kind != "linear" else None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hf_net_destroy(h)

    def signature(self):
        return (tuple(l.signature() for l in self.layers), self.loss, self.reduction, self.n_params, self.engine)

    def first_layer_span(self):
        """(offset, count) of the first trainable layer's slice of the flat vector (count 0 if not contiguous)."""
        off, cnt = C.c_int64(), C.c_int64()
        _lib.check(self.lib.hf_net_first_layer_span(self.handle, C.byref(off), C.byref(cnt)))
        return off.value, cnt.value

    def linearize(self, x, targets, hessian=False, loss_only=False):
        return Linearization(self, x, targets, hessian=hessian, loss_only=loss_only)


class Linearization:
    """Owner of an ``hf_lin`` handle, its workspace, and the chunk's inputs/targets."""

    def __init__(self, net: NativeNet, x, targets, hessian=False, loss_only=False):
        _lib.require_cuda(x, "inputs")
        _lib.require_cuda(targets, "targets")
        self.net, self.lib = net, net.lib
        x = x.detach()
        if net.input_shape is not None:
            if tuple(x.shape[1:]) != net.input_shape:
                raise ValueError(f"inputs have shape {tuple(x.shape[1:])}, the first convolution takes {net.input_shape}")
        else:
            if x.dim() != 2:
                x = x.reshape(x.shape[0], -1)
            if x.shape[1] != net.in_features:
                raise ValueError(f"inputs have {x.shape[1]} features, the first layer takes {net.in_features}")
        self.x = x.to(torch.float32).contiguous()
        self.n = int(x.shape[0])
        t = targets.detach()
        if net.loss == "ce":
            if t.dim() != 1 or t.dtype.is_floating_point:
                raise NotImplementedError("softmax cross-entropy is lowered for class-index targets only")
            self.targets = t.to(torch.int64).contiguous()
        else:
            self.targets = t.to(torch.float32).reshape(self.n, -1).contiguous()
            if self.targets.shape[1] != net.classes:
                raise ValueError("targets and network outputs disagree in shape")
        if self.targets.shape[0] != self.n:
            raise ValueError("inputs and targets disagree in batch size")
        self.flags = (LIN_HESSIAN if hessian else 0) | (LIN_LOSS_ONLY if loss_only else 0)
        nbytes = self.lib.hf_lin_workspace_bytes(net.handle, self.n, self.flags)
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=x.device)
        base = (self.workspace.data_ptr() + 255) // 256 * 256
        h = C.c_void_p()
        _lib.check(self.lib.hf_lin_create(net.handle, self.n, self.flags, base, nbytes, C.byref(h)))
        self.handle = h

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.hf_lin_destroy(h)

    def forward(self, theta, n_total, loss_acc=None):
        """Forward pass at ``theta``; adds this chunk's loss share to the float64 scalar ``loss_acc``."""
        _lib.check(self.lib.hf_lin_forward(self.handle, _lib.ptr(theta), self.x.data_ptr(), self.targets.data_ptr(),
                                           int(n_total), _lib.ptr(loss_acc), _lib.stream()))

    def gradient(self, theta, out, accumulate=False):
        _lib.check(self.lib.hf_lin_gradient(self.handle, _lib.ptr(theta), out.data_ptr(), int(accumulate), _lib.stream()))

    def ggn(self, theta, v, out, accumulate=False, skip_ptr=None):
        _lib.check(self.lib.hf_ggn_matvec(self.handle, _lib.ptr(theta), v.data_ptr(), out.data_ptr(), int(accumulate),
                                          skip_ptr, _lib.stream()))

    def hessian(self, theta, v, out, accumulate=False, skip_ptr=None):
        _lib.check(self.lib.hf_hessian_matvec(self.handle, _lib.ptr(theta), v.data_ptr(), out.data_ptr(),
                                              int(accumulate), skip_ptr, _lib.stream()))

    def matvec_phase(self, kind, theta, v, out, phase, accumulate=False, skip_ptr=None):
        """Two-phase form of :meth:`ggn` / :meth:`hessian` (kind "ggn" | "hessian"); see ``hf_matvec_phase``."""
        _lib.check(self.lib.hf_matvec_phase(self.handle, int(kind == "hessian"), _lib.ptr(theta), v.data_ptr(),
                                            out.data_ptr(), int(accumulate), skip_ptr, _lib.stream(), int(phase)))

    def fisher(self, theta, out, accumulate=False):
        _lib.check(self.lib.hf_fisher_diag(self.handle, _lib.ptr(theta), out.data_ptr(), int(accumulate), _lib.stream()))

    def logits(self):
        """The network outputs of the last forward pass, as a tensor view (tests)."""
        p = self.lib.hf_lin_logits(self.handle)
        off = (p - self.workspace.data_ptr()) // 4
        ld = (self.net.classes + 3) // 4 * 4  # library-owned [batch, width] buffers are pitched to 16 bytes
        flat = self.workspace.view(torch.float32)
        return flat[off: off + self.n * ld].view(self.n, ld)[:, : self.net.classes]
