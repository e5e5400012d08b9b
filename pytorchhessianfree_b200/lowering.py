"""Lower a PyTorch model + loss to the layer program the sm_100a kernels execute.

Two front ends, one result (:class:`Program`):

* :func:`lower_module` for the entry points that receive the model explicitly
  (``acc_step``, ``get_preconditioner``, ``test_reduction``; reference
  ``optimizer.py:519-529, 817, 928-937``);
* :func:`lower_graph` for ``step(forward)``, which only receives an opaque closure
  (``optimizer.py:137-151``): the structure is recovered from the autograd graph hanging off
  ``loss`` -- exactly the object the reference hands to BackPACK (``optimizer.py:241-247``).

Anything that cannot be lowered raises ``NotImplementedError`` naming the offending op.  There is no
autograd or CPU fallback inside this package; a caller with an exotic model can still pass its own
``mvp`` to ``step`` (``optimizer.py:160-164``), and the solver around it stays native.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
from torch import nn

from .native import LayerSpec

_ACT_MODULES = {nn.ReLU: "relu", nn.Sigmoid: "sigmoid", nn.Tanh: "tanh"}
_ACT_NODES = {"ReluBackward0": "relu", "SigmoidBackward0": "sigmoid", "TanhBackward0": "tanh"}
_REDUCTION_ENUM = {1: "mean", 2: "sum"}  # at::Reduction


@dataclass
class Program:
    layers: List[LayerSpec]
    loss: str
    reduction: str
    n_params: int
    inputs: Optional[torch.Tensor] = None   # graph front end only: the constant feeding the first lowered layer
    targets: Optional[torch.Tensor] = None
    param_slices: Dict[int, int] = field(default_factory=dict)


def flat_offsets(params_list):
    """id(param) -> offset in the flat trainable vector (reference ``optimizer.py:122`` order)."""
    table, off = {}, 0
    for p in params_list:
        table[id(p)] = off
        off += p.numel()
    return table, off


def _unsupported(what):
    raise NotImplementedError(
        f"{what} cannot be lowered to the sm_100a layer program (supported: Linear, Conv2d, ReLU, Sigmoid, Tanh, a "
        "global AvgPool2d / AdaptiveAvgPool2d(1), Flatten, Identity, eval-mode Dropout; MSELoss, CrossEntropyLoss, "
        "BCEWithLogitsLoss). Pass your own `mvp` to `step`, or restructure the model."
    )


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def lower_loss(loss_func):
    """(kind, reduction) of a supported loss module."""
    if isinstance(loss_func, nn.MSELoss):
        kind = "mse"
    elif isinstance(loss_func, nn.CrossEntropyLoss):
        if loss_func.weight is not None or loss_func.label_smoothing != 0.0:
            _unsupported("CrossEntropyLoss with class weights or label smoothing")
        kind = "ce"
    elif isinstance(loss_func, nn.BCEWithLogitsLoss):
        if loss_func.weight is not None or loss_func.pos_weight is not None:
            _unsupported("BCEWithLogitsLoss with weights")
        kind = "bce"
    else:
        _unsupported(f"loss {type(loss_func).__name__}")
    if loss_func.reduction not in ("mean", "sum"):
        _unsupported(f"loss reduction {loss_func.reduction!r}")
    return kind, loss_func.reduction


def _leaves(module):
    # `_modules`, not `children()`: the latter drops repeated instances, and a Sequential that applies one
    # `nn.ReLU()` object after every Linear would silently lose all but the first activation
    kids = [k for k in module._modules.values() if k is not None]
    if not kids:
        yield module
    else:
        if not isinstance(module, nn.Sequential):
            _unsupported(f"container {type(module).__name__} (only nn.Sequential nesting is walked)")
        for k in kids:
            yield from _leaves(k)


def _check_param_use(layers, offsets):
    """Every trainable parameter must feed exactly one layer: the transposed sweep writes each layer's slice of the
    flat vector once, so a shared weight would have one use overwrite the other."""
    used = [l.w_offset for l in layers if l.w_offset >= 0] + [l.b_offset for l in layers if l.has_bias and l.b_offset >= 0]
    if len(used) != len(set(used)):
        _unsupported("a parameter that is used by more than one layer (weight sharing)")
    return set(used)


def lower_module(model, loss_func, params_list, input_shape=None):
    """Walk a (nested) ``nn.Sequential`` of Linear / Conv2d / activation / pooling layers.

    ``input_shape`` = (channels, height, width) of one sample is needed when the model starts with a convolution (the
    entry points that call this pass the shape of the data they were given)."""
    offsets, n_params = flat_offsets(params_list)
    kind, reduction = lower_loss(loss_func)
    layers: List[LayerSpec] = []
    fmap = tuple(input_shape) if input_shape is not None and len(input_shape) == 3 else None  # (c, h, w) of the running feature map
    for m in _leaves(model):
        if isinstance(m, nn.Conv2d):
            if fmap is None:
                _unsupported("a Conv2d whose input feature-map size is unknown (inputs must be [batch, c, h, w])" if not layers
                             else "a Conv2d after a fully connected layer")
            if (m.groups != 1 or _pair(m.dilation) != (1, 1) or m.padding_mode != "zeros" or isinstance(m.padding, str)
                    or len(set(_pair(m.stride))) != 1 or len(set(_pair(m.padding))) != 1):
                _unsupported("a Conv2d with groups, dilation, string/non-zero-mode padding or anisotropic stride/padding")
            c, h, w_ = fmap
            if c != m.in_channels:
                raise ValueError(f"Conv2d expects {m.in_channels} input channels, the feature map has {c}")
            kh, kw = _pair(m.kernel_size)
            st, pd = _pair(m.stride)[0], _pair(m.padding)[0]
            ho, wo = (h + 2 * pd - kh) // st + 1, (w_ + 2 * pd - kw) // st + 1
            w, b = m.weight, m.bias
            spec = LayerSpec(c * kh * kw, m.out_channels, "none", b is not None, kind="conv2d",
                             geom=(c, h, w_, kh, kw, st, pd, ho, wo))
            fmap = (m.out_channels, ho, wo)
            for prm, off_name, frozen_name in ((w, "w_offset", "w_frozen"), (b, "b_offset", "b_frozen")):
                if prm is None:
                    continue
                if prm.requires_grad:
                    if id(prm) not in offsets:
                        raise ValueError("a trainable model parameter is not among the optimizer's parameters")
                    setattr(spec, off_name, offsets[id(prm)])
                else:
                    setattr(spec, frozen_name, prm.detach().reshape(m.out_channels, -1))
            layers.append(spec)
        elif isinstance(m, (nn.AvgPool2d, nn.AdaptiveAvgPool2d)):
            if fmap is None:
                _unsupported("a pooling layer outside a convolutional stack")
            c, h, w_ = fmap
            if isinstance(m, nn.AdaptiveAvgPool2d):
                whole = _pair(m.output_size) == (1, 1)
            else:
                whole = (_pair(m.kernel_size) == (h, w_) and _pair(m.padding) == (0, 0)
                         and (m.stride is None or _pair(m.stride) == (h, w_) or (h, w_) == _pair(m.kernel_size)))
            if not whole:
                _unsupported("an average pool that does not cover the whole feature map")
            layers.append(LayerSpec(c, c, "none", False, kind="avgpool", geom=(c, h, w_, 0, 0, 0, 0, 1, 1)))
            fmap = (c, 1, 1)
        elif isinstance(m, nn.Linear):
            if fmap is not None and fmap[1:] != (1, 1):
                _unsupported("a Linear layer on a feature map larger than 1x1 (pool first)")
            fmap = None
            w, b = m.weight, m.bias
            spec = LayerSpec(m.in_features, m.out_features, "none", b is not None)
            for prm, off_name, frozen_name in ((w, "w_offset", "w_frozen"), (b, "b_offset", "b_frozen")):
                if prm is None:
                    continue
                if prm.requires_grad:
                    if id(prm) not in offsets:
                        raise ValueError("a trainable model parameter is not among the optimizer's parameters")
                    setattr(spec, off_name, offsets[id(prm)])
                else:
                    setattr(spec, frozen_name, prm.detach())
            layers.append(spec)
        elif type(m) in _ACT_MODULES:
            if not layers or layers[-1].act != "none" or layers[-1].kind == "avgpool":
                _unsupported("an activation that does not directly follow a Linear or Conv2d layer")
            layers[-1].act = _ACT_MODULES[type(m)]
        elif isinstance(m, (nn.Identity, nn.Flatten)):
            if isinstance(m, nn.Flatten) and not layers and fmap is not None and (m.start_dim, m.end_dim) == (1, -1):
                # image-shaped inputs into a fully connected net (the MNIST MLPs): the samples are vectors of c*h*w
                # values from here on (native.Linearization flattens the inputs the same way)
                if fmap[0] * fmap[1] * fmap[2] <= 0:
                    raise ValueError("empty input feature map")
                fmap = None
            elif isinstance(m, nn.Flatten) and layers and not (fmap is not None and fmap[1:] == (1, 1)):
                _unsupported("Flatten after the first layer (other than of a 1x1 feature map)")
        elif isinstance(m, nn.Dropout):
            if m.training and m.p > 0:
                _unsupported("train-mode Dropout (non-deterministic curvature products)")
        else:
            _unsupported(f"module {type(m).__name__}")
    if not layers:
        _unsupported("a model without Linear or Conv2d layers")
    if fmap is not None and fmap[1:] != (1, 1):
        _unsupported("a convolutional net whose output is a feature map (end with a global average pool)")
    if _check_param_use(layers, offsets) != set(offsets.values()):
        raise ValueError("the optimizer holds trainable parameters that the model does not use")
    return Program(layers, kind, reduction, n_params)


# ---------------------------------------------------------------------------------------------------
# autograd-graph front end
# ---------------------------------------------------------------------------------------------------

def _name(fn):
    return type(fn).__name__


def _leaf_param(fn, want_transposed):
    """The parameter behind an AccumulateGrad node.  ``nn.Linear`` multiplies by ``weight.t()``, so its weight sits
    behind a TBackward0 node and is read as ``[out, in]``; a bare ``x @ W`` (no transpose) would be read transposed,
    so it is refused instead.  Biases must come without a transpose."""
    if fn is None:
        return None
    transposed = _name(fn) == "TBackward0"
    if transposed:
        fn = fn.next_functions[0][0]
        if fn is None:
            return None
    if _name(fn) != "AccumulateGrad":
        _unsupported(f"a weight produced by {_name(fn)}")
    if transposed != want_transposed:
        _unsupported("a matrix product whose weight is not used as `weight.t()` (nn.Linear convention)"
                     if want_transposed else "a transposed bias")
    return fn.variable


def lower_graph(loss, outputs, params_list):
    """Recover the layer program from ``loss.grad_fn`` (GGN needs ``outputs``; pass None for Hessian)."""
    offsets, n_params = flat_offsets(params_list)
    fn = loss.grad_fn
    if fn is None:
        raise ValueError("`forward` returned a loss that does not depend on the parameters")
    name = _name(fn)
    if name == "MseLossBackward0":
        kind, targets, top = "mse", fn._saved_target, fn.next_functions[0][0]
        out_val = fn._saved_self
    elif name == "BinaryCrossEntropyWithLogitsBackward0":
        if fn._saved_weight is not None or fn._saved_pos_weight is not None:
            _unsupported("binary_cross_entropy_with_logits with weights")
        kind, targets, top = "bce", fn._saved_target, fn.next_functions[0][0]
        out_val = fn._saved_self
    elif name == "NllLossBackward0":
        if fn._saved_weight is not None:
            _unsupported("nll_loss with class weights")
        ls = fn.next_functions[0][0]
        if ls is None or _name(ls) != "LogSoftmaxBackward0" or ls._saved_dim not in (1, -1):
            _unsupported("nll_loss that is not fed by log_softmax over the class dimension")
        kind, targets, top = "ce", fn._saved_target, ls.next_functions[0][0]
        out_val = None
        if (targets == fn._saved_ignore_index).any():
            _unsupported("cross-entropy targets equal to ignore_index")
    else:
        _unsupported(f"loss node {name}")
    reduction = _REDUCTION_ENUM.get(int(fn._saved_reduction))
    if reduction is None:
        _unsupported("loss reduction 'none'")
    if outputs is not None and top is not outputs.grad_fn:
        _unsupported("a loss that is not applied directly to the `outputs` returned by `forward`")

    # walk down the chain: [act] <- linear <- [act] <- linear ...
    rev: List[LayerSpec] = []
    pending_act, node, inputs = "none", top, None
    while node is not None:
        nm = _name(node)
        if nm in _ACT_NODES:
            if pending_act != "none":
                _unsupported("two activations in a row")
            pending_act = _ACT_NODES[nm]
            node = node.next_functions[0][0]
            if node is None:
                _unsupported("an activation applied to a constant")
            continue
        if nm in ("ViewBackward0", "ReshapeAliasBackward0", "UnsafeViewBackward0"):
            # flatten of a pooled [batch, c, 1, 1] map (or a no-op view): same numbers, same order
            sizes = tuple(node._saved_self_sym_sizes)
            if len(sizes) == 4 and sizes[2:] != (1, 1):
                _unsupported("flattening a feature map larger than 1x1 (pool first)")
            node = node.next_functions[0][0]
            if node is None:
                _unsupported("a view of a constant")
            continue
        if nm in ("AvgPool2DBackward0", "MeanBackward1"):
            if pending_act != "none":
                _unsupported("an activation directly after an average pool")
            if nm == "AvgPool2DBackward0":
                c, h, w_ = tuple(node._saved_self.shape[1:])
                whole = (tuple(node._saved_kernel_size) == (h, w_) and tuple(node._saved_padding) == (0, 0)
                         and node._saved_divisor_override is None)
            else:
                c, h, w_ = tuple(node._saved_self_sym_sizes)[1:]
                dims = sorted(d % 4 for d in node._saved_dim)
                whole = dims == [2, 3]
            if not whole:
                _unsupported("an average pool that does not cover the whole feature map")
            rev.append(LayerSpec(c, c, "none", False, kind="avgpool", geom=(c, h, w_, 0, 0, 0, 0, 1, 1)))
            node = node.next_functions[0][0]
            if node is None:
                _unsupported("a pool applied to a constant")
            continue
        if nm == "ConvolutionBackward0":
            if (node._saved_transposed or node._saved_groups != 1 or tuple(node._saved_dilation) != (1, 1)
                    or len(set(node._saved_stride)) != 1 or len(set(node._saved_padding)) != 1 or len(node._saved_stride) != 2):
                _unsupported("a convolution with groups, dilation, transposition or anisotropic stride/padding")
            nxt = node.next_functions[0][0]
            weight = _leaf_param(node.next_functions[1][0], False)
            bias = _leaf_param(node.next_functions[2][0], False) if len(node.next_functions) > 2 else None
            has_bias = node._saved_bias_sym_sizes_opt is not None and tuple(node._saved_bias_sym_sizes_opt) not in ((), (0,))
            if weight is None or (has_bias and bias is None):
                _unsupported("a convolution with frozen parameters inside the differentiated part of the graph")
            if id(weight) not in offsets or (bias is not None and id(bias) not in offsets):
                raise ValueError("the graph uses a trainable parameter that is not among the optimizer's parameters")
            saved_in = node._saved_input
            c, h, w_ = tuple(saved_in.shape[1:])
            cout, _, kh, kw = tuple(weight.shape)
            st, pd = int(node._saved_stride[0]), int(node._saved_padding[0])
            ho, wo = (h + 2 * pd - kh) // st + 1, (w_ + 2 * pd - kw) // st + 1
            rev.append(LayerSpec(c * kh * kw, cout, pending_act, bias is not None, offsets[id(weight)],
                                 offsets[id(bias)] if bias is not None else -1, kind="conv2d",
                                 geom=(c, h, w_, kh, kw, st, pd, ho, wo)))
            pending_act = "none"
            if nxt is None:
                inputs = saved_in
            node = nxt
            continue
        if nm == "AddmmBackward0":
            if float(node._saved_alpha) != 1.0 or float(node._saved_beta) != 1.0:
                _unsupported("addmm with alpha/beta != 1")
            bias, nxt, weight = (_leaf_param(node.next_functions[0][0], False), node.next_functions[1][0],
                                 _leaf_param(node.next_functions[2][0], True))
            saved_in = node._saved_mat1
            if bias is None:
                _unsupported("a Linear layer with a frozen bias inside the differentiated part of the graph")
        elif nm == "MmBackward0":
            bias, nxt, weight = None, node.next_functions[0][0], _leaf_param(node.next_functions[1][0], True)
            saved_in = node._saved_self
        else:
            _unsupported(f"graph node {nm}")
        if weight is None:
            _unsupported("a Linear layer with a frozen weight inside the differentiated part of the graph")
        if weight.dim() != 2 or id(weight) not in offsets or (bias is not None and id(bias) not in offsets):
            raise ValueError("the graph uses a trainable parameter that is not among the optimizer's parameters")
        spec = LayerSpec(weight.shape[1], weight.shape[0], pending_act, bias is not None, offsets[id(weight)],
                         offsets[id(bias)] if bias is not None else -1)
        rev.append(spec)
        pending_act = "none"
        if nxt is None:
            inputs = saved_in
        node = nxt
    if pending_act != "none" or not rev or inputs is None:
        _unsupported("a graph that does not end in a Linear or Conv2d layer fed by constant inputs")
    layers = rev[::-1]
    if _check_param_use(layers, offsets) != set(offsets.values()):
        raise ValueError("One of the optimizer's trainable parameters is not used in the graph of `loss`")
    if inputs.dim() != (2 if layers[0].kind == "linear" else 4):
        _unsupported("layers applied to inputs that are neither [batch, features] nor [batch, c, h, w]")
    for below, above in zip(layers, layers[1:]):
        if above.kind == "linear" and below.kind == "conv2d" and below.geom[7:] != (1, 1):
            _unsupported("a Linear layer on a feature map larger than 1x1 (pool first)")
    if kind != "ce" and out_val is not None and tuple(targets.shape) != tuple(out_val.shape):
        _unsupported("a loss with broadcast targets")
    return Program(layers, kind, reduction, n_params, inputs=inputs, targets=targets)
