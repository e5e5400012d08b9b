"""CPU: host-side logic that needs no kernels -- lowering, snapshot grid, step selection, validation."""
import warnings

import pytest
import torch
from torch import nn

from helpers import GOLDEN, SPECS, build_loss, build_model, make_data

from pytorchhessianfree_b200 import HessianFree, cg_backtracking, cg_efficient_backtracking, cg_storing_grid, simple_linesearch
from pytorchhessianfree_b200.cg import cg
from pytorchhessianfree_b200.lowering import lower_graph, lower_module

SEL = torch.load(f"{GOLDEN}/selection.pt", weights_only=False)
CG = torch.load(f"{GOLDEN}/cg.pt", weights_only=False)


def test_storing_grid_equals_reference():
    for m, grid in CG["grids"].items():
        assert cg_storing_grid(m) == grid
    with pytest.raises(ValueError):
        cg_storing_grid(10, gamma=0.5)


def test_backtracking_toy_list():  # reference tests/test_cg_backtracking.py:16-44
    assert cg_backtracking(lambda s: s, SEL["toy"]) == (1, 1.0)
    assert cg_efficient_backtracking(lambda s: s, SEL["toy"]) == (4, 2.4)
    assert cg_efficient_backtracking(lambda s: s, SEL["toy"], lookahead=3) == (4, 2.4)
    f = lambda s: s  # noqa: E731
    f.many = lambda ss: list(ss)
    assert cg_efficient_backtracking(f, SEL["toy"], lookahead=4) == (4, 2.4)
    assert cg_backtracking(f, SEL["toy"]) == (1, 1.0)


def test_linesearch_matches_reference_fixture():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for c in SEL["linesearch"]:
            f = lambda s: float(0.5 * s @ c["A"] @ s - c["b"] @ s)  # noqa: E731
            a, fa = simple_linesearch(f, -c["b"], c["step"])
            assert a == pytest.approx(c["alpha"]) and fa == pytest.approx(c["f"], rel=1e-5)
    with pytest.raises(ValueError):
        simple_linesearch(lambda s: 0.0, torch.zeros(2), torch.zeros(2), beta=1.0)
    with pytest.raises(ValueError):
        simple_linesearch(lambda s: 0.0, torch.zeros(2), torch.zeros(2), c=-1.0)


@pytest.mark.parametrize("name", sorted(SPECS))
def test_graph_and_module_lowering_agree(name):
    spec = SPECS[name]
    torch.manual_seed(0)
    model, loss_fn = build_model(spec), build_loss(spec, "sum")
    x, t = make_data(spec, 5, 1)
    params = [p for p in model.parameters() if p.requires_grad]
    out = model(x)
    g = lower_graph(loss_fn(out, t), out, params)
    m = lower_module(model, loss_fn, params)
    assert g.loss == m.loss == spec["loss"] and g.reduction == m.reduction == "sum"
    assert g.n_params == m.n_params == sum(p.numel() for p in params)
    # the graph only sees the differentiated suffix; frozen leading layers are folded into its inputs
    gl, ml = g.layers, m.layers[len(m.layers) - len(g.layers):]
    for a, b in zip(gl, ml):
        assert (a.in_features, a.out_features, a.act, a.has_bias, a.w_offset, a.b_offset) == \
               (b.in_features, b.out_features, b.act, b.has_bias, b.w_offset, b.b_offset)
    assert g.inputs.shape == (5, gl[0].in_features)
    assert torch.equal(g.targets, t)


def test_unlowerable_graphs_are_refused_loudly():
    p = torch.randn(4, requires_grad=True)
    with pytest.raises(NotImplementedError):
        lower_graph((p ** 4).sum(), None, [p])
    model = nn.Sequential(nn.Linear(4, 4), nn.Softplus(), nn.Linear(4, 2))
    with pytest.raises(NotImplementedError):
        lower_module(model, nn.MSELoss(), list(model.parameters()))
    with pytest.raises(NotImplementedError):
        lower_module(nn.Sequential(nn.Linear(4, 2)), nn.L1Loss(), [])


def test_constructor_validation_matches_reference():  # reference optimizer.py:80-115
    w = [nn.Parameter(torch.zeros(3))]
    for kw in (dict(curvature_opt="fisher"), dict(damping=-1.0), dict(cg_max_iter=0), dict(lr=-0.1)):
        with pytest.raises(ValueError):
            HessianFree(w, **kw)
    with pytest.raises(ValueError):
        HessianFree([dict(params=w), dict(params=[nn.Parameter(torch.zeros(2))])])
    with pytest.warns(UserWarning, match="won't get adapted"):
        opt = HessianFree(w, damping=0.0)
    assert opt.adapt_damping is False
    opt = HessianFree(w)
    assert set(opt.param_groups[0]) >= {"curvature_opt", "damping", "cg_max_iter", "lr"}
    assert opt.param_groups[0]["damping"] == 1.0 and opt.param_groups[0]["cg_max_iter"] == 250


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU path"):
        cg(lambda v: v, torch.ones(4))
    model = nn.Sequential(nn.Linear(3, 2))
    opt = HessianFree(model.parameters())
    x, t = torch.rand(4, 3), torch.rand(4, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step(lambda: (lambda o: (nn.functional.mse_loss(o, t), o))(model(x)))


def test_target_function_memo_and_linesearch_shortcuts():
    """The step asks for the same candidate losses several times; every distinct candidate is evaluated once, in as
    few passes as possible, and the selection results equal the unmemoised ones (host logic only: a stub evaluates
    the losses)."""
    from pytorchhessianfree_b200.problem import NativeProblem

    class Stub(NativeProblem):
        def __init__(self):  # no device state: only target_function()/losses_at() are exercised
            self.passes = []

        def losses_at(self, steps):
            self.passes.append(len(steps))
            return [float((s - 1.0).pow(2).sum()) for s in steps]

    stub = Stub()
    f = stub.target_function()
    cands = [torch.full((4,), v) for v in (0.0, 0.5, 0.9, 1.2, 2.0)]
    zero = torch.zeros(4)
    primed = f.prime([zero, cands[0], cands[-1], cands[-1], cands[-2], cands[-3]])
    assert stub.passes == [5] and primed[2] == primed[3]  # duplicates share one evaluation
    want_best = min(range(len(cands)), key=lambda i: float((cands[i] - 1.0).pow(2).sum()))
    best, val = cg_efficient_backtracking(f, cands, lookahead=3)
    assert best == want_best and val == pytest.approx(float((cands[best] - 1.0).pow(2).sum()))
    assert stub.passes == [5, 1]  # the walk went one candidate beyond the primed window
    grad = -torch.ones(4)
    alpha, f_alpha = simple_linesearch(f, grad, cands[best], init_alpha=1.0, f_0=primed[0])
    assert alpha == 1.0 and f_alpha == val and stub.passes == [5, 1]  # nothing new to evaluate
    plain = lambda s: float((s - 1.0).pow(2).sum())  # noqa: E731
    assert simple_linesearch(plain, grad, cands[best], init_alpha=1.0) == (alpha, f_alpha)
    # a tensor modified in place is a new candidate
    cands[0].add_(3.0)
    assert f(cands[0]) == pytest.approx(4 * 4.0) and stub.passes[-1] == 1


def test_repeated_module_instances_are_kept():
    """A Sequential that applies ONE activation object after every Linear must lower to activation after every layer
    (``children()`` de-duplicates instances; the walk uses ``_modules``)."""
    act = nn.ReLU()
    model = nn.Sequential(nn.Linear(4, 5), act, nn.Linear(5, 6), act, nn.Linear(6, 3))
    prog = lower_module(model, nn.MSELoss(), list(model.parameters()))
    assert [l.act for l in prog.layers] == ["relu", "relu", "none"]


def test_shared_weights_are_refused():
    """Two uses of one Linear would write the same slice of the flat vector twice in the transposed sweep."""
    lin = nn.Linear(5, 5)
    model = nn.Sequential(nn.Linear(4, 5), nn.Tanh(), lin, nn.Tanh(), lin)
    with pytest.raises(NotImplementedError, match="more than one layer"):
        lower_module(model, nn.MSELoss(), list(model.parameters()))
    x = torch.rand(3, 4)
    out = model(x)
    with pytest.raises(NotImplementedError, match="more than one layer"):
        lower_graph(nn.MSELoss()(out, torch.rand(3, 5)), out, list(model.parameters()))


def test_untransposed_matmul_is_refused():
    """``x @ W`` with a square W has the shapes of a Linear layer but the transposed meaning: refuse, do not guess."""
    W = nn.Parameter(torch.rand(4, 4))
    x = torch.rand(3, 4)
    out = x @ W
    with pytest.raises(NotImplementedError, match="weight.t"):
        lower_graph(nn.MSELoss()(out, torch.rand(3, 4)), out, [W])


def test_conv_lowering_geometry():
    """nn.Conv2d / global average pool -> layer program: the unfolded width c_in*k_h*k_w, the output map, flat offsets."""
    model = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 6, 3, stride=2, padding=1, bias=False), nn.ReLU(),
                          nn.AvgPool2d(6), nn.Flatten(), nn.Linear(6, 4))
    prog = lower_module(model, nn.CrossEntropyLoss(), list(model.parameters()), input_shape=(3, 12, 12))
    kinds = [(l.kind, l.in_features, l.out_features, l.act, l.has_bias) for l in prog.layers]
    assert kinds == [("conv2d", 27, 8, "relu", True), ("conv2d", 72, 6, "relu", False), ("avgpool", 6, 6, "none", False),
                     ("linear", 6, 4, "none", True)]
    assert prog.layers[0].geom == (3, 12, 12, 3, 3, 1, 1, 12, 12) and prog.layers[1].geom == (8, 12, 12, 3, 3, 2, 1, 6, 6)
    assert [l.w_offset for l in prog.layers] == [0, 224, -1, 656] and prog.n_params == 656 + 24 + 4
    bad = nn.Sequential(nn.Conv2d(3, 8, 3), nn.AvgPool2d(2), nn.Flatten())
    with pytest.raises(NotImplementedError, match="whole feature map"):
        lower_module(bad, nn.MSELoss(), list(bad.parameters()), input_shape=(3, 12, 12))
    bad = nn.Sequential(nn.Conv2d(3, 8, 3), nn.AdaptiveAvgPool2d(1), nn.Flatten())
    with pytest.raises(NotImplementedError, match="feature-map size is unknown"):
        lower_module(bad, nn.MSELoss(), list(bad.parameters()))
    bad = nn.Sequential(nn.Conv2d(3, 8, 3), nn.ReLU())
    with pytest.raises(NotImplementedError, match="global average pool"):
        lower_module(bad, nn.MSELoss(), list(bad.parameters()), input_shape=(3, 12, 12))


def test_graph_and_module_lowering_agree_on_a_conv_net():
    model = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 6, 3, stride=2, padding=1, bias=False), nn.ReLU(),
                          nn.AvgPool2d(6), nn.Flatten(), nn.Linear(6, 4))
    x, t = torch.rand(2, 3, 12, 12), torch.tensor([1, 2])
    out = model(x)
    g = lower_graph(nn.CrossEntropyLoss()(out, t), out, list(model.parameters()))
    m = lower_module(model, nn.CrossEntropyLoss(), list(model.parameters()), input_shape=(3, 12, 12))
    assert [l.signature() for l in g.layers] == [l.signature() for l in m.layers]
    assert g.inputs is x or torch.equal(g.inputs, x)
    pooled = nn.Sequential(nn.Conv2d(3, 8, 3), nn.Tanh(), nn.AdaptiveAvgPool2d(1), nn.Flatten())
    o = pooled(x)
    g2 = lower_graph(nn.MSELoss()(o, torch.rand(2, 8)), o, list(pooled.parameters()))
    assert [(l.kind, l.act) for l in g2.layers] == [("conv2d", "tanh"), ("avgpool", "none")]


def test_exchange_ranges_are_whole_quanta_and_disjoint():
    """hf_allreduce_multimem wants offset % 4 == 0 and count % (4 * world) == 0; the sliced form of the exchange cuts
    the vector at a widened end point and must not reduce any element twice."""
    from pytorchhessianfree_b200.dist import padded_range

    for world in (2, 4, 8):
        q = 4 * world
        for numel in (1, q - 1, q, 2837314, 669706):
            limit = (numel + q - 1) // q * q
            lo, hi = padded_range(0, numel, q, limit)
            assert (lo, hi) == (0, limit) and (hi - lo) % q == 0
            for split in (4, 784000, numel // 3):
                if not 0 < split < numel:
                    continue
                cut = padded_range(0, split, q, limit)[1]
                first, second = padded_range(0, cut, q, limit), padded_range(cut, numel, q, limit)
                assert first == (0, cut) and second == (cut, limit)  # disjoint, together the whole padded vector
                assert cut >= split and cut % q == 0 and cut % 4 == 0


def test_flatten_in_front_of_a_fully_connected_net_is_lowered():
    """Image-shaped inputs into an MLP (nn.Flatten first): the layer program starts at the first Linear; a Linear applied
    to an unflattened feature map (torch would contract the last axis only) stays refused."""
    import torch.nn as nn

    from pytorchhessianfree_b200.lowering import lower_module

    loss = nn.CrossEntropyLoss()
    mlp = nn.Sequential(nn.Flatten(), nn.Linear(784, 32), nn.ReLU(), nn.Linear(32, 10))
    prog = lower_module(mlp, loss, list(mlp.parameters()), input_shape=(1, 28, 28))
    assert [(l.kind, l.in_features, l.out_features, l.act) for l in prog.layers] == [("linear", 784, 32, "relu"), ("linear", 32, 10, "none")]
    bare = nn.Sequential(nn.Linear(28, 32), nn.ReLU(), nn.Linear(32, 10))
    with pytest.raises(NotImplementedError, match="feature map"):
        lower_module(bare, loss, list(bare.parameters()), input_shape=(1, 28, 28))
