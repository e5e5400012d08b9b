"""GPU: BASELINE.json configs at FULL size through size-independent properties (the oracle cannot run these in
seconds): sharding equivalence, linearity, symmetry, positive semi-definiteness, engine agreement, and the public
optimizer API end to end."""
import warnings

import pytest
import torch

from helpers import build_loss, build_model

from pytorchhessianfree_b200 import HessianFree
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem

pytestmark = pytest.mark.gpu
DEV = "cuda"
AE = dict(widths=[784, 1000, 500, 250, 30, 250, 500, 1000, 784], act="sigmoid", bias=[True] * 8, frozen=[], loss="bce",
          linear_after=[3])  # Martens' deep autoencoder, linear 30-unit code layer (BASELINE.json configs[2])
MLP = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")


def l2rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


@pytest.mark.parametrize("curv", ["ggn", "hessian"])
def test_autoencoder_60000_sharded_equals_unsharded(curv):
    """configs[2]: batch 60 000 as 10 chunks of 6 000 (the acc_step sharding) vs 2 chunks of 30 000."""
    torch.manual_seed(0)
    model = build_model(AE).to(DEV)
    loss_fn = build_loss(AE, "mean")
    n = 60000 if curv == "ggn" else 12000  # the Hessian path keeps 4 extra buffers per layer
    x = torch.rand(n, 784, device=DEV)
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params])
    assert theta.numel() == 2837314
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine="tc")
    v, w = torch.randn_like(theta), torch.randn_like(theta)
    c = n // 10
    many = NativeProblem(net, theta, curv, [(x[i * c:(i + 1) * c], x[i * c:(i + 1) * c]) for i in range(10)])
    l1, g1 = many.linearize(), many.gradient()
    Bv1, Bw1 = many.mvp(v), many.mvp(w)
    del many
    few = NativeProblem(net, theta, curv, [(x[: n // 2], x[: n // 2]), (x[n // 2:], x[n // 2:])])
    l2, g2 = few.linearize(), few.gradient()
    Bv2 = few.mvp(v)
    assert abs(l1.item() - l2.item()) <= 1e-6 * abs(l2.item())
    # same kernels, different partition of the 60 000-term FP32 batch sums (split-K boundaries move with the chunk
    # size): agreement is limited by FP32 accumulation order, not by the algorithm
    assert rel(g1, g2) < 3e-4 and rel(Bv1, Bv2) < 3e-4
    assert l2rel(g1, g2) < 1e-4 and l2rel(Bv1, Bv2) < 1e-4
    a, b = torch.dot(w.double(), Bv1.double()).item(), torch.dot(v.double(), Bw1.double()).item()
    # w.Bv and v.Bw are sums of 2.8 M terms that cancel to ~1e-3 of |v||Bw|: compare on that scale
    assert abs(a - b) <= 1e-5 * (v.double().norm() * Bw1.double().norm()).item(), "curvature matrix must be symmetric"
    if curv == "ggn":
        assert torch.dot(v.double(), Bv1.double()).item() >= 0.0


def test_engines_agree_on_the_mlp_config():
    torch.manual_seed(0)
    model = build_model(MLP).to(DEV)
    loss_fn = build_loss(MLP, "mean")
    x, t = torch.rand(4096, 784, device=DEV), torch.randint(0, 10, (4096,), device=DEV)
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params])
    v = torch.randn_like(theta)
    out = {}
    for engine in ("simt", "tc"):
        for curv in ("ggn", "hessian"):
            prob = NativeProblem(NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine), theta,
                                 curv, [(x, t)])
            prob.linearize()
            out[engine, curv] = (prob.gradient(), prob.mvp(v), prob.fisher_diag())
    for curv in ("ggn", "hessian"):
        for a, b in zip(out["tc", curv], out["simt", curv]):
            # split-precision tensor tiles vs FP32 FMA tiles.  With 4096 x 1024 ReLU units a few pre-activations sit within
            # rounding of 0 and flip their mask between the engines (as they do between torch-CPU and torch-CUDA);
            # each flip moves single entries by O(1/N), so the bulk is compared in L2 and the tail loosely.
            assert l2rel(a, b) < 1e-4 and rel(a, b) < 5e-3


@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_public_api_trains_the_mlp_config(engine):
    """configs[1] through HessianFree.acc_step with the Fisher preconditioner: the loss must go down and both
    engines must follow the same trajectory."""
    torch.manual_seed(0)
    model = build_model(MLP).to(DEV)
    loss_fn = build_loss(MLP, "mean")
    g = torch.Generator(device=DEV).manual_seed(1)
    proto = torch.randn(10, 784, device=DEV, generator=g)
    t = torch.randint(0, 10, (4096,), device=DEV, generator=g)
    x = proto[t] + 0.5 * torch.randn(4096, 784, device=DEV, generator=g)  # learnable synthetic classes
    opt = HessianFree(model.parameters(), engine=engine)
    losses = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(3):
            M = opt.get_preconditioner(model, loss_fn, x, t, "mean")
            losses.append(opt.acc_step(model, loss_fn, [(x[:2048], t[:2048]), (x[2048:], t[2048:])], M_func=M))
    st = opt.state
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < 0.5 * st["init_losses"][0]
    assert st["init_losses"][0] == pytest.approx(2.33, abs=0.15)
    assert all(1 <= n <= 250 for n in st["num_cg_iters"])
    test_public_api_trains_the_mlp_config.seen = getattr(test_public_api_trains_the_mlp_config, "seen", {})
    test_public_api_trains_the_mlp_config.seen[engine] = losses
    seen = test_public_api_trains_the_mlp_config.seen
    if len(seen) == 2:
        for a, b in zip(seen["tc"], seen["simt"]):
            assert a == pytest.approx(b, rel=1e-3)


@pytest.mark.parametrize("curv", ["ggn", "hessian"])
def test_products_are_bitwise_repeatable(curv):
    """The same product launched repeatedly, and two independently allocated linearisations of the same problem, return
    the same bits: every reduction is fixed-order, nothing is atomically accumulated, and no kernel writes outside its
    tile (a 7 500-row chunk has an odd number of 128-row blocks, so the CTA-pair grid is padded -- the padding CTA of
    an early build scribbled over row 0 of a cotangent, which only this test caught)."""
    torch.manual_seed(0)
    model = build_model(AE).to(DEV)
    loss_fn = build_loss(AE, "mean")
    x = torch.rand(2 * 7500, 784, device=DEV)
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params])
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine="tc")
    v = torch.randn_like(theta)

    def fresh():
        prob = NativeProblem(net, theta, curv, [(x[:7500], x[:7500]), (x[7500:], x[7500:])])
        prob.linearize()
        return prob, prob.gradient()

    (p1, g1), (p2, g2) = fresh(), fresh()
    assert torch.equal(g1, g2)
    outs = [p1.mvp(v) for _ in range(4)] + [p2.mvp(v) for _ in range(2)]
    for o in outs[1:]:
        assert torch.equal(outs[0], o)
    assert torch.equal(p1.fisher_diag(), p2.fisher_diag())
