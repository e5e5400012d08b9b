"""GPU: the first convolutional slice (reference examples/run_allcnnc_cifar100_deepobs.py,
eval mode): Conv2d / ReLU / global average pool / Linear nets lowered to the layer program -- loss, gradient and GGN
products against the CPU oracle (autograd on the same module), chunked == full batch, optimizer steps against the
oracle's, Hessian products, the empirical-Fisher diagonal (per-sample gradients summed over positions BEFORE the
square) against a per-sample autograd loop, and a loud refusal for what is not lowered."""
import copy
import warnings

import pytest
import torch
from torch import nn

import hf_oracle as O
from helpers import GOLDEN, conv_fixture_case

from pytorchhessianfree_b200 import HessianFree
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem

pytestmark = pytest.mark.gpu
DEV = "cuda"


def small_cnn(act=nn.ReLU):
    # stride 2, 'same' and 'valid' 3x3, a 1x1 convolution, bias on and off, global average pool: the shapes of All-CNN-C in small
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), act(), nn.Conv2d(8, 8, 3, stride=2, padding=1, bias=False), act(),
                         nn.Conv2d(8, 16, 3), act(), nn.Conv2d(16, 10, 1), act(), nn.AvgPool2d(4), nn.Flatten())


def cnn_with_head():
    return nn.Sequential(nn.Conv2d(2, 6, 3, padding=1), nn.Sigmoid(), nn.Conv2d(6, 12, 5, stride=2, padding=2), nn.Sigmoid(),
                         nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(12, 7), nn.Tanh(), nn.Linear(7, 5))


def allcnnc(classes=100):
    """All-CNN-C (Springenberg et al.; the DeepOBS cifar100_allcnnc architecture with symmetric padding), eval mode."""
    def block(cin, cout, k, stride=1, pad=0):
        return [nn.Conv2d(cin, cout, k, stride=stride, padding=pad), nn.ReLU()]
    layers = (block(3, 96, 3, pad=1) + block(96, 96, 3, pad=1) + block(96, 96, 3, stride=2, pad=1) + block(96, 192, 3, pad=1)
              + block(192, 192, 3, pad=1) + block(192, 192, 3, stride=2, pad=1) + block(192, 192, 3) + block(192, 192, 1)
              + block(192, classes, 1) + [nn.AvgPool2d(6), nn.Flatten()])
    return nn.Sequential(*layers)


def errs(got, want):
    got, want = got.double().cpu(), want.double()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item(), ((got - want).norm() / want.norm().clamp_min(1e-30)).item()


def device_problem(model, loss_fn, chunks, engine, curv="ggn"):
    m = copy.deepcopy(model).to(DEV)
    params = [p for p in m.parameters() if p.requires_grad]
    prog = lower_module(m, loss_fn, params, input_shape=tuple(chunks[0][0].shape[1:]))
    theta = torch.cat([p.detach().reshape(-1) for p in params])
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    return NativeProblem(net, theta, curv, [(x.to(DEV), t.to(DEV)) for x, t in chunks])


CASES = {
    "small_cnn_ce": (small_cnn, nn.CrossEntropyLoss, (3, 12, 12), 10, "ce"),
    "small_cnn_tanh_mse": (lambda: small_cnn(nn.Tanh), nn.MSELoss, (3, 12, 12), 10, "mse"),
    "cnn_head_bce": (cnn_with_head, nn.BCEWithLogitsLoss, (2, 9, 9), 5, "bce"),
}


def make_case(name, n, seed):
    build, loss_cls, shape, classes, kind = CASES[name]
    torch.manual_seed(seed)
    model = build()
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.rand(n, *shape, generator=g)
    t = torch.randint(0, classes, (n,), generator=g) if kind == "ce" else torch.rand(n, classes, generator=g)
    return model, loss_cls(), x, t


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("n", [1, 6, 37])
@pytest.mark.parametrize("name", sorted(CASES))
def test_conv_products_match_oracle(name, n, engine):
    for seed in (0, 1, 42):
        model, loss_fn, x, t = make_case(name, n, seed)
        params = list(model.parameters())
        out = model(x)
        loss = loss_fn(out, t)
        v = torch.randn(sum(p.numel() for p in params))
        want_g = O.flatten(torch.autograd.grad(loss, params, retain_graph=True))
        want_G = O.Gv(loss, out, params, v)
        prob = device_problem(model, loss_fn, [(x, t)], engine)
        got_loss = float(prob.linearize().item())
        assert abs(got_loss - float(loss)) <= 1e-5 * abs(float(loss))
        e_g, e_G = errs(prob.gradient(), want_g), errs(prob.mvp(v.to(DEV)), want_G)
        assert max(e_g) < 1e-4, f"gradient: max {e_g[0]:.1e} l2 {e_g[1]:.1e}"
        assert max(e_G) < 1e-4, f"GGN product: max {e_G[0]:.1e} l2 {e_G[1]:.1e}"
        # Hessian product (Pearlmutter sweep through fold / unpool with the second-order activation terms)
        want_H = O.Hv(loss, params, v)
        hprob = device_problem(model, loss_fn, [(x, t)], engine, curv="hessian")
        hprob.linearize(), hprob.gradient()
        e_H = errs(hprob.mvp(v.to(DEV)), want_H)
        assert max(e_H) < 1e-4, f"Hessian product: max {e_H[0]:.1e} l2 {e_H[1]:.1e}"


CONV_GOLD = torch.load(f"{GOLDEN}/conv.pt", weights_only=False)


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("i", range(len(CONV_GOLD)))
def test_conv_products_match_reference_fixture(i, engine):
    """Against tests/golden/conv.pt: loss, gradient, `_Gv`, `_Hv` as the UNMODIFIED reference returned them for these
    nets (the GPU box has no reference), and against the float64 dense `J^T H J v` / `H v` stored with them."""
    c = CONV_GOLD[i]
    model, loss_fn, x, t, v = conv_fixture_case(c)
    prob = device_problem(model, loss_fn, [(x, t)], engine)
    got_loss = float(prob.linearize().item())
    assert abs(got_loss - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    assert max(errs(prob.gradient(), c["grad"])) < 1e-4
    G = prob.mvp(v.to(DEV))
    assert max(errs(G, c["Gv"])) < 1e-4 and max(errs(G, c["Gv_dense64"])) < 1e-4
    hprob = device_problem(model, loss_fn, [(x, t)], engine, curv="hessian")
    hprob.linearize(), hprob.gradient()
    H = hprob.mvp(v.to(DEV))
    assert max(errs(H, c["Hv"])) < 1e-4 and max(errs(H, c["Hv_dense64"])) < 1e-4


@pytest.mark.parametrize("name", sorted(CASES))
def test_conv_chunked_equals_full_batch(name):
    model, loss_fn, x, t = make_case(name, 23, 3)
    full = device_problem(model, loss_fn, [(x, t)], "tc")
    parts = device_problem(model, loss_fn, [(x[:7], t[:7]), (x[7:8], t[7:8]), (x[8:], t[8:])], "tc")
    v = torch.randn_like(full.theta)
    assert abs(full.linearize().item() - parts.linearize().item()) <= 1e-6 * abs(full.linearize().item())
    for a, b in ((full.gradient(), parts.gradient()), (full.mvp(v), parts.mvp(v))):
        assert max(errs(b, a.cpu())) < 2e-5
    hfull = device_problem(model, loss_fn, [(x, t)], "tc", curv="hessian")
    hparts = device_problem(model, loss_fn, [(x[:7], t[:7]), (x[7:8], t[7:8]), (x[8:], t[8:])], "tc", curv="hessian")
    for p in (hfull, hparts):
        p.linearize(), p.gradient()
    assert max(errs(hparts.mvp(v), hfull.mvp(v).cpu())) < 2e-5


def test_allcnnc_ggn_product_at_batch_64():
    """BASELINE.json configs[4]'s architecture at the size BASELINE.md quotes the CPU time for (N = 64, 537 ms per
    reference `_Gv`): gradient and GGN product of the tensor-core path against the CPU oracle, rtol 1e-4."""
    torch.manual_seed(0)
    model = allcnnc()
    loss_fn = nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(7)
    x, t = torch.rand(64, 3, 32, 32, generator=g), torch.randint(0, 100, (64,), generator=g)
    params = list(model.parameters())
    assert sum(p.numel() for p in params) == 1387108  # P of configs[4]
    out = model(x)
    loss = loss_fn(out, t)
    v = torch.randn(1387108, generator=g)
    want_g = O.flatten(torch.autograd.grad(loss, params, retain_graph=True))
    want_G = O.Gv(loss, out, params, v)
    prob = device_problem(model, loss_fn, [(x, t)], "tc")
    assert abs(prob.linearize().item() - float(loss)) <= 1e-5 * float(loss)
    e_g, e_G = errs(prob.gradient(), want_g), errs(prob.mvp(v.to(DEV)), want_G)
    # BASELINE.json configs[4] asks for curvature_opt='hessian' (CPU reference at this size: 1 282 ms per `_Hv`).  The
    # Hessian of a ReLU net is made of cross terms that cancel, so float32 noise is larger relative to |Hv| than for
    # the GGN: the comparison is against the float64 oracle, and the float32 oracle's own distance from it is printed.
    m64 = copy.deepcopy(model).double()
    p64 = list(m64.parameters())
    want_H32 = O.Hv(loss, params, v)
    want_H = O.Hv(loss_fn(m64(x.double()), t), p64, v.double())
    print(f"\nfloat32 oracle vs float64 oracle, Hessian product: max {errs(want_H32, want_H)[0]:.1e} l2 {errs(want_H32, want_H)[1]:.1e}")
    hprob = device_problem(model, loss_fn, [(x, t)], "tc", curv="hessian")
    hprob.linearize(), hprob.gradient()
    e_H = errs(hprob.mvp(v.to(DEV)), want_H)
    print(f"\nAll-CNN-C N=64: gradient max {e_g[0]:.1e} l2 {e_g[1]:.1e}; GGN product max {e_G[0]:.1e} l2 {e_G[1]:.1e}; "
          f"Hessian product max {e_H[0]:.1e} l2 {e_H[1]:.1e}")
    assert max(e_g) < 1e-4 and max(e_G) < 1e-4 and max(e_H) < 1e-4
    # symmetric, positive semi-definite
    w = torch.randn_like(v)
    Bv, Bw = prob.mvp(v.to(DEV)), prob.mvp(w.to(DEV))
    a, b = torch.dot(w.to(DEV).double(), Bv.double()).item(), torch.dot(v.to(DEV).double(), Bw.double()).item()
    assert abs(a - b) <= 1e-4 * max(abs(a), abs(b)) and torch.dot(v.to(DEV).double(), Bv.double()).item() >= 0.0


def test_acc_step_on_a_conv_net_follows_the_oracle():
    model, loss_fn, x, t = make_case("small_cnn_ce", 24, 5)
    ref_model = copy.deepcopy(model)
    model = model.to(DEV)
    opt = HessianFree(model.parameters(), cg_max_iter=15)
    orc = O.OracleHF(ref_model.parameters(), cg_max_iter=15)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for step in range(3):
            chunks = [(x[:10], t[:10]), (x[10:], t[10:])]
            opt.acc_step(model, loss_fn, [(a.to(DEV), b.to(DEV)) for a, b in chunks], reduction="mean")
            orc.acc_step(ref_model, loss_fn, chunks, reduction="mean")
            assert opt.state["init_losses"][-1] == pytest.approx(orc.log["init_losses"][-1], rel=1e-4)
    for p, q in zip(model.parameters(), ref_model.parameters()):
        assert torch.allclose(p.data.cpu(), q.data, atol=1e-4)


def test_step_and_static_products_through_the_autograd_graph():
    """`step(forward)` and the static `_Gv` only see an autograd graph (reference optimizer.py:137-151, :457-462): the
    conv layer program is recovered from ConvolutionBackward0 / AvgPool2DBackward0 / ViewBackward0 nodes."""
    model, loss_fn, x, t = make_case("small_cnn_ce", 12, 8)
    ref_model = copy.deepcopy(model)
    model = model.to(DEV)
    xd, td = x.to(DEV), t.to(DEV)
    params, ref_params = list(model.parameters()), list(ref_model.parameters())
    v = torch.randn(sum(p.numel() for p in params))
    out, ref_out = model(xd), ref_model(x)
    got = HessianFree._Gv(loss_fn(out, td), out, params, v.to(DEV))
    want = O.Gv(loss_fn(ref_out, t), ref_out, ref_params, v)
    assert max(errs(got, want)) < 1e-4
    opt = HessianFree(model.parameters(), cg_max_iter=12)
    orc = O.OracleHF(ref_model.parameters(), cg_max_iter=12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(3):
            f1 = opt.step(lambda: (lambda o: (loss_fn(o, td), o))(model(xd)))
            f2 = orc.step(lambda: (lambda o: (loss_fn(o, t), o))(ref_model(x)))
            assert f1 == pytest.approx(f2, rel=1e-3)
    assert opt.state["num_cg_iters"] == orc.log["num_cg_iters"]


def test_allcnnc_fisher_diagonal_at_batch_12():
    """The preconditioner of the reference's All-CNN-C example: empirical-Fisher diagonal of all nine convolutions
    (3x3 with and without stride, 1x1, K from 27 to 1 728, up to 1 024 positions per sample) against the per-sample loop."""
    torch.manual_seed(0)
    model = allcnnc()
    loss_fn = nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(11)
    x, t = torch.rand(12, 3, 32, 32, generator=g), torch.randint(0, 100, (12,), generator=g)
    want = per_sample_ef(model, loss_fn, x, t, "mean")
    prob = device_problem(model, loss_fn, [(x[:5], t[:5]), (x[5:], t[5:])], "tc")
    prob.linearize()
    e = errs(prob.fisher_diag(), want)
    print(f"\nAll-CNN-C N=12 Fisher diagonal: max {e[0]:.1e} l2 {e[1]:.1e}")
    assert max(e) < 1e-4


def per_sample_ef(model, loss_fn, x, t, reduction):
    """sum_n g_n^2 ("sum") or (1/N) sum_n g_n^2 ("mean") with g_n the gradient of sample n's own loss: the reference's
    autograd variant (preconditioners.py:63-105) with the batch dimension kept for the convolutions."""
    params = [p for p in model.parameters() if p.requires_grad]
    acc = torch.zeros(sum(p.numel() for p in params), dtype=torch.float64)
    for i in range(x.shape[0]):
        li = loss_fn(model(x[i:i + 1]), t[i:i + 1])
        acc += O.flatten(torch.autograd.grad(li, params)).double() ** 2
    return acc / x.shape[0] if reduction == "mean" else acc


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("n", [1, 5, 70])
@pytest.mark.parametrize("name", sorted(CASES))
def test_conv_fisher_diagonal_matches_per_sample_loop(name, n, engine):
    model, loss_fn, x, t = make_case(name, n, 1)
    want = per_sample_ef(model, loss_fn, x, t, "mean")
    prob = device_problem(model, loss_fn, [(x, t)], engine)
    prob.linearize()
    e = errs(prob.fisher_diag(), want)
    assert max(e) < 1e-4, f"one chunk: max {e[0]:.1e} l2 {e[1]:.1e}"
    if n >= 5:  # chunked == full (the chunk sum is exact for a sum over samples)
        parts = device_problem(model, loss_fn, [(x[:2], t[:2]), (x[2:], t[2:])], engine)
        parts.linearize()
        e = errs(parts.fisher_diag(), want)
        assert max(e) < 1e-4, f"two chunks: max {e[0]:.1e} l2 {e[1]:.1e}"


def test_conv_preconditioner_through_the_public_api():
    """get_preconditioner + acc_step on a conv net (the call sequence of the reference's DeepOBS example)."""
    model, loss_fn, x, t = make_case("small_cnn_ce", 24, 2)
    m = copy.deepcopy(model).to(DEV)
    opt = HessianFree(m.parameters(), cg_max_iter=20)
    xd, td = x.to(DEV), t.to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        M = opt.get_preconditioner(m, loss_fn, xd, td, "mean")
        before = float(loss_fn(m(xd), td))
        opt.acc_step(m, loss_fn, [(xd, td)], M_func=M)
        after = float(loss_fn(m(xd), td))
    assert after < before


def test_image_shaped_inputs_into_a_flatten_first_mlp():
    """nn.Flatten in front of a fully connected net, inputs [n, c, h, w] (the MNIST-MLP pattern): gradient, GGN product
    and Fisher diagonal against the oracle, and an `acc_step` through the public API."""
    torch.manual_seed(3)
    model = nn.Sequential(nn.Flatten(), nn.Linear(3 * 6 * 6, 24), nn.Tanh(), nn.Linear(24, 5))
    loss_fn = nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(5)
    x, t = torch.rand(19, 3, 6, 6, generator=g), torch.randint(0, 5, (19,), generator=g)
    params = list(model.parameters())
    out = model(x)
    loss = loss_fn(out, t)
    v = torch.randn(sum(p.numel() for p in params), generator=g)
    prob = device_problem(model, loss_fn, [(x, t)], "tc")
    assert abs(prob.linearize().item() - float(loss)) <= 1e-5 * float(loss)
    assert max(errs(prob.gradient(), O.flatten(torch.autograd.grad(loss, params, retain_graph=True)))) < 1e-4
    assert max(errs(prob.mvp(v.to(DEV)), O.Gv(loss, out, params, v))) < 1e-4
    assert max(errs(prob.fisher_diag(), per_sample_ef(model, loss_fn, x, t, "mean"))) < 1e-4
    m = copy.deepcopy(model).to(DEV)
    opt = HessianFree(m.parameters(), cg_max_iter=10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt.acc_step(m, loss_fn, [(x.to(DEV), t.to(DEV))])
    assert float(loss_fn(m(x.to(DEV)), t.to(DEV))) < float(loss)


def test_unlowered_conv_features_are_refused_loudly():
    model, loss_fn, x, t = make_case("small_cnn_ce", 4, 0)
    bad = nn.Sequential(nn.Conv2d(3, 4, 3, groups=1, dilation=2), nn.AdaptiveAvgPool2d(1), nn.Flatten())
    with pytest.raises(NotImplementedError):
        lower_module(bad, loss_fn, list(bad.parameters()), input_shape=(3, 8, 8))
