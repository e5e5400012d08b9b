"""GPU: loss, gradient, GGN-/Hessian-vector products and the Fisher diagonal of the sm_100a kernels against
the fixtures minted from the reference (tests/golden/matvec.pt) -- tolerance rtol 1e-4 (north_star)."""
import copy

import pytest
import torch

import hf_oracle as O
from helpers import GOLDEN, SPECS, build_loss, build_model, make_data

from pytorchhessianfree_b200 import HessianFree, diag_EF_autograd, diag_EF_backpack
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem

pytestmark = pytest.mark.gpu
DEV = "cuda"
MV = torch.load(f"{GOLDEN}/matvec.pt", weights_only=False)
RTOL = 1e-4


def close(got, want, what, rtol=RTOL):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    atol = 1e-5 * max(want.abs().max().item(), 1e-30)
    assert torch.allclose(got, want, rtol=rtol, atol=atol), \
        f"{what}: max abs err {(got - want).abs().max().item():.3e} (scale {want.abs().max().item():.3e})"


@pytest.mark.parametrize("i", range(len(MV)))
@pytest.mark.parametrize("front", ["graph", "module"])
def test_against_reference_fixture(i, front):
    c = MV[i]
    spec = SPECS[c["net"]]
    model = build_model(spec)
    model.load_state_dict(c["state"])
    model.to(DEV)
    loss_fn = build_loss(spec, c["reduction"])
    x, t, v = c["x"].to(DEV), c["t"].to(DEV), c["v"].to(DEV)
    params = [p for p in model.parameters() if p.requires_grad]
    if front == "graph":  # the path step(forward) takes: structure recovered from the autograd graph
        out = model(x)
        loss = loss_fn(out, t)
        close(HessianFree._Gv(loss, out, params, v), c["Gv"], "Gv")
        close(HessianFree._Hv(loss, params, v), c["Hv"], "Hv")
    else:  # the path acc_step / get_preconditioner take
        prog = lower_module(model, loss_fn, params)
        theta = torch.cat([p.detach().reshape(-1) for p in params])
        for curv, key in (("ggn", "Gv"), ("hessian", "Hv")):
            prob = NativeProblem(NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params), theta, curv, [(x, t)])
            close(prob.linearize().float(), c["loss"], "loss", rtol=1e-5)
            close(prob.gradient(), c["grad"], "gradient")
            close(prob.mvp(v), c[key], key)
            if "Gv_dense64" in c:  # explicit float64 J^T H J / Hessian known answer
                close(prob.mvp(v), c[key + "_dense64"], key + " vs dense float64")
        close(diag_EF_backpack(model, loss_fn, x, t, c["reduction"]), c["ef"], "Fisher diagonal")
        close(diag_EF_autograd(model, loss_fn, x, t, c["reduction"]), c["ef"], "Fisher diagonal")


@pytest.mark.parametrize("name", ["mlp_ce", "ae_bce", "small_nn"])
@pytest.mark.parametrize("curv", ["ggn", "hessian"])
def test_chunked_equals_full_batch(name, curv):
    """Sharding equivalence (reference tests/test_optimizer_acc.py): ragged chunks == concatenated batch."""
    spec = SPECS[name]
    torch.manual_seed(3)
    model = build_model(spec).to(DEV)
    loss_fn = build_loss(spec, "mean")
    x, t = (a.to(DEV) for a in make_data(spec, 37, 5))
    params = [p for p in model.parameters() if p.requires_grad]
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params])
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params)
    v = torch.randn_like(theta)
    full = NativeProblem(net, theta, curv, [(x, t)])
    parts = NativeProblem(net, theta, curv, [(x[:7], t[:7]), (x[7:8], t[7:8]), (x[8:], t[8:])])
    close(parts.linearize(), full.linearize(), "loss", rtol=1e-5)
    close(parts.gradient(), full.gradient(), "gradient")
    close(parts.mvp(v), full.mvp(v), "mvp")
    close(parts.fisher_diag(), full.fisher_diag(), "fisher")


@pytest.mark.parametrize("widths,loss,n", [([784, 512, 512, 10], "ce", 512), ([784, 1000, 500, 250, 30, 250, 500, 1000, 784], "bce", 300),
                                            ([33, 130, 67, 9], "mse", 257)])
@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_wide_layers_against_oracle(widths, loss, n, engine):
    """BASELINE.json configs[1]/[2] layer shapes (reduced batch) and a ragged-width net, against the CPU oracle."""
    spec = dict(widths=widths, act="relu" if loss == "ce" else "sigmoid", bias=[True] * (len(widths) - 1), frozen=[],
                loss=loss, linear_after=[3] if loss == "bce" else [])
    loss_fn = build_loss(spec, "mean")
    for seed in range(20):
        # A ReLU whose pre-activation sits within rounding of 0 flips its mask between any two float32
        # implementations (torch-CPU vs torch-CUDA differ there too); draw until no unit is that close.
        torch.manual_seed(seed)
        ref = build_model(spec)
        x, t = make_data(spec, n, 11 + seed)
        h, margin = x.double(), float("inf")
        for m in copy.deepcopy(ref).double():
            if isinstance(m, torch.nn.ReLU):
                margin = min(margin, h.abs().min().item())
            h = m(h)
        if margin > 1e-5:
            break
    params = [p for p in ref.parameters() if p.requires_grad]
    out = ref(x)
    l = loss_fn(out, t)
    v = torch.randn(sum(p.numel() for p in params))
    want_g = O.flatten(torch.autograd.grad(l, params, retain_graph=True))
    want_G, want_H = O.Gv(l, out, params, v), O.Hv(l, params, v)
    model = copy.deepcopy(ref).to(DEV)
    dparams = [p for p in model.parameters() if p.requires_grad]
    prog = lower_module(model, loss_fn, dparams)
    theta = torch.cat([p.detach().reshape(-1) for p in dparams])
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    for curv, want in (("ggn", want_G), ("hessian", want_H)):
        prob = NativeProblem(net, theta, curv, [(x.to(DEV), t.to(DEV))])
        close(prob.linearize().float(), l.detach(), "loss", rtol=1e-5)
        close(prob.gradient(), want_g, "gradient")
        close(prob.mvp(v.to(DEV)), want, curv)


@pytest.mark.parametrize("loss,classes", [("ce", 10), ("mse", 3), ("bce", 32)])
def test_fused_output_head_many_rows(loss, classes):
    """The fused output head (csrc/head.cuh) with more rows than 32 x SM count, so that every CTA takes several
    32-row steps and carries its head gradient across them; ragged widths and row count; against float64 autograd."""
    n = 32 * 148 * 2 + 77
    spec = dict(widths=[21, 70, classes], act="tanh", bias=[True, True], frozen=[], loss=loss)
    loss_fn = build_loss(spec, "mean")
    torch.manual_seed(3)
    ref = build_model(spec)
    x, t = make_data(spec, n, 5)
    m64 = copy.deepcopy(ref).double()
    p64 = list(m64.parameters())
    out = m64(x.double())
    l = loss_fn(out, t.double() if t.is_floating_point() else t)
    v = torch.randn(sum(p.numel() for p in p64))
    want = O.Gv(l, out, p64, v.double())
    model = copy.deepcopy(ref).to(DEV)
    dparams = list(model.parameters())
    prog = lower_module(model, loss_fn, dparams)
    theta = torch.cat([p.detach().reshape(-1) for p in dparams])
    for engine in ("simt", "tc"):
        net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
        prob = NativeProblem(net, theta, "ggn", [(x.to(DEV), t.to(DEV))])
        prob.linearize()
        prob.gradient()
        close(prob.mvp(v.to(DEV)), want, f"ggn ({engine})")
        # two chunks (one of them a single ragged step) accumulate into the same vector
        parts = NativeProblem(net, theta, "ggn", [(x[:5000].to(DEV), t[:5000].to(DEV)), (x[5000:].to(DEV), t[5000:].to(DEV))])
        parts.linearize()
        parts.gradient()
        close(parts.mvp(v.to(DEV)), want, f"ggn chunked ({engine})")


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_linearity_and_symmetry_at_full_width(engine):
    """Size-independent properties at BASELINE configs[1] full size: B(av+bw) = aBv+bBw, v.Bw = w.Bv, v.Bv >= 0."""
    spec = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
    torch.manual_seed(0)
    model = build_model(spec).to(DEV)
    loss_fn = build_loss(spec, "mean")
    x, t = (a.to(DEV) for a in make_data(spec, 4096, 0))
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params])
    prob = NativeProblem(NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine), theta, "ggn",
                         [(x, t)])
    prob.linearize()
    v, w = torch.randn_like(theta), torch.randn_like(theta)
    Bv, Bw = prob.mvp(v), prob.mvp(w)
    close(prob.mvp(2.0 * v - 0.5 * w), 2.0 * Bv - 0.5 * Bw, "linearity")
    a, b = torch.dot(v.double(), Bw.double()).item(), torch.dot(w.double(), Bv.double()).item()
    assert abs(a - b) <= 1e-4 * max(abs(a), abs(b), 1e-12)
    assert torch.dot(v.double(), Bv.double()).item() >= 0.0


@pytest.mark.parametrize("curv", ["ggn", "hessian"])
@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_two_phase_product_equals_single_call(curv, engine):
    """hf_matvec_phase (the data-parallel overlap seam) must reproduce hf_ggn_matvec / hf_hessian_matvec bit for bit."""
    spec = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
    torch.manual_seed(0)
    model = build_model(spec).to(DEV)
    loss_fn = build_loss(spec, "mean")
    x, t = (a.to(DEV) for a in make_data(spec, 1024, 0))
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params])
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    prob = NativeProblem(net, theta, curv, [(x, t)])
    prob.linearize(), prob.gradient()
    v = torch.randn_like(theta)
    want = prob.mvp(v)
    off, cnt = net.first_layer_span()
    assert (off, cnt) == (0, 784 * 512 + 512)
    got = torch.full_like(theta, float("nan"))
    lin = prob.mvp_lins[0]
    lin.matvec_phase(curv, theta, v, got, 0)
    torch.cuda.synchronize()
    assert torch.isnan(got[:cnt]).all() and torch.equal(got[cnt:], want[cnt:])
    lin.matvec_phase(curv, theta, v, got, 1)
    assert torch.equal(got, want)
