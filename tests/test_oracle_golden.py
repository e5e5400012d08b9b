"""CPU: the travelling oracle (oracle/hf_oracle.py) against the fixtures minted from the
unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import copy
import warnings

import pytest
import torch

import hf_oracle as O
from helpers import GOLDEN, SPECS, build_loss, build_model, spd_system  # noqa: F401

CG = torch.load(f"{GOLDEN}/cg.pt", weights_only=False)
MV = torch.load(f"{GOLDEN}/matvec.pt", weights_only=False)
ST = torch.load(f"{GOLDEN}/steps.pt", weights_only=False)
SEL = torch.load(f"{GOLDEN}/selection.pt", weights_only=False)


@pytest.mark.parametrize("i", range(0, len(CG["cases"]), 3))
def test_pcg_matches_reference(i):
    c = CG["cases"][i]
    A, b = c["A"], c["b"]
    M = (lambda v: c["dinv"] * v) if c["precond"] else None
    n = 10 * c["dim"]
    xs, ms, why = O.pcg(lambda v: A @ v, b, x0=c["x0"], M=M, max_iter=n, tol=1e-6, atol=1e-6,
                        martens_conv_crit=True, store_x_at_iters=list(range(n)))
    assert why == c["reason"] and len(xs) == len(c["x_iters"])
    assert torch.allclose(torch.stack(xs), c["x_iters"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(torch.stack(ms), c["m_iters"], rtol=1e-5, atol=1e-7)


def test_storing_grid_matches_reference():
    for m, grid in CG["grids"].items():
        assert O.storing_grid(m) == grid
    assert O.storing_grid(250) == [0, 1, 2, 3, 4, 6, 8, 10, 13, 17, 23, 30, 39, 51, 66, 86, 112, 146, 190, 247, 321]


def test_martens_criterion_fires():
    c = CG["martens"]
    xs, ms, why = O.pcg(lambda v: c["A"] @ v, c["b"], max_iter=250, martens_conv_crit=True, store_x_at_iters=None, tol=1e-10)
    assert why == c["reason"] == O.REASON_MARTENS
    idx = [i for i, x in enumerate(xs) if x is not None]
    assert idx == c["idx"].tolist()
    assert torch.allclose(torch.stack(ms), c["m_iters"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("i", range(len(MV)))
def test_matvec_matches_reference(i):
    c = MV[i]
    spec = SPECS[c["net"]]
    model = build_model(spec)
    model.load_state_dict(c["state"])
    loss_fn = build_loss(spec, c["reduction"])
    params = [p for p in model.parameters() if p.requires_grad]
    out = model(c["x"])
    loss = loss_fn(out, c["t"])
    assert torch.allclose(loss, c["loss"], rtol=1e-6)
    assert torch.allclose(O.Gv(loss, out, params, c["v"]), c["Gv"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(O.Hv(loss, params, c["v"]), c["Hv"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(O.ef_diag(model, loss_fn, c["x"], c["t"], c["reduction"]), c["ef"], rtol=1e-5, atol=1e-9)
    # the one-pass (d^2)^T (a^2) form used as the oracle at benchmark batch sizes, against the reference's per-sample values
    assert torch.allclose(O.ef_diag_layerwise(model, loss_fn, c["x"], c["t"], c["reduction"]), c["ef"], rtol=1e-4, atol=1e-9)
    if "Gv_dense64" in c:  # explicit J^T H J known answer (float64)
        assert torch.allclose(c["Gv"].double(), c["Gv_dense64"], rtol=1e-4, atol=1e-6)
        assert torch.allclose(c["Hv"].double(), c["Hv_dense64"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("i", range(0, len(ST), 2))
def test_step_trajectory_matches_reference(i):
    c = ST[i]
    spec = SPECS[c["net"]]
    model = build_model(spec)
    model.load_state_dict(c["init_state"])
    loss_fn = build_loss(spec, c["reduction"])
    orc = O.OracleHF(model.parameters(), curvature_opt=c["curv"], **c["hf_kw"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for x, t in c["data"]:
            M = None
            if c["precond"]:
                M = O.diag_precond(O.ef_diag(model, loss_fn, x, t, c["reduction"]), orc.damping)
            if c["chunks"] is None:
                orc.step(lambda: (lambda o: (loss_fn(o, t), o))(model(x)), M_func=M)
            else:
                dl, off = [], 0
                for n in c["chunks"]:
                    dl.append((x[off:off + n], t[off:off + n]))
                    off += n
                orc.acc_step(model, loss_fn, dl, M_func=M, reduction=c["reduction"])
    assert orc.log["cg_reasons"] == c["cg_reasons"]
    assert orc.log["num_cg_iters"] == c["num_cg_iters"]
    assert torch.allclose(torch.tensor(orc.log["init_losses"]), torch.tensor(c["init_losses"]), rtol=1e-4)
    assert torch.allclose(torch.tensor(orc.log["dampings"]), torch.tensor(c["dampings"]), rtol=1e-6)
    for k, w in model.state_dict().items():
        assert torch.allclose(w, c["final_state"][k], rtol=1e-4, atol=1e-5)


def test_selection_matches_reference():
    assert O.backtrack_all(lambda s: s, SEL["toy"]) == SEL["toy_all"] == (1, 1.0)
    assert O.backtrack_efficient(lambda s: s, SEL["toy"]) == SEL["toy_eff"] == (4, 2.4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for c in SEL["linesearch"]:
            f = lambda s: float(0.5 * s @ c["A"] @ s - c["b"] @ s)  # noqa: E731
            a, fa = O.armijo(f, -c["b"], c["step"])
            assert a == pytest.approx(c["alpha"]) and fa == pytest.approx(c["f"], rel=1e-5)
    p = SEL["precond"]
    assert torch.allclose(O.diag_precond(p["d"], p["damping"], p["exponent"])(p["v"]), p["out"])


BS = torch.load(f"{GOLDEN}/benchsize.pt", weights_only=False)


@pytest.mark.parametrize("i", range(len(BS["cases"])))
def test_benchsize_fixture_inputs_regenerate(i):
    """The benchmark-size fixtures store strided samples of the reference's results; inputs and weights are re-created
    from seeds (a 4096 x 784 batch does not belong in git).  Check here, on CPU, that the seeded re-creation still lands
    on the tensors the fixture was minted from; the GPU test relies on it."""
    from helpers import benchsize_problem

    c = BS["cases"][i]
    model, loss_fn, x, t, v = benchsize_problem(c["cfg"], c["seed"])
    assert float(x.double().sum()) == pytest.approx(c["x_sum"], rel=1e-12)
    assert float(v.double().sum()) == pytest.approx(c["v_sum"], rel=1e-12)
    w = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert float(w.double().sum()) == pytest.approx(c["w_sum"], rel=1e-12)


def test_benchsize_oracle_reproduces_one_fixture():
    """One benchmark-size case end to end on CPU (the MLP at batch 4096, seed 0): oracle vs the stored reference sample."""
    from helpers import benchsize_problem

    c = next(c for c in BS["cases"] if c["cfg"] == "cfg2" and c["seed"] == 0)
    model, loss_fn, x, t, v = benchsize_problem("cfg2", 0)
    params = list(model.parameters())
    out = model(x)
    loss = loss_fn(out, t)
    idx = torch.arange(0, v.numel(), BS["stride"])
    assert torch.allclose(O.Gv(loss, out, params, v)[idx], c["Gv"], rtol=1e-4, atol=1e-8)
    assert torch.allclose(O.ef_diag_layerwise(model, loss_fn, x, t, "mean")[idx], c["ef"], rtol=1e-4, atol=1e-12)


CV = torch.load(f"{GOLDEN}/conv.pt", weights_only=False)


@pytest.mark.parametrize("i", range(len(CV)))
def test_conv_matvec_matches_reference(i):
    """The oracle on the small CNNs against what the unmodified reference returned for them, and the reference's float32
    products against float64 dense known answers."""
    from helpers import conv_fixture_case

    c = CV[i]
    model, loss_fn, x, t, v = conv_fixture_case(c)
    params = list(model.parameters())
    out = model(x)
    loss = loss_fn(out, t)
    assert torch.allclose(loss, c["loss"], rtol=1e-6)
    assert torch.allclose(O.flatten(torch.autograd.grad(loss, params, retain_graph=True)), c["grad"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(O.Gv(loss, out, params, v), c["Gv"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(O.Hv(loss, params, v), c["Hv"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(c["Gv"].double(), c["Gv_dense64"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(c["Hv"].double(), c["Hv_dense64"], rtol=1e-4, atol=1e-6)
