"""GPU: the fused PCG kernel against the reference's own cg tests (tests/test_cg.py, run on CUDA), the
golden fixtures minted from the reference, and the CPU oracle."""
import warnings

import pytest
import torch

import hf_oracle as O
from helpers import GOLDEN, spd_system

from pytorchhessianfree_b200 import cg, diag_to_preconditioner, pcg_device

pytestmark = pytest.mark.gpu
DEV = "cuda"
SEEDS, DIMS = [0, 1, 42], [3, 10, 50]
CG = torch.load(f"{GOLDEN}/cg.pt", weights_only=False)
EPS = 5e-6  # incremental vs. true residual, as in the reference test


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("tol", [1e-3, 1e-6])
@pytest.mark.parametrize("atol", [1e-3, 1e-6])
@pytest.mark.parametrize("precond", [True, False])
def test_cg_residuals(seed, dim, tol, atol, precond):  # reference tests/test_cg.py:34-87
    A, b, _ = spd_system(dim, seed)
    A, b = A.to(DEV), b.to(DEV)
    Minv = torch.diag(torch.diag(A) ** -1)
    M = (lambda v: Minv @ v) if precond else None
    xs, _, why = cg(lambda v: A @ v, b, M=M, max_iter=10 * dim, tol=tol, atol=atol)
    assert len(xs) - 1 <= 10 * dim
    res = torch.linalg.norm(A @ xs[-1] - b).item()
    if why == "Convergence (tolerances)":
        assert res <= max(tol * torch.linalg.norm(b).item(), atol) + EPS


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("x0_none", [True, False])
@pytest.mark.parametrize("precond", [True, False])
def test_cg_m_iters(seed, dim, x0_none, precond):  # reference tests/test_cg.py:98-156, atol 1e-7
    A, b, _ = spd_system(dim, seed)
    x0 = None if x0_none else (2 * (torch.rand(dim) - 0.5)).to(DEV)
    A, b = A.to(DEV), b.to(DEV)
    Minv = torch.diag(torch.diag(A) ** -1)
    M = (lambda v: Minv @ v) if precond else None
    xs, ms, _ = cg(lambda v: A @ v, b, x0=x0, M=M, max_iter=10 * dim, tol=1e-6, atol=1e-6, martens_conv_crit=True,
                   store_x_at_iters=list(range(10 * dim)))
    quad = torch.stack([0.5 * torch.dot(x, A @ x) - torch.dot(b, x) for x in xs]).cpu()
    assert torch.allclose(quad, torch.stack(ms).cpu(), atol=1e-7)


@pytest.mark.parametrize("seed", SEEDS)
@pytest.mark.parametrize("dim", DIMS)
def test_pcg(seed, dim):  # reference tests/test_cg.py:159-224 (float64, bit-identical None vs identity)
    A, b, _ = spd_system(dim, seed)
    A, b = A.double().to(DEV), b.double().to(DEV)
    Ainv = torch.linalg.inv(A)
    runs = []
    for M in (None, lambda v: v, lambda v: Ainv @ v):
        xs, _, _ = cg(lambda v: A @ v, b, M=M, max_iter=10 * dim, tol=1e-6, atol=1e-6,
                      store_x_at_iters=list(range(10 * dim)))
        runs.append(xs)
    assert len(runs[0]) == len(runs[1])
    for u, v in zip(runs[0], runs[1]):
        assert torch.equal(u, v), "`None` and `identity` don't yield the same result"
    assert len(runs[2]) - 1 <= 1


@pytest.mark.parametrize("i", range(len(CG["cases"])))
def test_cg_against_reference_fixture(i):
    """Iterates, quadratic values, iteration count and stopping reason of the reference itself."""
    c = CG["cases"][i]
    f64 = "float64" in c["dtype"]
    A, b = c["A"].to(DEV), c["b"].to(DEV)
    x0 = None if c["x0"] is None else c["x0"].to(DEV)
    M = None
    if c["precond"]:
        dinv = c["dinv"].to(DEV)
        M = lambda v: dinv * v  # noqa: E731
    n = 10 * c["dim"]
    xs, ms, why = cg(lambda v: A @ v, b, x0=x0, M=M, max_iter=n, tol=1e-6, atol=1e-6, martens_conv_crit=True,
                     store_x_at_iters=list(range(n)))
    want_x, want_m = c["x_iters"], c["m_iters"]
    # CG is chaotic once the residual is at rounding level: compare the well-conditioned head exactly-ish and
    # the end point through the residual (north_star: CG iterates rtol 1e-3)
    k = min(len(xs), len(want_x), 4)
    scale = want_x.abs().max().item()
    tol = 1e-9 if f64 else 1e-3
    for j in range(k):
        assert torch.allclose(xs[j].cpu(), want_x[j], rtol=tol, atol=tol * scale), f"iterate {j}"
        assert torch.allclose(ms[j].cpu(), want_m[j], rtol=tol, atol=tol * abs(want_m[-1].item()) + 1e-12)
    if f64:
        # float64: same stopping reason and iteration count; iterates agree until the residual reaches rounding level
        assert why == c["reason"] and abs(len(xs) - len(want_x)) <= 1
        n_cmp = min(len(xs), len(want_x))
        assert torch.allclose(torch.stack(xs[:n_cmp]).cpu(), want_x[:n_cmp], rtol=1e-3, atol=1e-3 * scale)
    else:
        assert abs(len(xs) - len(want_x)) <= max(3, len(want_x) // 4)
        r_ref = torch.linalg.norm(c["A"] @ want_x[-1] - c["b"]).item()
        r_got = torch.linalg.norm(c["A"] @ xs[-1].cpu() - c["b"]).item()
        assert r_got <= 10 * max(r_ref, 1e-6 * torch.linalg.norm(c["b"]).item())


def test_martens_criterion_and_grid_against_reference_fixture():
    c = CG["martens"]
    A, b = c["A"].to(DEV), c["b"].to(DEV)
    for solver in ("generic", "device"):
        if solver == "generic":
            xs, ms, why = cg(lambda v: A @ v, b, max_iter=250, martens_conv_crit=True, store_x_at_iters=None, tol=1e-10)
        else:
            def mv(v, out, skip):
                torch.mv(A, v, out=out)
            xs, ms, why = pcg_device(mv, b, max_iter=250, martens_conv_crit=True, store_x_at_iters=None, tol=1e-10,
                                     poll=4)
        assert why == c["reason"] == "Convergence (Martens)"
        assert abs((len(xs) - 1) - (len(c["m_iters"]) - 1)) <= 2
        idx = [i for i, x in enumerate(xs) if x is not None]
        common = [i for i in idx if i in c["idx"].tolist()]
        assert common[:10] == c["idx"].tolist()[:10]
        for i in common[:8]:
            want = c["x"][c["idx"].tolist().index(i)]
            assert torch.allclose(xs[i].cpu(), want, rtol=1e-3, atol=1e-3 * want.abs().max().item())
        n = min(len(ms), len(c["m_iters"]))
        assert torch.allclose(torch.stack(ms)[:n].cpu(), c["m_iters"][:n], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("P", [1, 5, 1000, 4097, 1 << 20, 3_000_001])
@pytest.mark.parametrize("precond", [False, True])
def test_fused_kernel_matches_oracle_on_diagonal_systems(P, precond):
    """Sizes that exercise one CTA, ragged tails, the resident and the streaming (non-resident) paths."""
    g = torch.Generator().manual_seed(P)
    d = torch.rand(P, generator=g) + 0.5
    b = torch.randn(P, generator=g)
    lam = 0.3
    pre = torch.rand(P, generator=g) + 0.1
    M_o = O.diag_precond(pre, lam) if precond else None
    xs_o, ms_o, why_o = O.pcg(lambda v: d * v + lam * v, b, M=M_o, max_iter=12, martens_conv_crit=True,
                              store_x_at_iters=[0, 3, 12], tol=1e-12)
    dd, bd = d.to(DEV), b.to(DEV)
    minv = diag_to_preconditioner(pre.to(DEV), lam).minv if precond else None

    def mv(v, out, skip):
        torch.mul(dd, v, out=out)
    xs, ms, why = pcg_device(mv, bd, minv=minv, damping=lam, max_iter=12, martens_conv_crit=True,
                             store_x_at_iters=[0, 3, 12], tol=1e-12, poll=5)
    assert why == why_o and len(xs) == len(xs_o)
    assert [x is None for x in xs] == [x is None for x in xs_o]
    for u, v in zip(xs, xs_o):
        if u is not None:
            assert torch.allclose(u.cpu(), v, rtol=1e-3, atol=1e-5 * v.abs().max().item())
    assert torch.allclose(torch.stack(ms).cpu(), torch.stack(ms_o), rtol=1e-3, atol=1e-6 * abs(ms_o[-1].item()))


def test_nonpositive_curvature_warns_like_the_reference():
    A = torch.diag(torch.tensor([1.0, -2.0, 3.0])).to(DEV)
    b = torch.tensor([1.0, 1.0, 1.0], device=DEV)
    with pytest.warns(UserWarning, match="Directional curvature pAp"):
        cg(lambda v: A @ v, b, max_iter=3)


def test_divergence_and_maxiter_reasons():
    b = torch.ones(8, device=DEV)
    _, _, why = cg(lambda v: v * float("nan"), b, max_iter=5)
    assert why == "Divergence"
    A = torch.diag(torch.linspace(1, 100, 8)).to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xs, _, why = cg(lambda v: A @ v, b, max_iter=2, tol=1e-12)
    assert why == "Number of iterations" and len(xs) == 3
