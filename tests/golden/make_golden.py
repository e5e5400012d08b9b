"""Mint the golden fixtures from the UNMODIFIED reference  --  run in the build container only.

    python tests/golden/make_golden.py

Imports ``hessianfree`` from ``/root/reference`` (read-only) with
``oracle/backpack_shim`` standing in for the absent BackPACK package, runs the
reference's own ``cg``, ``HessianFree._Gv/_Hv``, ``diag_EF_*``,
``HessianFree.step/acc_step`` on seeded inputs, ASSERTS that the travelling
oracle (``oracle/hf_oracle.py``) reproduces every result, and writes the
results as small ``.pt`` fixtures next to this file.  The GPU box has no
``/root/reference``; there the tests compare the CUDA path with these files
and with the oracle.
"""
import copy
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"),
                os.path.join(ROOT, "oracle", "backpack_shim"), "/root/reference"]

import hf_oracle as O  # noqa: E402
from helpers import SPECS, build_loss, build_model, make_data, spd_system  # noqa: E402

from hessianfree.cg import cg as ref_cg  # noqa: E402
from hessianfree.cg_backtracking import cg_backtracking, cg_efficient_backtracking  # noqa: E402
from hessianfree.linesearch import simple_linesearch  # noqa: E402
from hessianfree.optimizer import HessianFree as RefHF  # noqa: E402
from hessianfree.preconditioners import (  # noqa: E402
    diag_EF_autograd, diag_EF_backpack, diag_EF_preconditioner, diag_to_preconditioner)

warnings.simplefilter("ignore")
SEEDS = [0, 1, 42]  # the reference's own seed set (tests/test_cg.py:9)


def same(a, b, what, tol=0.0):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if tol == 0.0:
        assert torch.equal(a, b), f"oracle != reference: {what}"
    else:
        assert torch.allclose(a, b, rtol=tol, atol=tol), f"oracle != reference: {what}"


def stack_opt(xs):
    """list with None entries -> (indices, stacked tensor)."""
    idx = [i for i, x in enumerate(xs) if x is not None]
    return torch.tensor(idx), torch.stack([xs[i] for i in idx])


# ---------------------------------------------------------------------------
def gold_cg():
    out = []
    for dtype in (torch.float32, torch.float64):
        for seed in SEEDS:
            for dim in (3, 10, 50):
                for precond in (False, True):
                    for x0_none in (True, False):
                        A, b, _ = spd_system(dim, seed, dtype)
                        x0 = None if x0_none else (2 * (torch.rand(dim) - 0.5)).to(dtype)
                        dinv = torch.diag(A) ** -1
                        M = (lambda v: dinv * v) if precond else None
                        kw = dict(x0=x0, M=M, max_iter=10 * dim, tol=1e-6, atol=1e-6, martens_conv_crit=True,
                                  store_x_at_iters=list(range(10 * dim)))
                        xs, ms, why = ref_cg(lambda v: A @ v, b, **kw)
                        xs2, ms2, why2 = O.pcg(lambda v: A @ v, b, **kw)
                        assert why == why2 and len(xs) == len(xs2)
                        for u, v in zip(xs, xs2):
                            same(u, v, "cg x_iters")
                        same(torch.stack(ms), torch.stack(ms2), "cg m_iters")
                        out.append(dict(dtype=str(dtype), seed=seed, dim=dim, precond=precond, A=A, b=b, x0=x0,
                                        dinv=dinv if precond else None, x_iters=torch.stack(xs),
                                        m_iters=torch.stack(ms), reason=why))
    # the automatic snapshot grid and the Martens criterion actually firing (neither is tested upstream)
    grids = {m: O.storing_grid(m) for m in (1, 4, 10, 50, 250, 1000)}
    A, b, _ = spd_system(200, 7)
    A = A + 0.05 * torch.eye(200)
    xs, ms, why = ref_cg(lambda v: A @ v, b, max_iter=250, martens_conv_crit=True, store_x_at_iters=None, tol=1e-10)
    xs2, ms2, why2 = O.pcg(lambda v: A @ v, b, max_iter=250, martens_conv_crit=True, store_x_at_iters=None, tol=1e-10)
    assert why == why2 == O.REASON_MARTENS, why
    idx, X = stack_opt(xs)
    idx2, X2 = stack_opt(xs2)
    same(idx, idx2, "grid idx"), same(X, X2, "grid x"), same(torch.stack(ms), torch.stack(ms2), "grid m")
    martens = dict(A=A, b=b, idx=idx, x=X, m_iters=torch.stack(ms), reason=why)
    torch.save(dict(cases=out, grids=grids, martens=martens), os.path.join(HERE, "cg.pt"))
    print(f"cg.pt: {len(out)} systems, grids {list(grids)}, martens case stops after {len(xs) - 1} iterations")


# ---------------------------------------------------------------------------
def gold_matvec():
    out = []
    for name, spec in SPECS.items():
        for seed in SEEDS:
            for reduction in ("mean", "sum"):
                for n in (1, 16):
                    torch.manual_seed(seed)
                    model = build_model(spec)
                    loss_fn = build_loss(spec, reduction)
                    x, t = make_data(spec, n, seed + 100)
                    params = [p for p in model.parameters() if p.requires_grad]
                    g = torch.Generator().manual_seed(seed + 200)
                    v = torch.randn(sum(p.numel() for p in params), generator=g)
                    outputs = model(x)
                    loss = loss_fn(outputs, t)
                    grad = O.flatten(torch.autograd.grad(loss, params, create_graph=True)).detach()
                    Gv = RefHF._Gv(loss, outputs, params, v)
                    Hv = RefHF._Hv(loss, params, v)
                    same(Gv, O.Gv(loss, outputs, params, v), "Gv")
                    same(Hv, O.Hv(loss, params, v), "Hv")
                    ef_a = diag_EF_autograd(model, loss_fn, x, t, reduction)
                    same(ef_a, O.ef_diag(model, loss_fn, x, t, reduction), "ef diag")
                    m2 = copy.deepcopy(model)
                    ef_b = diag_EF_backpack(m2, loss_fn, x, t, reduction)
                    same(ef_a, ef_b, "ef backpack-shim vs autograd", tol=1e-5)
                    # independent pin of `_Gv` (all seeds, all nets): float64 forward-mode GGN, not the R-op recipe
                    G64 = func_ggn64(model, loss_fn, x, t, v)
                    same(G64, Gv.double(), "func64 GGN vs Gv", tol=2e-5)
                    rec = dict(net=name, seed=seed, reduction=reduction, n=n, x=x, t=t, v=v,
                               state={k: w.detach().clone() for k, w in model.state_dict().items()},
                               loss=loss.detach(), grad=grad, Gv=Gv, Hv=Hv, ef=ef_a)
                    if n == 16 and seed == 0:
                        # float64 dense known answers (absent upstream): G = J^T H J and the full Hessian
                        m64 = copy.deepcopy(model).double()
                        t64 = t.double() if t.is_floating_point() else t
                        G = O.explicit_ggn(m64, loss_fn, x.double(), t64)
                        H = O.explicit_hessian(m64, loss_fn, x.double(), t64)
                        same(G, G.T, "GGN symmetric", tol=1e-12)
                        assert torch.linalg.eigvalsh(G).min() > -1e-10, "GGN must be p.s.d."
                        # only the products travel (the dense matrices would make the fixture MBs large)
                        rec["Gv_dense64"], rec["Hv_dense64"] = G @ v.double(), H @ v.double()
                        same(rec["Gv_dense64"], Gv.double(), "dense GGN vs Gv", tol=1e-5)
                        same(rec["Hv_dense64"], Hv.double(), "dense H vs Hv", tol=1e-5)
                    out.append(rec)
    torch.save(out, os.path.join(HERE, "matvec.pt"))
    print(f"matvec.pt: {len(out)} cases")


# ---------------------------------------------------------------------------
def func_ggn64(model, loss_fn, x, t, v):
    """G v = J^T H_loss J v in float64 by forward-mode AD (torch.func.jvp), an analytic-free loss Hessian
    (jvp of the loss gradient) and one vjp -- a derivation that shares nothing with the R-op recipe the oracle and
    the BackPACK shim restate, so it pins `_Gv` independently of them."""
    from torch.func import functional_call, grad, jvp, vjp

    m64 = copy.deepcopy(model).double()
    named = dict(m64.named_parameters())
    train = [k for k, p in named.items() if p.requires_grad]
    frozen = {k: p.detach() for k, p in named.items() if not p.requires_grad}
    x64 = x.double()
    t64 = t.double() if t.is_floating_point() else t
    plist = tuple(named[k].detach() for k in train)
    vlist = tuple(O.unflatten(v.double(), plist))

    def net(*ps):
        return functional_call(m64, {**frozen, **dict(zip(train, ps))}, (x64,))

    out, Jv = jvp(net, plist, vlist)
    _, HJv = jvp(grad(lambda z: loss_fn(z, t64)), (out,), (Jv,))
    _, pull = vjp(net, *plist)
    return O.flatten(pull(HJv))


def gold_benchsize(stride=997):
    """BASELINE.json configs[1] at batch 4096 and configs[2] at one 7 500-sample shard, the reference's seeds: gradient,
    `_Gv`, `_Hv`, `diag_EF_backpack` from the unmodified reference.  Inputs and weights are re-created from seeds by
    tests/helpers.py::benchsize_problem (their sums are stored to catch a drifting generator); of every P-vector only
    each 997th entry and the L2 norm travel."""
    from helpers import BENCH_CFGS, benchsize_problem

    cases = []
    for cfg in BENCH_CFGS:
        for seed in SEEDS:
            model, loss_fn, x, t, v = benchsize_problem(cfg, seed)
            params = list(model.parameters())
            outputs = model(x)
            loss = loss_fn(outputs, t)
            grad = O.flatten(torch.autograd.grad(loss, params, create_graph=True)).detach()
            Gv = RefHF._Gv(loss, outputs, params, v)
            Hv = RefHF._Hv(loss, params, v)
            same(Gv, O.Gv(loss, outputs, params, v), "benchsize Gv")
            same(Hv, O.Hv(loss, params, v), "benchsize Hv")
            ef = diag_EF_backpack(copy.deepcopy(model), loss_fn, x, t, "mean")
            same(ef, O.ef_diag_layerwise(model, loss_fn, x, t, "mean"), "benchsize ef", tol=1e-6)
            G64 = func_ggn64(model, loss_fn, x, t, v)
            rel = ((Gv.double() - G64).norm() / G64.norm()).item()
            # (float32 reference vs float64 truth: up to ~2e-5 at these sizes -- a few ReLU units whose pre-activation sits
            # within float32 rounding of zero take the other branch in float64; the stored value calibrates the GPU tolerance)
            assert rel < 1e-4, f"reference _Gv vs independent float64 GGN: {rel:.2e}"
            idx = torch.arange(0, v.numel(), stride)
            w = torch.cat([p.detach().reshape(-1) for p in params])
            cases.append(dict(cfg=cfg, seed=seed, n=x.shape[0], loss=loss.detach(), x_sum=float(x.double().sum()),
                              v_sum=float(v.double().sum()), w_sum=float(w.double().sum()),
                              grad=grad[idx].clone(), Gv=Gv[idx].clone(), Hv=Hv[idx].clone(), ef=ef[idx].clone(),
                              Gv_func64=G64[idx].clone(), grad_norm=float(grad.double().norm()), Gv_norm=float(Gv.double().norm()),
                              Hv_norm=float(Hv.double().norm()), ef_norm=float(ef.double().norm()), Gv_vs_func64=rel))
            print(f"benchsize {cfg} seed {seed}: loss {loss.item():.6f}, |Gv| {cases[-1]['Gv_norm']:.4e}, ref-vs-func64 {rel:.1e}")
    torch.save(dict(stride=stride, cases=cases), os.path.join(HERE, "benchsize.pt"))
    print(f"benchsize.pt: {len(cases)} cases")


def gold_conv():
    """The convolutional slice: the small CNNs of tests/test_gpu_conv.py (3x3 'same' / strided / 'valid', 1x1, bias on and
    off, global average pool, a Linear head; ReLU / tanh / sigmoid; CE / MSE / BCE), the reference's seeds, batch 1 and 6:
    loss, gradient, `_Gv`, `_Hv` from the unmodified reference + float64 dense `J^T H J v` and `H v` known answers."""
    from test_gpu_conv import CASES, make_case

    out = []
    for name in sorted(CASES):
        for seed in SEEDS:
            for n in (1, 6):
                model, loss_fn, x, t = make_case(name, n, seed)
                params = list(model.parameters())
                g = torch.Generator().manual_seed(seed + 200)
                v = torch.randn(sum(p.numel() for p in params), generator=g)
                outputs = model(x)
                loss = loss_fn(outputs, t)
                grad = O.flatten(torch.autograd.grad(loss, params, create_graph=True)).detach()
                Gv = RefHF._Gv(loss, outputs, params, v)
                Hv = RefHF._Hv(loss, params, v)
                same(Gv, O.Gv(loss, outputs, params, v), "conv Gv")
                same(Hv, O.Hv(loss, params, v), "conv Hv")
                # independent pin (no R-op recipe): float64 dense J^T H J and the full Hessian of these ~2 000-parameter nets
                m64 = copy.deepcopy(model).double()
                t64 = t.double() if t.is_floating_point() else t
                G = O.explicit_ggn(m64, loss_fn, x.double(), t64)
                H = O.explicit_hessian(m64, loss_fn, x.double(), t64)
                Gd, Hd = G @ v.double(), H @ v.double()
                same(Gd, Gv.double(), "conv dense GGN vs Gv", tol=2e-5)
                same(Hd, Hv.double(), "conv dense H vs Hv", tol=2e-5)
                out.append(dict(net=name, seed=seed, n=n, x=x, t=t, v=v,
                                state={k: w.detach().clone() for k, w in model.state_dict().items()},
                                loss=loss.detach(), grad=grad, Gv=Gv, Hv=Hv, Gv_dense64=Gd, Hv_dense64=Hd))
    torch.save(out, os.path.join(HERE, "conv.pt"))
    print(f"conv.pt: {len(out)} cases")


# ---------------------------------------------------------------------------
def run_ref_steps(name, seed, reduction, curv, n_steps, n, precond, chunks=None, **hf_kw):
    """n_steps of the reference optimizer; the oracle class must land on the same trajectory."""
    spec = SPECS[name]
    torch.manual_seed(seed)
    model = build_model(spec)
    init_state = {k: w.detach().clone() for k, w in model.state_dict().items()}
    twin = copy.deepcopy(model)
    loss_fn = build_loss(spec, reduction)
    opt = RefHF(model.parameters(), curvature_opt=curv, **hf_kw)
    orc = O.OracleHF(twin.parameters(), curvature_opt=curv, **hf_kw)
    data, finals = [], []
    for s in range(n_steps):
        x, t = make_data(spec, n, 1000 * seed + s)
        data.append((x, t))
        M = M2 = None
        if precond:
            # the reference's get_preconditioner drops its return value (optimizer.py:943); drive the
            # function it wraps so both sides really are preconditioned
            M = diag_EF_preconditioner(model, loss_fn, x, t, reduction, damping=opt._group["damping"], use_backpack=False)
            M2 = O.diag_precond(O.ef_diag(twin, loss_fn, x, t, reduction), orc.damping)
        if chunks is None:
            f1 = opt.step(lambda: (lambda o: (loss_fn(o, t), o))(model(x)), M_func=M)
            f2 = orc.step(lambda: (lambda o: (loss_fn(o, t), o))(twin(x)), M_func=M2)
        else:
            dl, off = [], 0
            for c in chunks:
                dl.append((x[off:off + c], t[off:off + c]))
                off += c
            f1 = opt.acc_step(model, loss_fn, dl, M_func=M, reduction=reduction)
            f2 = orc.acc_step(twin, loss_fn, dl, M_func=M2, reduction=reduction)
        if f1 is not None or f2 is not None:
            same(f1, f2, f"final loss step {s}", tol=1e-6)
        finals.append(f1)
    st = opt.state
    for k in ("init_losses", "dampings", "num_cg_iters", "learning_rates"):
        same(torch.tensor(st[k], dtype=torch.float64), torch.tensor(orc.log[k], dtype=torch.float64), k, tol=1e-6)
    assert st["cg_reasons"] == orc.log["cg_reasons"]
    if hf_kw.get("use_cg_backtracking", True):
        assert [int(b) for b in st["best_cg_iters"]] == [int(b) for b in orc.log["best_cg_iters"]]
    for p, q in zip(model.parameters(), twin.parameters()):
        same(p.data, q.data, "final params", tol=1e-6)
    # the same trajectory in float64: how far the reference itself moves under rounding (test tolerances)
    m64 = build_model(spec, torch.float64)
    m64.load_state_dict({k: w.double() for k, w in init_state.items()})
    o64 = O.OracleHF(m64.parameters(), curvature_opt=curv, **hf_kw)
    for x, t in data:
        x, t = x.double(), (t.double() if t.is_floating_point() else t)
        M64 = O.diag_precond(O.ef_diag(m64, loss_fn, x, t, reduction), o64.damping) if precond else None
        if chunks is None:
            o64.step(lambda: (lambda o: (loss_fn(o, t), o))(m64(x)), M_func=M64)
        else:
            dl, off = [], 0
            for c in chunks:
                dl.append((x[off:off + c], t[off:off + c]))
                off += c
            o64.acc_step(m64, loss_fn, dl, M_func=M64, reduction=reduction)
    return dict(net=name, seed=seed, reduction=reduction, curv=curv, n=n, precond=precond, chunks=chunks, hf_kw=hf_kw,
                init_losses64=list(o64.log["init_losses"]), num_cg_iters64=list(o64.log["num_cg_iters"]),
                final_state64={k: w.detach().clone() for k, w in m64.state_dict().items()},
                init_state=init_state, data=data, final_losses=finals,
                init_losses=list(st["init_losses"]), dampings=list(st["dampings"]), cg_reasons=list(st["cg_reasons"]),
                num_cg_iters=list(st["num_cg_iters"]), best_cg_iters=[int(b) for b in st["best_cg_iters"]],
                learning_rates=list(st["learning_rates"]),
                final_state={k: w.detach().clone() for k, w in model.state_dict().items()},
                x0=st["x0"].clone())


def gold_steps():
    out = []
    for seed in SEEDS:
        # BASELINE.json configs[0]: examples/run_mwe.py, 10 steps (north_star: loss trajectory over 10 steps)
        out.append(run_ref_steps("mwe", seed, "mean", "ggn", 10, 16, precond=False))
        out.append(run_ref_steps("mlp_ce", seed, "mean", "ggn", 10, 32, precond=True))
        out.append(run_ref_steps("ae_bce", seed, "mean", "ggn", 10, 32, precond=True))
        out.append(run_ref_steps("small_nn", seed, "mean", "hessian", 3, 16, precond=False))
        out.append(run_ref_steps("tanh_mse", seed, "sum", "hessian", 5, 16, precond=False))
        # reference tests/test_optimizer_acc.py::test_step: chunks [7,8], cg_max_iter=4
        for curv in ("ggn", "hessian"):
            for red in ("mean", "sum"):
                out.append(run_ref_steps("small_nn", seed, red, curv, 3, 15, precond=False, chunks=[7, 8], cg_max_iter=4))
    out.append(run_ref_steps("mwe", 0, "mean", "ggn", 4, 16, precond=False, use_cg_backtracking=False, use_linesearch=False,
                             adapt_damping=False, lr=0.5))
    torch.save(out, os.path.join(HERE, "steps.pt"))
    print(f"steps.pt: {len(out)} trajectories")


# ---------------------------------------------------------------------------
def gold_resume():
    """Checkpoint fidelity: the UNMODIFIED reference optimizer takes 3 steps, its
    ``state_dict()`` and the model travel in the fixture, and the reference's own next 3 steps are recorded.  The
    drop-in optimizer must load that state_dict and continue on the same trajectory (tests/test_gpu_steps.py)."""
    out = []
    for name, curv in (("mlp_ce", "ggn"), ("ae_bce", "ggn"), ("tanh_mse", "hessian")):
        spec = SPECS[name]
        torch.manual_seed(5)
        model = build_model(spec)
        loss_fn = build_loss(spec, "mean")
        opt = RefHF(model.parameters(), curvature_opt=curv, damping=0.7, cg_max_iter=30)
        data = [make_data(spec, 24, 300 + s) for s in range(6)]
        for x, t in data[:3]:
            opt.step(lambda: (lambda o: (loss_fn(o, t), o))(model(x)))
        blob = copy.deepcopy(opt.state_dict())
        mid_state = {k: w.detach().clone() for k, w in model.state_dict().items()}
        n_before = len(opt.state["init_losses"])
        for x, t in data[3:]:
            opt.step(lambda: (lambda o: (loss_fn(o, t), o))(model(x)))
        out.append(dict(net=name, curv=curv, optimizer_state_dict=blob, model_state=mid_state, data=data[3:],
                        init_losses=list(opt.state["init_losses"][n_before:]), dampings=list(opt.state["dampings"][n_before:]),
                        num_cg_iters=list(opt.state["num_cg_iters"][n_before:]), cg_reasons=list(opt.state["cg_reasons"][n_before:]),
                        final_state={k: w.detach().clone() for k, w in model.state_dict().items()}))
        assert isinstance(blob["state"]["x0"], torch.Tensor) and blob["param_groups"][0]["damping"] == opt.state["dampings"][n_before]
    torch.save(out, os.path.join(HERE, "resume.pt"))
    print(f"resume.pt: {len(out)} checkpoints taken from the reference optimizer")


# ---------------------------------------------------------------------------
def gold_selection():
    toy = [2.0, 1.0, None, 2.7, 2.4, None, None, 7.3]  # reference tests/test_cg_backtracking.py:8
    b1, v1 = cg_backtracking(lambda s: s, toy)
    b2, v2 = cg_efficient_backtracking(lambda s: s, toy)
    assert (int(b1), v1) == O.backtrack_all(lambda s: s, toy) and (b2, v2) == O.backtrack_efficient(lambda s: s, toy)
    ls = []
    for seed in SEEDS:
        A, b, _ = spd_system(10, seed)
        f = lambda s: float(0.5 * s @ A @ s - b @ s)  # noqa: E731
        for scale in (1.0, 30.0, -1.0):
            step = scale * torch.linalg.solve(A, b)
            r = simple_linesearch(f, -b, step, init_alpha=1.0)
            assert r == O.armijo(f, -b, step, init_alpha=1.0)
            ls.append(dict(A=A, b=b, step=step, alpha=r[0], f=r[1]))
    d = torch.rand(7)
    v = torch.randn(7)
    pre = dict(d=d, v=v, damping=0.3, exponent=0.75, out=diag_to_preconditioner(d, 0.3)(v))
    same(pre["out"], O.diag_precond(d, 0.3)(v), "diag precond")
    torch.save(dict(toy=toy, toy_all=(int(b1), v1), toy_eff=(b2, v2), linesearch=ls, precond=pre),
               os.path.join(HERE, "selection.pt"))
    print("selection.pt written")


if __name__ == "__main__":
    if "benchsize" in sys.argv:  # minutes of single-threaded autograd: minted on its own
        gold_benchsize()
        sys.exit(0)
    if "resume" in sys.argv:
        torch.set_num_threads(1)
        gold_resume()
        sys.exit(0)
    if "conv" in sys.argv:
        torch.set_num_threads(1)
        gold_conv()
        sys.exit(0)
    torch.set_num_threads(1)  # bit-stable reductions while minting
    gold_cg()
    gold_matvec()
    gold_steps()
    gold_resume()
    gold_selection()
    gold_conv()
