import sys, copy, torch
sys.path[:0] = ['tests', 'oracle', '.']
import hf_oracle as O
from helpers import build_model, build_loss, make_data
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
DEV = 'cuda'
for n in (128, 256, 512):
    spec = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
    torch.manual_seed(0)
    ref = build_model(spec); loss_fn = build_loss(spec, "mean")
    x, t = make_data(spec, n, 11)
    params = list(ref.parameters())
    l = loss_fn(ref(x), t)
    want = torch.autograd.grad(l, params)
    model = copy.deepcopy(ref).to(DEV)
    dparams = list(model.parameters())
    prog = lower_module(model, loss_fn, dparams)
    theta = torch.cat([p.detach().reshape(-1) for p in dparams])
    prob = NativeProblem(NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params), theta, "ggn", [(x.to(DEV), t.to(DEV))])
    print(n, "loss", prob.linearize().item(), l.item())
    g = prob.gradient().cpu()
    off = 0
    for p, w in zip(params, want):
        got = g[off:off + p.numel()].view_as(w); off += p.numel()
        err = (got - w).abs()
        idx = err.argmax().item()
        print("  ", tuple(w.shape), "max err %.3e scale %.3e" % (err.max().item(), w.abs().max().item()), "argmax", idx if w.dim()==1 else (idx // w.shape[1], idx % w.shape[1]),
              "n_bad", int((err > 1e-4 * w.abs().max()).sum()))
    # gpu torch reference too
    lg = loss_fn(model(x.to(DEV)), t.to(DEV)); wg = torch.autograd.grad(lg, dparams)
    print("   torch-gpu vs cpu:", max((a.cpu() - b).abs().max().item() for a, b in zip(wg, want)))
