"""Phase-by-phase wall time (synchronised) of one optimizer step on configs[1]: where a step's milliseconds go."""
import sys, time, warnings, torch
sys.path[:0] = ['tests', '.']
from helpers import build_model, build_loss
import pytorchhessianfree_b200.optimizer as O
import pytorchhessianfree_b200.problem as PR
from pytorchhessianfree_b200 import HessianFree
DEV = 'cuda'
MLP = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
torch.manual_seed(0)
model = build_model(MLP).to(DEV); loss_fn = build_loss(MLP, "mean")
x, t = torch.rand(4096, 784, device=DEV), torch.randint(0, 10, (4096,), device=DEV)
opt = HessianFree(model.parameters())
warnings.simplefilter("ignore")
acc = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc.setdefault(name, []).append(1e3 * (time.perf_counter() - t0))
        return out
    return w
O.lower_module = timed("lower_module", O.lower_module)
O.NativeProblem = timed("NativeProblem()", O.NativeProblem)
PR.NativeProblem.linearize = timed("linearize", PR.NativeProblem.linearize)
PR.NativeProblem.gradient = timed("gradient", PR.NativeProblem.gradient)
PR.NativeProblem.fisher_diag = timed("fisher_diag", PR.NativeProblem.fisher_diag)
PR.NativeProblem.losses_at = timed("losses_at", PR.NativeProblem.losses_at)
O.pcg_device = timed("pcg_device", O.pcg_device)
O.cg_efficient_backtracking = timed("backtracking(incl. losses)", O.cg_efficient_backtracking)
O.simple_linesearch = timed("linesearch(incl. losses)", O.simple_linesearch)
opt.get_preconditioner = timed("get_preconditioner(total)", opt.get_preconditioner)
def one():
    M = opt.get_preconditioner(model, loss_fn, x, t, "mean")
    opt.acc_step(model, loss_fn, [(x, t)], M_func=M)
for i in range(5):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    one()
    torch.cuda.synchronize(); total = 1e3 * (time.perf_counter() - t0)
print("step %.2f ms, cg iters %d" % (total, opt.state["num_cg_iters"][-1]))
for k, v in acc.items():
    print("  %-28s n=%d  total %.3f ms" % (k, len(v), sum(v)))
