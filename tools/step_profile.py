"""Host-side profile (cProfile) of one optimizer step through the public API on configs[1], after warm-up."""
import cProfile, pstats, sys, time, warnings, torch
sys.path[:0] = ['tests', '.']
from helpers import build_model, build_loss
from pytorchhessianfree_b200 import HessianFree
DEV = 'cuda'
MLP = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
torch.manual_seed(0)
model = build_model(MLP).to(DEV); loss_fn = build_loss(MLP, "mean")
x, t = torch.rand(4096, 784, device=DEV), torch.randint(0, 10, (4096,), device=DEV)
opt = HessianFree(model.parameters())
warnings.simplefilter("ignore")
def one():
    M = opt.get_preconditioner(model, loss_fn, x, t, "mean")
    opt.acc_step(model, loss_fn, [(x, t)], M_func=M)
for _ in range(3):
    one()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable(); one(); torch.cuda.synchronize(); pr.disable()
print("step wall ms:", 1e3 * (time.perf_counter() - t0), "cg iters", opt.state["num_cg_iters"][-1])
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
