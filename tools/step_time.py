"""Wall time of whole optimizer steps through the public API on BASELINE.json configs[1] (MLP, batch 4096)."""
import sys, time, warnings, torch
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
for mode in ("acc_step", "step"):
    times = []
    for i in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        M = opt.get_preconditioner(model, loss_fn, x, t, "mean")
        if mode == "acc_step":
            opt.acc_step(model, loss_fn, [(x, t)], M_func=M)
        else:
            opt.step(lambda: (lambda o: (loss_fn(o, t), o))(model(x)), M_func=M)
        torch.cuda.synchronize(); times.append(1e3 * (time.perf_counter() - t0))
    st = opt.state
    print(mode, "ms per step:", [round(v, 1) for v in times], "cg iters:", st["num_cg_iters"][-8:], "reasons:", set(st["cg_reasons"][-8:]))
