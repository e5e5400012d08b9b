import sys, copy, torch
sys.path[:0] = ['tests', 'oracle', '.']
import hf_oracle as O
from helpers import build_model, build_loss
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
DEV = 'cuda'
import os
if os.environ.get("NET") == "ae":
    MLP = dict(widths=[784, 1000, 500, 250, 30, 250, 500, 1000, 784], act="sigmoid", bias=[True] * 8, frozen=[], loss="bce", linear_after=[3])
else:
    MLP = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
NB = int(os.environ.get("NB", "4096"))
torch.manual_seed(0)
model = build_model(MLP).to(DEV); loss_fn = build_loss(MLP, "mean")
x = torch.rand(NB, 784, device=DEV)
t = x.clone() if MLP["loss"] == "bce" else torch.randint(0, 10, (NB,), device=DEV)
params = list(model.parameters())
prog = lower_module(model, loss_fn, params)
theta = torch.cat([p.detach().reshape(-1) for p in params])
v = torch.randn_like(theta)
m64 = copy.deepcopy(model).double(); p64 = list(m64.parameters())
out = m64(x.double()); loss = loss_fn(out, t.double() if t.is_floating_point() else t)
g64 = O.flatten(torch.autograd.grad(loss, p64, create_graph=True)).detach()
G64 = O.Gv(loss, out, p64, v.double())
res = {}
for engine in ("simt", "tc"):
    prob = NativeProblem(NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine), theta, "ggn", [(x, t)])
    prob.linearize(); res[engine] = (prob.gradient(), prob.mvp(v))
def report(name, got, want):
    off = 0
    for p in params:
        n = p.numel(); a = got[off:off+n].double(); b = want[off:off+n].double(); off += n
        print(f"   {name} {tuple(p.shape)}: l2rel {((a-b).norm()/b.norm()).item():.2e}  maxrel {((a-b).abs().max()/b.abs().max()).item():.2e}")
for engine in ("simt", "tc"):
    print(engine); report("grad", res[engine][0], g64); report("Gv  ", res[engine][1], G64)
