import sys, torch
sys.path[:0] = ['tests', '.']
from helpers import build_model, build_loss
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
DEV = 'cuda'
AE = dict(widths=[784, 1000, 500, 250, 30, 250, 500, 1000, 784], act="sigmoid", bias=[True] * 8, frozen=[], loss="bce", linear_after=[3])
torch.manual_seed(0)
model = build_model(AE).to(DEV); loss_fn = build_loss(AE, "mean")
params = list(model.parameters()); prog = lower_module(model, loss_fn, params)
theta = torch.cat([p.detach().reshape(-1) for p in params])
net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine="tc")
x = torch.rand(7500, 784, device=DEV)
prob = NativeProblem(net, theta, "ggn", [(x, x)])
prob.linearize(); v = torch.randn_like(theta); out = torch.empty_like(theta)
for _ in range(3): prob.matvec(v, out)
torch.cuda.synchronize()
