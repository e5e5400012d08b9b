import sys, torch
sys.path[:0] = ['.']
import bench
from pytorchhessianfree_b200 import DiagonalPreconditioner, pcg_device
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
dev = torch.device('cuda')
model = bench.build_mlp(0).to(dev); loss_fn = torch.nn.CrossEntropyLoss()
params = list(model.parameters()); prog = lower_module(model, loss_fn, params)
theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params)
x, t = bench.synth(1); x, t = x.to(dev), t.to(dev)
prob = NativeProblem(net, theta, "ggn", [(x, t)]); print("loss", prob.linearize().item())
g = prob.gradient(); d = prob.fisher_diag(); M = DiagonalPreconditioner(d, 1.0)
print("g norm", g.norm().item(), "fisher min/max", d.min().item(), d.max().item(), "minv", M.minv.min().item(), M.minv.max().item())
for K in (5, 20, 50):
    xs, ms, why = pcg_device(prob.matvec, -g, minv=M.minv, damping=1.0, max_iter=K, tol=0.0, martens_conv_crit=True, store_x_at_iters=None, poll=K)
    print(K, why, len(xs), [float(m) for m in ms][-3:])
    from pytorchhessianfree_b200.cg import _Solver
