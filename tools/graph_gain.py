"""Steady-state time per CG iteration with plain launches vs CUDA-graph replay (configs[1], 250-iteration solves)."""
import sys, time, torch
sys.path[:0] = ['tests', '.']
from helpers import build_model, build_loss
from pytorchhessianfree_b200.cg import DiagonalPreconditioner, pcg_device
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
from torch.nn.utils import parameters_to_vector
DEV = 'cuda'
MLP = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
torch.manual_seed(0)
model = build_model(MLP).to(DEV); loss_fn = build_loss(MLP, "mean")
x, t = torch.rand(4096, 784, device=DEV), torch.randint(0, 10, (4096,), device=DEV)
params = list(model.parameters())
prog = lower_module(model, loss_fn, params)
theta = parameters_to_vector(params).detach().clone()
net = NativeNet(prog.layers, prog.loss, prog.reduction, theta.numel(), engine="tc")
prob = NativeProblem(net, theta, "ggn", [(x, t)])
prob.linearize(); g = prob.gradient(); M = DiagonalPreconditioner(prob.fisher_diag(), 1e-3)
for K in (50, 250):
    for graph in (False, True):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            xs, ms, why = pcg_device(prob.matvec, -g, minv=M.minv, damping=1e-3, max_iter=K, tol=0.0, martens_conv_crit=False,
                                     store_x_at_iters=[0], poll=K, use_graph=graph)
            torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
        print(f"K={K} graph={graph}: solve ms {[round(v, 2) for v in ts]}  -> {1e3 * min(ts) / K:.1f} us/iteration  ({why}, {len(xs) - 1} its)")
