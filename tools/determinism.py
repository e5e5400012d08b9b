import sys, torch
sys.path[:0] = ['.']
import bench
from pytorchhessianfree_b200 import DiagonalPreconditioner, pcg_device
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
dev = torch.device('cuda')
model = bench.build_ae(0).to(dev); loss_fn = torch.nn.BCEWithLogitsLoss()
params = list(model.parameters()); prog = lower_module(model, loss_fn, params)
theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine='tc')
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
chunks = [bench.ae_chunk(c).to(dev) for c in range(nch)]
def fresh():
    prob = NativeProblem(net, theta, 'ggn', [(x, x) for x in chunks]); prob.linearize(); return prob
p1, p2 = fresh(), fresh()
v = torch.randn_like(theta)
g1, g2 = p1.gradient(), p2.gradient(); print('gradient identical across problems:', torch.equal(g1, g2))
f1, f2 = p1.fisher_diag(), p2.fisher_diag(); print('fisher identical:', torch.equal(f1, f2))
outs = [p1.mvp(v) for _ in range(4)] + [p2.mvp(v) for _ in range(2)]
print('mvp repeat identical:', [torch.equal(outs[0], o) for o in outs[1:]], 'max diff', max((outs[0]-o).abs().max().item() for o in outs[1:]))
M = DiagonalPreconditioner(f1, 1e-3)
sol = [pcg_device(p.matvec, -g1, minv=M.minv, damping=1e-3, max_iter=50, tol=0.0, martens_conv_crit=False, store_x_at_iters=None, poll=50)[0][-1].clone() for p in (p1, p1, p2)]
print('solve identical:', torch.equal(sol[0], sol[1]), torch.equal(sol[0], sol[2]), 'rel diff', ((sol[0]-sol[2]).norm()/sol[0].norm()).item())
