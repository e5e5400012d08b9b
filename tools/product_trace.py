"""In-kernel %globaltimer timeline of the tensor-engine launches of ONE curvature product on configs[1]:
when each launch's first CTA starts and last CTA ends, i.e. kernel spans AND the gaps between dependent kernels."""
import sys, torch
sys.path[:0] = ['tests', '.']
from helpers import build_model, build_loss
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
from torch.nn.utils import parameters_to_vector
DEV = 'cuda'
MLP = dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")
torch.manual_seed(0)
model = build_model(MLP).to(DEV); loss_fn = build_loss(MLP, "mean")
x, t = torch.rand(4096, 784, device=DEV), torch.randint(0, 10, (4096,), device=DEV)
params = [p for p in model.parameters()]
prog = lower_module(model, loss_fn, params)
theta = parameters_to_vector(params).detach().clone()
net = NativeNet(prog.layers, prog.loss, prog.reduction, theta.numel(), engine="tc")
prob = NativeProblem(net, theta, "ggn", [(x, t)])
prob.linearize(); prob.gradient()
v = torch.randn_like(theta); out = torch.empty_like(theta)
for _ in range(5): prob.matvec(v, out)
lib = _lib.load()
EPOCHS = 32
tr = torch.zeros(EPOCHS * 1024 * 8, dtype=torch.int64, device=DEV)
torch.cuda.synchronize()
_lib.check(lib.hf_debug_tc_trace(tr.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(2_000_000); e0.record(); prob.matvec(v, out); prob.matvec(v, out); e1.record(); torch.cuda.synchronize()
_lib.check(lib.hf_debug_tc_trace(None))
print("two products by events: %.1f us" % (e0.elapsed_time(e1) * 1e3))
T = tr.view(EPOCHS, 1024, 8).cpu().double()
base = None
prev_end = None
for e in range(EPOCHS):
    used = T[e, :, 0] > 0
    if not used.any(): continue
    ent, end = T[e, used, 0], T[e, used, 4]
    if base is None: base = ent.min()
    line = "launch %2d  ctas %4d  first entry %8.2f  last entry %8.2f  first end %8.2f  last end %8.2f  span %6.2f" % (
        e, int(used.sum()), (ent.min() - base) / 1e3, (ent.max() - base) / 1e3, (end.min() - base) / 1e3, (end.max() - base) / 1e3,
        (end.max() - ent.min()) / 1e3)
    print(line)
