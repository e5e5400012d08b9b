import sys, torch, os
sys.path[:0] = ['tests', 'oracle', '.']
from torch import nn
from test_gpu_conv import allcnnc, device_problem
N = int(sys.argv[1]); engine = sys.argv[2]
torch.manual_seed(0)
model = allcnnc(); loss_fn = nn.CrossEntropyLoss()
x, t = torch.rand(N, 3, 32, 32), torch.randint(0, 100, (N,))
prob = device_problem(model, loss_fn, [(x, t)], engine)
print("created", flush=True)
l = prob.linearize(); torch.cuda.synchronize(); print("linearize ok", float(l), flush=True)
g = prob.gradient(); torch.cuda.synchronize(); print("gradient ok", float(g.norm()), flush=True)
v = torch.randn_like(prob.theta)
o = prob.mvp(v); torch.cuda.synchronize(); print("mvp ok", float(o.norm()), flush=True)
