"""BASELINE.json configs[2] (Martens autoencoder, batch 60 000 as 8 chunks of 7 500) on one GPU: time of one GGN
product over all chunks and of one chunk (= the per-GPU work at 8 GPUs)."""
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
S = sum(a * b for a, b in zip(AE["widths"][:-1], AE["widths"][1:])); m1 = 784 * 1000
for engine in ("tc", "simt"):
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine=engine)
    for nchunks in (8, 1):
        x = torch.rand(7500 * nchunks, 784, device=DEV)
        prob = NativeProblem(net, theta, "ggn", [(x[i * 7500:(i + 1) * 7500],) * 2 for i in range(nchunks)])
        prob.linearize(); v = torch.randn_like(theta); out = torch.empty_like(theta)
        for _ in range(2): prob.matvec(v, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if engine == "tc" else 2
        e0.record()
        for _ in range(reps): prob.matvec(v, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * 7500 * nchunks * (4 * S - 2 * m1)
        print(f"engine={engine} batch={7500*nchunks}: {ms:.2f} ms per GGN product = {flops/ms/1e9:.1f} TFLOP/s algorithmic")
        del prob
