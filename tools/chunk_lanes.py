"""Chunk products in flight (NativeProblem._local_products, HF_CHUNK_LANES) against one after the other, bench.py's
workload on one GPU: the 50-iteration solve over n resident 7 500-sample chunks, n = 8 (the N = 1 bench), 4, 2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

from pytorchhessianfree_b200 import DiagonalPreconditioner, pcg_device  # noqa: E402
from pytorchhessianfree_b200.lowering import lower_module  # noqa: E402
from pytorchhessianfree_b200.native import NativeNet  # noqa: E402
from pytorchhessianfree_b200.problem import NativeProblem  # noqa: E402

dev = torch.device("cuda", 0)
model = B.build_ae(0).to(dev)
loss_fn = torch.nn.BCEWithLogitsLoss()
params = list(model.parameters())
prog = lower_module(model, loss_fn, params)
theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine="tc")


def solve_ms(chunks, lanes):
    os.environ["HF_CHUNK_LANES"] = str(lanes)
    prob = NativeProblem(net, theta, "ggn", [(x, x) for x in chunks])
    prob.linearize()
    g = prob.gradient()
    M = DiagonalPreconditioner(prob.fisher_diag(), B.DAMPING)

    def solve():
        return pcg_device(prob.matvec, -g, minv=M.minv, damping=B.DAMPING, max_iter=B.K_CG, tol=0.0, martens_conv_crit=False,
                          store_x_at_iters=None, poll=B.K_CG)
    for _ in range(2):
        xs = solve()[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        solve()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3, xs[-1]


for n in (8, 4, 2):
    chunks = [B.ae_chunk(c).to(dev) for c in range(n)]
    base, x0 = solve_ms(chunks, 1)
    line = f"{n} chunks: one after the other {base:.1f} ms ({1e3 * base / B.K_CG / n:.0f} us per chunk product)"
    for lanes in (2, 3, 4):
        if lanes > n:
            continue
        ms, x1 = solve_ms(chunks, lanes)
        line += f"; {lanes} lanes {ms:.1f} ms ({1e3 * ms / B.K_CG / n:.0f} us, iterates differ by {((x1 - x0).norm() / x0.norm()).item():.0e})"
    print(line, flush=True)
    del chunks
