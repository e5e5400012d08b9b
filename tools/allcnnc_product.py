"""BASELINE.json configs[4]'s architecture (All-CNN-C, CIFAR-100 shape, eval mode): time of one GGN-vector product and of
the linearisation at the per-GPU batch of the 8-GPU experiment (1024) and at the batch BASELINE.md quotes the CPU
reference for (64: 537 ms per `_Gv` on 8 host cores)."""
import sys, torch
sys.path[:0] = ['tests', 'oracle', '.']
from torch import nn
from test_gpu_conv import allcnnc, device_problem
torch.manual_seed(0)
model = allcnnc(); loss_fn = nn.CrossEntropyLoss()
# MACs per sample of the nine convolutions on their output maps (S = 271 420 416, m1 = 2 654 208)
S, m1 = 271420416, 2654208
for N in (64, 256, 1024):
    x, t = torch.rand(N, 3, 32, 32), torch.randint(0, 100, (N,))
    for engine in ("tc", "simt"):
        if engine == "simt" and N > 256: continue
        prob = device_problem(model, loss_fn, [(x, t)], engine)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.linearize(); e1.record(); torch.cuda.synchronize(); t_lin = e0.elapsed_time(e1)
        v = torch.randn_like(prob.theta); out = torch.empty_like(prob.theta)
        for _ in range(2): prob.matvec(v, out)
        torch.cuda.synchronize()
        reps = 5
        e0.record()
        for _ in range(reps): prob.matvec(v, out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * N * (4 * S - 2 * m1)
        ws = sum(l.workspace.numel() for l in prob.mvp_lins) / 2**30
        prob.fisher_diag(); torch.cuda.synchronize()
        e0.record(); prob.fisher_diag(); e1.record(); torch.cuda.synchronize()
        print(f"All-CNN-C N={N} engine={engine}: GGN product {ms:.2f} ms = {flops/ms/1e9:.1f} TFLOP/s algorithmic; linearise {t_lin:.1f} ms; "
              f"Fisher diagonal {e0.elapsed_time(e1):.1f} ms; workspace {ws:.1f} GiB", flush=True)
        del prob
