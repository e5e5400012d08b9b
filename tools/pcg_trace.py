import sys, torch
sys.path[:0] = ['.']
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200.cg import _Solver
from pytorchhessianfree_b200._lib import PCG_FUSED
lib = _lib.load(); dev = 'cuda'
for P in (669706, 2837314):
    b = torch.randn(P, device=dev); s = _Solver(b, 10**6)
    minv = torch.rand(P, device=dev) + 0.5; Bp = torch.randn(P, device=dev)
    s.init(None, None, minv, 1e-3, 0.0, None, False, False)
    tr = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
    for _ in range(5): s.iterate(PCG_FUSED, Bp=Bp, minv=minv, lam=1e-3)
    torch.cuda.synchronize()
    # back-to-back timing without trace
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000); e0.record()
    for _ in range(50): s.iterate(PCG_FUSED, Bp=Bp, minv=minv, lam=1e-3)
    e1.record(); torch.cuda.synchronize()
    print(f"P={P}: back-to-back {e0.elapsed_time(e1)*1e3/50:.2f} us per launch")
    _lib.check(lib.hf_debug_pcg_trace(tr.data_ptr()))
    if len(sys.argv) > 1 and sys.argv[1] == "cold":
        torch.empty(256 << 20, dtype=torch.uint8, device=dev).fill_(1); torch.cuda.synchronize()
    s.iterate(PCG_FUSED, Bp=Bp, minv=minv, lam=1e-3); torch.cuda.synchronize()
    _lib.check(lib.hf_debug_pcg_trace(None))
    n = lib.hf_device_sm_count()
    t = tr.view(256, 8)[:n, :6].cpu().double()
    t0 = t[:, 0].min()
    names = ["start", "p.Ap partial", "alpha known", "x,r updated", "beta known", "p written"]
    for k in range(6):
        col = t[:, k] - t0
        print(f"   {names[k]:14s} min {col.min().item()/1e3:7.2f}  median {col.median().item()/1e3:7.2f}  max {col.max().item()/1e3:7.2f} us")
