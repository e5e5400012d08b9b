import sys, torch
sys.path[:0] = ['.']
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200._lib import Operand
lib = _lib.load(); dev = 'cuda'; st = torch.cuda.current_stream().cuda_stream
for (M, N, K) in [(4096, 512, 32), (128, 128, 64), (4096, 512, 784)]:
    a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev); c = torch.empty(M, N, device=dev)
    A = (Operand * 1)(Operand(a.data_ptr(), K, 1)); B = (Operand * 1)(Operand(b.data_ptr(), K, 1))
    run = lambda: lib.hf_contract(1, M, N, K, 1, A, B, c.data_ptr(), N, None, 0, st)
    for _ in range(5): run()
    n_ctas = max(1, (M // 128) * (N // 128))
    tr = torch.zeros(n_ctas * 8, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    _lib.check(lib.hf_debug_tc_trace(tr.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(1_000_000); e0.record(); run(); e1.record(); torch.cuda.synchronize()
    _lib.check(lib.hf_debug_tc_trace(None))
    t = tr.view(n_ctas, 8)[:, :8].cpu().double(); t0 = t[:, 0].min()
    print(f"M={M} N={N} K={K}: event time {e0.elapsed_time(e1)*1e3:.2f} us")
    for k, nm in enumerate(["entry", "prologue done", "first stage landed", "accumulator done", "epilogue done", "tmem->smem done", "first 4 rows stored", "all rows stored"]):
        col = t[:, k] - t0
        print(f"   {nm:20s} min {col.min().item()/1e3:7.2f}  median {col.median().item()/1e3:7.2f}  max {col.max().item()/1e3:7.2f} us")
