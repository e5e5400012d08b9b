"""In-kernel %globaltimer phases of one CTA-pair contraction (gemm_tc2.cu): where a launch spends its time outside the
main loop.  Prints, over all CTAs, the median / max time of each phase boundary relative to the earliest CTA entry."""
import os, sys, torch
sys.path[:0] = ['.']
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200._lib import Operand
lib = _lib.load(); dev = 'cuda'; st = torch.cuda.current_stream().cuda_stream
NAMES = ["entry", "prologue done", "first stage landed", "accumulator complete", "first half staged", "first half stored", "all stored", "pair released"]
def trace(M, N, K):
    a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev); c = torch.empty(M, N, device=dev)
    A = (Operand * 1)(Operand(a.data_ptr(), K, 1)); B = (Operand * 1)(Operand(b.data_ptr(), K, 1))
    nb = lib.hf_contract_workspace_bytes(M, N, K, 1); ws = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 255) // 256 * 256
    assert lib.hf_contract(2, M, N, K, 1, A, B, c.data_ptr(), N, wp, nb, st) == 0
    for _ in range(3): lib.hf_contract(3, M, N, K, 1, A, B, c.data_ptr(), N, wp, nb, st)
    buf = torch.zeros(4096 * 8, dtype=torch.int64, device=dev)
    lib.hf_debug_tc2_trace(buf.data_ptr()); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.hf_contract(3, M, N, K, 1, A, B, c.data_ptr(), N, wp, nb, st); e1.record(); torch.cuda.synchronize()
    lib.hf_debug_tc2_trace(None)
    t = buf.view(4096, 8).cpu()
    ctas = 2 * (-(-M // 256)) * (-(-N // 256))
    t = t[:ctas].double()
    t0 = t[:, 0].min()
    print(f"M={M} N={N} K={K}: {ctas} CTAs, HF_TC2_PERSIST={os.environ.get('HF_TC2_PERSIST', '1')}, event time {e0.elapsed_time(e1)*1e3:.1f} us")
    for i, nm in enumerate(NAMES):
        col = t[:, i]; col = col[col > 0]
        if len(col) == 0: continue
        r = (col - t0) / 1e3
        print(f"   {nm:22s} median {r.median():7.2f} us   min {r.min():7.2f}   max {r.max():7.2f}   ({len(col)} CTAs)")
for shape in [(7500, 1000, 784), (7500, 1000, 32), (4096, 512, 784)]:
    trace(*shape)
