"""Per-k-block pipeline timeline of one CTA of the tensor-tile kernel (M=4096 N=512 K=2048): when the producer got the
slot, issued the TMA, when the splitters saw the bytes, finished, and when the MMAs of the block were issued.
Needs a trace build: HF_B200_DEFS=-DHF_TC_ITER_TRACE=1 python -m pytorchhessianfree_b200.build --force (the hooks cost
~15 % of the main loop, so they are compiled out by default)."""
import sys, torch
sys.path[:0] = ['.']
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200._lib import Operand
lib = _lib.load(); dev = 'cuda'; st = torch.cuda.current_stream().cuda_stream
M, N, K = 4096, 512, 2048
a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev); c = torch.empty(M, N, device=dev)
A = (Operand * 1)(Operand(a.data_ptr(), K, 1)); B = (Operand * 1)(Operand(b.data_ptr(), K, 1))
run = lambda: lib.hf_contract(1, M, N, K, 1, A, B, c.data_ptr(), N, None, 0, st)
for _ in range(5): run()
tr = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
_lib.check(lib.hf_debug_tc_trace_iters(tr.data_ptr()))
run(); torch.cuda.synchronize()
_lib.check(lib.hf_debug_tc_trace_iters(None))
t = tr.view(64, 8).cpu().double(); t0 = t[0, 4]
names = ["slot free", "tma issued", "raw seen", "split done", "mma issued"]
print("  it " + " ".join(f"{n:>11s}" for n in names) + "   (us since the first slot wait; deltas: tma->raw, raw->split, split->mma)")
for it in range(40):
    row = [t[it, 4], t[it, 0], t[it, 1], t[it, 2], t[it, 3]]
    v = [(x - t0).item() / 1e3 for x in row]
    print(f"{it:4d} " + " ".join(f"{x:11.2f}" for x in v) + f"    {v[2]-v[1]:6.2f} {v[3]-v[2]:6.2f} {v[4]-v[3]:6.2f}")
