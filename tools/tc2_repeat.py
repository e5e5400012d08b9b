"""Bitwise repeatability of the pair engine: the same contraction launched many times must return the same bits."""
import sys, torch
sys.path[:0] = ['.']
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200._lib import Operand
lib = _lib.load(); dev = 'cuda'; st = torch.cuda.current_stream().cuda_stream
def run(M, N, K, layout, pairs=1, reps=30):
    a = [torch.randn(M, K, device=dev) if layout[0] else torch.randn(K, M, device=dev) for _ in range(pairs)]
    b = [torch.randn(N, K, device=dev) if layout[1] else torch.randn(K, N, device=dev) for _ in range(pairs)]
    A = (Operand * pairs)(*[Operand(t.data_ptr(), K, 1) if layout[0] else Operand(t.data_ptr(), 1, M) for t in a])
    B = (Operand * pairs)(*[Operand(t.data_ptr(), K, 1) if layout[1] else Operand(t.data_ptr(), 1, N) for t in b])
    nb = lib.hf_contract_workspace_bytes(M, N, K, pairs); ws = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 255) // 256 * 256
    outs = []
    for i in range(reps):
        c = torch.full((M, N), float('nan'), device=dev)
        assert lib.hf_contract(2 if i == 0 else 3, M, N, K, pairs, A, B, c.data_ptr(), N, wp, nb, st) == 0
        outs.append(c)
    torch.cuda.synchronize()
    bad = sum(not torch.equal(outs[0], o) for o in outs[1:])
    md = max((outs[0] - o).abs().max().item() for o in outs[1:])
    print(f"M={M} N={N} K={K} layout={layout} pairs={pairs}: {bad}/{reps-1} differ, max diff {md:.2e} (|C| max {outs[0].abs().max().item():.1f})")
for lay in [(True, True), (True, False), (False, False)]:
    run(7500, 1000, 784, lay)
run(7500, 500, 1000, (True, True), pairs=2)
run(1000, 784, 7500, (False, False))
run(7500, 250, 500, (True, True), pairs=2)
