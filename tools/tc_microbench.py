"""Back-to-back launch time of the tensor-tile contractions over shapes, K and operand layouts.
engine 1 = 128x128 tiles with the in-kernel splitter, engine 2 = 256x256 CTA-pair tiles on pre-split images
(timed through engine 3: images already built)."""
import sys, torch
sys.path[:0] = ['.']
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200._lib import Operand
lib = _lib.load(); dev = 'cuda'
st = torch.cuda.current_stream().cuda_stream
def bench(M, N, K, eng, reps=100, layout=(True, True), pairs=1):
    a = [torch.randn(M, K, device=dev) if layout[0] else torch.randn(K, M, device=dev) for _ in range(pairs)]
    b = [torch.randn(N, K, device=dev) if layout[1] else torch.randn(K, N, device=dev) for _ in range(pairs)]
    c = torch.empty(M, N, device=dev)
    A = (Operand * pairs)(*[Operand(t.data_ptr(), K, 1) if layout[0] else Operand(t.data_ptr(), 1, M) for t in a])
    B = (Operand * pairs)(*[Operand(t.data_ptr(), K, 1) if layout[1] else Operand(t.data_ptr(), 1, N) for t in b])
    nb = lib.hf_contract_workspace_bytes(M, N, K, pairs) if eng >= 2 else 0
    ws = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 255) // 256 * 256
    if lib.hf_contract(min(eng, 2), M, N, K, pairs, A, B, c.data_ptr(), N, wp, nb, st) != 0:
        return float('nan')
    e = 3 if eng == 2 else eng
    for _ in range(5):
        lib.hf_contract(e, M, N, K, pairs, A, B, c.data_ptr(), N, wp, nb, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.hf_contract(e, M, N, K, pairs, A, B, c.data_ptr(), N, wp, nb, st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
KK = (True, True); MM = (False, False); KM = (True, False)
cases = [  # (M, N, K, layout, pairs, what)
    (4096, 512, 784, KK, 1, "cfg2 R-op 1"), (4096, 512, 512, KK, 2, "cfg2 R-op 2"), (4096, 512, 512, KM, 1, "cfg2 transposed"),
    (512, 784, 4096, MM, 1, "cfg2 weight grad 1"),
    (7500, 1000, 784, KK, 1, "cfg3 R-op 1"), (7500, 500, 1000, KK, 2, "cfg3 R-op 2"), (7500, 784, 1000, KK, 2, "cfg3 R-op 8"),
    (7500, 1000, 784, KM, 1, "cfg3 transposed 8"), (1000, 784, 7500, MM, 1, "cfg3 weight grad 1 (unsplit)"),
    (60000, 1000, 784, KK, 1, "cfg3 R-op 1, one 60000 chunk"),
    (7500, 250, 500, KK, 2, "cfg3 R-op 3 (narrow)"), (7500, 30, 250, KK, 2, "cfg3 R-op 4 (code layer)"), (7500, 250, 30, KK, 2, "cfg3 R-op 5"),
    (7500, 500, 250, KK, 2, "cfg3 R-op 6"), (7500, 500, 250, KM, 1, "cfg3 transposed 3"), (7500, 250, 500, KM, 1, "cfg3 transposed 6"),
    (250, 500, 7500, MM, 1, "cfg3 weight grad 3 (unsplit)"),
    (8192, 8192, 2048, KK, 1, "large square"), (16384, 4096, 4096, KK, 1, "large"),
]
for M, N, K, lay, pairs, what in cases:
    f = 2.0 * M * N * K * pairs
    row = f"{what:32s} M={M:6d} N={N:5d} K={K:5d} x{pairs}:"
    for eng in (1, 2):
        t = bench(M, N, K, eng, layout=lay, pairs=pairs)
        row += f"   engine {eng}: {t:8.2f} us {f / t / 1e6:7.1f} TF/s"
    print(row, flush=True)
for K in [32, 256, 1024, 4096]:
    t1, t2 = bench(9472, 2048, K, 1), bench(9472, 2048, K, 2)   # 37 x 8 pair tiles = 4 full waves of 74 pairs
    print(f"slope M=9472 N=2048 K={K}: engine 1 {t1:8.2f} us, engine 2 {t2:8.2f} us")
