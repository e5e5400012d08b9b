import sys, torch
sys.path[:0] = ['.']
from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200._lib import Operand
lib = _lib.load(); dev = 'cuda'
st = torch.cuda.current_stream().cuda_stream
def bench(M, N, K, eng, reps=200, layout=(True, True)):
    a = torch.randn(M, K, device=dev) if layout[0] else torch.randn(K, M, device=dev)
    b = torch.randn(N, K, device=dev) if layout[1] else torch.randn(K, N, device=dev)
    c = torch.empty(M, N, device=dev)
    A = (Operand * 1)(Operand(a.data_ptr(), K, 1) if layout[0] else Operand(a.data_ptr(), 1, M))
    B = (Operand * 1)(Operand(b.data_ptr(), K, 1) if layout[1] else Operand(b.data_ptr(), 1, N))
    for _ in range(5):
        if lib.hf_contract(eng, M, N, K, 1, A, B, c.data_ptr(), N, None, 0, st) != 0:
            return float('nan')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.hf_contract(eng, M, N, K, 1, A, B, c.data_ptr(), N, None, 0, st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
x = torch.zeros(1024, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): x.add_(1)
e1.record(); torch.cuda.synchronize()
print("tiny torch kernel back-to-back: %.2f us" % (e0.elapsed_time(e1) * 1e3 / 200))
for (M, N) in [(128, 128), (4096, 512), (4096, 10)]:
    for K in [32, 64, 128, 256, 512, 1024, 2048]:
        t = bench(M, N, K, 1)
        print(f"tc   M={M} N={N} K={K}: {t:7.2f} us   ({2*M*N*K/t/1e6:8.2f} TFLOP/s)")
for K in [512, 4096]:
    for lay in [(False, False), (True, False)]:
        t = bench(512, 784, K, 1, layout=lay)
        print(f"tc   M=512 N=784 K={K} layout={lay}: {t:7.2f} us ({2*512*784*K/t/1e6:8.2f} TFLOP/s)")
for (M, N, K) in [(4096, 512, 784), (4096, 10, 512), (4096, 512, 10)]:
    print(f"simt M={M} N={N} K={K}: {bench(M, N, K, 0, reps=50):7.2f} us")
