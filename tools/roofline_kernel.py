"""The launch bench.py's `roofline` block times -- the 7500 x 1000 x 784 contraction of the autoencoder's first R-op on
the pair engine, operand images built once, L2 flushed before every launch -- on its own, for ncu:

  ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/prof_roofline \
      python tools/roofline_kernel.py

(`dram__bytes_read.sum + dram__bytes_write.sum` of that launch is the `traffic` of the bench line.)"""
import sys

import torch

sys.path[:0] = ["."]
from pytorchhessianfree_b200 import _lib  # noqa: E402
from pytorchhessianfree_b200._lib import Operand  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (7500, 1000, 784)
lib, dev = _lib.load(), "cuda"
a, b, c = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.empty(M, N, device=dev)
A, B = (Operand * 1)(Operand(a.data_ptr(), K, 1)), (Operand * 1)(Operand(b.data_ptr(), K, 1))
nb = lib.hf_contract_workspace_bytes(M, N, K, 1)
ws = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
wp = (ws.data_ptr() + 255) // 256 * 256
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_lib.check(lib.hf_contract(2, M, N, K, 1, A, B, c.data_ptr(), N, wp, nb, st))  # builds the images, then one launch
for _ in range(4):
    flush.fill_(1)
    _lib.check(lib.hf_contract(3, M, N, K, 1, A, B, c.data_ptr(), N, wp, nb, st))
torch.cuda.synchronize()
print("done")
