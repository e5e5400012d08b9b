"""GPU: the stand-alone contraction ``hf_contract`` (the tile kernels every curvature product is made of)
against a float64 torch matmul, for the three operand layouts, ragged shapes, two pairs and both engines."""
import pytest
import torch

from pytorchhessianfree_b200 import _lib
from pytorchhessianfree_b200._lib import Operand

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run_contract(engine, A_list, B_list, layouts):
    """A_list/B_list: logical [M,K] / [N,K] float32 tensors; layouts[i] = (a_kc, b_kc): True = K contiguous."""
    lib = _lib.load()
    M, K = A_list[0].shape
    N = B_list[0].shape[0]
    n = len(A_list)
    keep, A, B = [], (Operand * n)(), (Operand * n)()
    for i, (a, b, (a_kc, b_kc)) in enumerate(zip(A_list, B_list, layouts)):
        for arr, t, kc in ((A, a, a_kc), (B, b, b_kc)):
            if kc:
                s = t.contiguous()
                arr[i] = Operand(s.data_ptr(), s.shape[1], 1)
            else:
                s = t.t().contiguous()  # stored [K, MN]
                arr[i] = Operand(s.data_ptr(), 1, s.shape[1])
            keep.append(s)
    C = torch.full((M, N), float("nan"), device=DEV)
    ws_bytes = lib.hf_contract_workspace_bytes(M, N, K, n) if engine == 2 else 0
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=DEV) if ws_bytes else None
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256 if ws is not None else None
    rc = lib.hf_contract(engine, M, N, K, n, A, B, C.data_ptr(), N, ws_ptr, ws_bytes, torch.cuda.current_stream().cuda_stream)
    if rc == -2:
        pytest.skip("shape not supported by this engine: " + lib.hf_last_error_string().decode())
    _lib.check(rc)
    torch.cuda.synchronize()
    return C


# engine 0: FP32 FMA tiles.  engines 1, 2: split-precision tensor-core tiles (TF32 hi = truncated word, BF16 corrections
# from the exact remainder, lo*lo dropped): ~2^-19 relative, two orders inside the rtol 1e-4 the curvature products have to
# meet.  Engine 1 derives the BF16 forms in its main loop (128x128 tiles); engine 2 reads pre-split operand images
# (256x256 CTA-pair tiles).
TOL = {0: 2e-6, 1: 2e-5, 2: 2e-5}
SHAPES = [(16, 10, 10), (128, 128, 64), (512, 784, 512), (4096, 512, 784), (512, 784, 4096), (257, 67, 130),
          (10, 512, 512), (512, 10, 512), (64, 64, 16), (300, 1000, 500), (1024, 1024, 8), (96, 200, 1000)]
LAYOUTS = [(True, True), (True, False), (False, False), (False, True)]
# shapes that exercise the pair engine's own corners: ragged 256-tiles, odd number of 128-row blocks, K not a multiple
# of 32 or 8, a single k-block, more k-blocks than ring stages, the autoencoder's widths
SHAPES2 = [(7500, 1000, 784), (1000, 784, 7500), (7500, 250, 30), (300, 520, 100), (256, 256, 32), (129, 257, 33),
           (384, 512, 2048)]


@pytest.mark.parametrize("engine", [0, 1, 2])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("layout", LAYOUTS)
def test_single_pair(engine, shape, layout):
    M, N, K = shape
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device=DEV, generator=g)
    b = torch.randn(N, K, device=DEV, generator=g)
    got = run_contract(engine, [a], [b], [layout])
    want = a.double() @ b.double().t()
    err = (got.double() - want).abs().max().item() / want.abs().max().item()
    assert err < TOL[engine] * max(1.0, (K / 512) ** 0.5), f"rel err {err:.2e}"  # FP32 accumulation grows ~sqrt(K)


# the persistent pair kernel takes launches of >= 4 waves of 256x256 tiles by itself (the last two shapes: 314 and 320
# tiles on 74 pairs, ragged edges, 17 and 19 k-blocks); HF_TC2_PERSIST=2 sends every shape through it,
# HF_TC2P_BK=16 through its six-stage ring
SHAPES2P = SHAPES2 + [(40000, 512, 520), (20300, 1000, 600)]


@pytest.mark.parametrize("persist", ["1", "2", "2/16", "0"])
@pytest.mark.parametrize("shape", SHAPES2P)
@pytest.mark.parametrize("layout", LAYOUTS)
def test_pair_engine_shapes(shape, layout, persist, monkeypatch):
    monkeypatch.setenv("HF_TC2_PERSIST", persist.split("/")[0])
    monkeypatch.setenv("HF_TC2P_BK", persist.split("/")[1] if "/" in persist else "32")
    M, N, K = shape
    g = torch.Generator(device=DEV).manual_seed(M * 5 + N * 11 + K)
    a = torch.randn(M, K, device=DEV, generator=g)
    b = torch.randn(N, K, device=DEV, generator=g)
    got = run_contract(2, [a], [b], [layout])
    want = a.double() @ b.double().t()
    err = (got.double() - want).abs().max().item() / want.abs().max().item()
    assert err < TOL[2] * max(1.0, (K / 512) ** 0.5), f"rel err {err:.2e}"


@pytest.mark.parametrize("engine", [0, 1, 2])
@pytest.mark.parametrize("shape", [(512, 512, 784), (4096, 784, 512), (100, 36, 52), (512, 784, 1024)])
@pytest.mark.parametrize("layout", LAYOUTS[:3])
def test_two_pairs_accumulate(engine, shape, layout):
    M, N, K = shape
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = [torch.randn(M, K, device=DEV, generator=g) for _ in range(2)]
    b = [torch.randn(N, K, device=DEV, generator=g) for _ in range(2)]
    got = run_contract(engine, a, b, [layout, layout])
    want = a[0].double() @ b[0].double().t() + a[1].double() @ b[1].double().t()
    err = (got.double() - want).abs().max().item() / want.abs().max().item()
    assert err < TOL[engine] * max(1.0, (K / 512) ** 0.5), f"rel err {err:.2e}"
