"""Shared helpers for the test-suite: tiny-net builders and path set-up.

The nets mirror the fixtures of the reference's own tests
(reference tests/test_utils.py:19-52: Linear(7,5)-ReLU-[Linear(5,5)-ReLU]-Linear(5,3),
first layer frozen, MSE) and the BASELINE.json configs at toy widths.
"""
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

ACTS = {"relu": nn.ReLU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh}
LOSSES = {"mse": nn.MSELoss, "ce": nn.CrossEntropyLoss, "bce": nn.BCEWithLogitsLoss}

# name -> spec.  widths[0] is the input width; one activation between consecutive Linear layers.
SPECS = {
    # reference tests/test_utils.py:19-52 (nested Sequential, frozen first layer)
    "small_nn": dict(widths=[7, 5, 5, 3], act="relu", bias=[True, True, True], frozen=[0], loss="mse", nested=True),
    # reference examples/run_mwe.py:16-21 (BASELINE.json configs[0])
    "mwe": dict(widths=[10, 10, 10], act="relu", bias=[False, True], frozen=[], loss="mse"),
    # BASELINE.json configs[1] at toy width
    "mlp_ce": dict(widths=[12, 16, 16, 5], act="relu", bias=[True, True, True], frozen=[], loss="ce"),
    # BASELINE.json configs[2] at toy width (sigmoid hidden units, linear code layer, logits out)
    "ae_bce": dict(widths=[12, 8, 4, 8, 12], act="sigmoid", bias=[True] * 4, frozen=[], loss="bce", linear_after=[1]),
    "tanh_mse": dict(widths=[6, 9, 4], act="tanh", bias=[True, True], frozen=[], loss="mse"),
}


def build_model(spec, dtype=torch.float32):
    """Sequential of Linear / activation layers described by ``spec``."""
    w = spec["widths"]
    mods = []
    n_lin = len(w) - 1
    for i in range(n_lin):
        lin = nn.Linear(w[i], w[i + 1], bias=spec["bias"][i])
        if i in spec.get("frozen", []):
            for p in lin.parameters():
                p.requires_grad = False
        mods.append(lin)
        if i < n_lin - 1 and i not in spec.get("linear_after", []):
            mods.append(ACTS[spec["act"]]())
    if spec.get("nested"):
        # same shape as the reference fixture: the middle Linear+act wrapped in a Sequential
        mods = [mods[0], mods[1], nn.Sequential(mods[2], mods[3])] + mods[4:]
    return nn.Sequential(*mods).to(dtype)


def build_loss(spec, reduction="mean"):
    return LOSSES[spec["loss"]](reduction=reduction)


def make_data(spec, n, seed, dtype=torch.float32):
    """Seeded inputs/targets: U[0,1) inputs; U[0,1) targets (mse, bce) or class ids (ce)."""
    g = torch.Generator().manual_seed(seed)
    w = spec["widths"]
    x = torch.rand(n, w[0], generator=g, dtype=dtype)
    if spec["loss"] == "ce":
        t = torch.randint(0, w[-1], (n,), generator=g)
    else:
        t = torch.rand(n, w[-1], generator=g, dtype=dtype)
    return x, t


def spd_system(dim, seed=0, dtype=torch.float32):
    """A = R R^T + 1e-3 I, b = A x (same construction as reference tests/test_utils.py:6-16)."""
    torch.manual_seed(seed)
    R = torch.rand((dim, dim)) - 0.5
    A = R @ R.T + 1e-3 * torch.eye(dim)
    x = torch.rand(dim) - 0.5
    return A.to(dtype), (A @ x).to(dtype), x.to(dtype)


# ---- BASELINE.json configs at benchmark size (tests/golden/benchsize.pt, tests/test_gpu_benchsize.py) ----
BENCH_CFGS = {
    # configs[1]: MLP 784-512-512-10 ReLU, CrossEntropy, batch 4096
    "cfg2": dict(spec=dict(widths=[784, 512, 512, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce"), n=4096),
    # configs[2]: Martens autoencoder, one 7 500-sample shard (the per-GPU share of the 60 000 batch at 8 GPUs)
    "cfg3": dict(spec=dict(widths=[784, 1000, 500, 250, 30, 250, 500, 1000, 784], act="sigmoid", bias=[True] * 8, frozen=[],
                           loss="bce", linear_after=[3]), n=7500),
}


def benchsize_problem(cfg, seed):
    """(model, loss_fn, x, t, v) of a benchmark-size parity case, re-created from the seed alone."""
    c = BENCH_CFGS[cfg]
    torch.manual_seed(seed)
    model = build_model(c["spec"])
    loss_fn = build_loss(c["spec"], "mean")
    x, t = make_data(c["spec"], c["n"], 1000 + seed)
    if c["spec"]["loss"] == "bce":
        t = x.clone()  # the autoencoder reconstructs its inputs
    g = torch.Generator().manual_seed(2000 + seed)
    v = torch.randn(sum(p.numel() for p in model.parameters() if p.requires_grad), generator=g)
    return model, loss_fn, x, t, v


def conv_fixture_case(c):
    """Rebuild model, loss and data of one record of tests/golden/conv.pt (minted from the unmodified reference by
    tests/golden/make_golden.py conv; the architectures live in tests/test_gpu_conv.py)."""
    from test_gpu_conv import CASES

    build, loss_cls, _, _, _ = CASES[c["net"]]
    model = build()
    model.load_state_dict(c["state"])
    return model, loss_cls(), c["x"], c["t"], c["v"]
