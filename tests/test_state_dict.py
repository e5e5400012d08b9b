"""CPU: checkpoint/resume surface.  ``damping`` lives in ``param_groups[0]``, the CG
warm start ``x0`` and the log lists live in ``state`` under string keys (reference optimizer.py:104-110, :183-192);
``state_dict()`` / ``load_state_dict()`` must carry them over unchanged, as the reference's do."""
import torch
from torch import nn

from pytorchhessianfree_b200 import HessianFree


def test_state_dict_round_trip():
    model = nn.Sequential(nn.Linear(4, 3), nn.ReLU(), nn.Linear(3, 2))
    opt = HessianFree(model.parameters(), damping=0.5, cg_max_iter=17, lr=0.7)
    st = opt._log_state()
    st["x0"] = torch.arange(23, dtype=torch.float32)
    st["dampings"].extend([0.5, 0.75])
    st["cg_reasons"].append("Convergence (Martens)")
    opt._group["damping"] = 0.75
    blob = opt.state_dict()
    assert blob["param_groups"][0]["damping"] == 0.75 and blob["param_groups"][0]["cg_max_iter"] == 17

    twin = nn.Sequential(nn.Linear(4, 3), nn.ReLU(), nn.Linear(3, 2))
    opt2 = HessianFree(twin.parameters())
    opt2.load_state_dict(blob)
    g = opt2.param_groups[0]
    assert (g["damping"], g["cg_max_iter"], g["lr"], g["curvature_opt"]) == (0.75, 17, 0.7, "ggn")
    assert torch.equal(opt2.state["x0"], st["x0"])
    assert opt2.state["dampings"] == [0.5, 0.75] and opt2.state["cg_reasons"] == ["Convergence (Martens)"]
    # the group the optimizer reads in step() must be the loaded one
    assert opt2._group is opt2.param_groups[0] and opt2._group["damping"] == 0.75
