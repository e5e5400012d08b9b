"""Fused output head vs float64 autograd on the toy specs, single chunk and chunked, repeated calls."""
import sys, copy, torch
sys.path[:0] = ['tests', 'oracle', '.']
import hf_oracle as O
from helpers import SPECS, build_model, build_loss, make_data
from pytorchhessianfree_b200.lowering import lower_module
from pytorchhessianfree_b200.native import NativeNet
from pytorchhessianfree_b200.problem import NativeProblem
DEV = 'cuda'
for name in ("small_nn", "mwe", "mlp_ce", "ae_bce"):
    for red in ("mean", "sum"):
        spec = SPECS[name]
        torch.manual_seed(42)
        model = build_model(spec).to(DEV); loss_fn = build_loss(spec, red)
        x, t = make_data(spec, 15, seed=42)
        x, t = x.to(DEV), t.to(DEV)
        params = [p for p in model.parameters() if p.requires_grad]
        prog = lower_module(model, loss_fn, params)
        theta = torch.cat([p.detach().reshape(-1) for p in params])
        m64 = copy.deepcopy(model).double(); p64 = [p for p in m64.parameters() if p.requires_grad]
        out = m64(x.double()); loss = loss_fn(out, t.double() if t.is_floating_point() else t)
        for chunks in (None, [7, 8]):
            data = [(x, t)] if chunks is None else [(x[:7], t[:7]), (x[7:], t[7:])]
            prob = NativeProblem(NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine="simt"), theta, "ggn", data)
            prob.linearize(); prob.gradient()
            for rep in range(3):
                v = torch.randn_like(theta)
                want = O.Gv(loss, out, p64, v.double())
                got = prob.mvp(v)
                off, worst = 0, []
                for p in params:
                    n = p.numel(); a = got[off:off + n].double(); b = want[off:off + n]; off += n
                    worst.append(((a - b).abs().max() / (b.abs().max() + 1e-30)).item())
                print(f"{name:9s} {red:4s} chunks={chunks} rep {rep}: maxrel per slice " + " ".join(f"{w:.1e}" for w in worst))
