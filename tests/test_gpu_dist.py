"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the data-parallel path through the PUBLIC API.
Each rank runs ``HessianFree(..., process_group=WORLD).acc_step`` on its round-robin shard of the chunk list; the
replicas must stay bit-identical to each other and land on the parameters a single process reaches with the whole
chunk list (the sharding equivalence the reference tests on one device, tests/test_optimizer_acc.py:116-175)."""
import os
import sys
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, build_loss, build_model

pytestmark = pytest.mark.gpu
SPEC = dict(widths=[64, 96, 48, 10], act="relu", bias=[True] * 3, frozen=[], loss="ce")


def _data(step):
    g = torch.Generator().manual_seed(100 + step)
    x = torch.rand(6 * 64, 64, generator=g)
    t = torch.randint(0, 10, (6 * 64,), generator=g)
    return [(x[i * 64:(i + 1) * 64], t[i * 64:(i + 1) * 64]) for i in range(6)]


def _train(rank, world, group):
    from pytorchhessianfree_b200 import HessianFree
    from pytorchhessianfree_b200.dist import shard_chunks

    dev = torch.device("cuda", rank if world > 1 else 0)
    torch.manual_seed(0)
    model = build_model(SPEC).to(dev)
    loss_fn = build_loss(SPEC, "mean")
    opt = HessianFree(model.parameters(), process_group=group, cg_max_iter=20)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for step in range(2):
            chunks = [(x.to(dev), t.to(dev)) for x, t in shard_chunks(_data(step), rank, world)]
            opt.acc_step(model, loss_fn, chunks, reduction="mean")
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()
    return flat, opt.state["num_cg_iters"], opt.state["init_losses"]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        q.put((rank,) + _train(rank, world, dist.group.WORLD))
    except Exception as e:  # noqa: BLE001 -- report instead of leaving the parent to time out
        import traceback

        q.put((rank, "error", f"{e!r}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["nccl", "switch"])
def test_two_rank_acc_step_equals_single_process(exchange, monkeypatch):
    # the spawned ranks inherit the environment: HF_NVLS=1 routes the per-iteration exchange through
    # hf_allreduce_multimem (falls back to NCCL by itself where the fabric has no multicast)
    monkeypatch.setenv("HF_NVLS", "1" if exchange == "switch" else "0")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=150) for _ in procs], key=lambda r: r[0])
    for r in got:
        assert r[1] != "error", r[2]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _train(0, 1, None)
    assert torch.equal(got[0][1], got[1][1]), "replicas must stay bit-identical"
    assert got[0][2] == got[1][2] == single[1], "same CG iteration counts"
    assert got[0][3] == pytest.approx(single[2], rel=1e-5)
    assert torch.allclose(got[0][1], single[0], rtol=1e-3, atol=1e-5)


def _switch_worker(rank, world, port, q):
    """hf_allreduce_multimem against ncclAllReduce on the same partial vectors: whole vector, two disjoint slices (the
    overlapped form of the product exchange), a skipped launch, and back-to-back reuse of the self-resetting barrier."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from pytorchhessianfree_b200.dist import SymmetricVector

        n = 300_007  # not a multiple of 4 * world: the padded tail is reduced too and must stay zero
        sv = SymmetricVector.try_create(n, dev, dist.group.WORLD, force=True)
        if sv is None:
            q.put((rank, "unavailable"))
            return
        res = {}
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        for rep in range(3):
            part = torch.randn(n, device=dev, generator=g)
            want = part.clone()
            dist.all_reduce(want)
            sv.vec.copy_(part)
            sv.all_reduce_()
            res[f"whole{rep}"] = (sv.vec - want).abs().max().item() / want.abs().max().item()
            res[f"tail{rep}"] = sv.buf[n:].abs().sum().item()
            sv.vec.copy_(part)
            cut = sv.padded_range(0, 100_001)[1]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            sv.all_reduce_(cut, n, stream=side.cuda_stream)
            torch.cuda.current_stream().wait_stream(side)
            sv.all_reduce_(0, cut)
            res[f"sliced{rep}"] = (sv.vec - want).abs().max().item() / want.abs().max().item()
        flag = torch.ones(1, dtype=torch.int32, device=dev)
        sv.vec.copy_(part)
        sv.all_reduce_(skip_ptr=flag.data_ptr())
        res["skipped"] = torch.equal(sv.vec, part)
        sv.all_reduce_()
        gathered = [torch.empty_like(sv.vec) for _ in range(world)]
        dist.all_gather(gathered, sv.vec.contiguous())
        res["identical"] = all(torch.equal(gathered[0], t) for t in gathered)
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_switch_allreduce_matches_nccl():
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_switch_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if any(r[1] == "unavailable" for r in got):
        pytest.skip("no multicast mapping on this box: the exchange stays on NCCL")
    for _, res in got:
        for k, v in res.items():
            if k.startswith(("whole", "sliced")):
                assert v < 1e-6, (k, v)  # same addends, another summation order
            elif k.startswith("tail"):
                assert v == 0.0
            else:
                assert v is True, k
