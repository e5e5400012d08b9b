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
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_acc_step_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _train(0, 1, None)
    assert torch.equal(got[0][1], got[1][1]), "replicas must stay bit-identical"
    assert got[0][2] == got[1][2] == single[1], "same CG iteration counts"
    assert got[0][3] == pytest.approx(single[2], rel=1e-5)
    assert torch.allclose(got[0][1], single[0], rtol=1e-3, atol=1e-5)
