"""CPU, world_size 2 over gloo: the host-side sharding logic of the data-parallel path.  Each rank takes its
round-robin share of the chunks, forms N-weighted local sums with the GLOBAL count, all-reduces, and must land on
the oracle's full-batch gradient / GGN product / loss (reference tests/test_optimizer_acc.py checks the same
identity on one device)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hf_oracle as O
from helpers import ROOT, SPECS, build_loss, build_model, make_data

from pytorchhessianfree_b200.dist import all_reduce_sum, global_count, shard_chunks, world


def _worker(rank, world_size, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert world(dist.group.WORLD) == (rank, world_size)
        spec = SPECS["mlp_ce"]
        torch.manual_seed(0)
        model, loss_fn = build_model(spec), build_loss(spec, "sum")
        params = list(model.parameters())
        x, t = make_data(spec, 23, 5)
        sizes, chunks, off = [7, 1, 9, 6], [], 0
        for n in sizes:
            chunks.append((x[off:off + n], t[off:off + n]))
            off += n
        mine = shard_chunks(chunks, rank, world_size)
        n_total = global_count(sum(c[1].shape[0] for c in mine), dist.group.WORLD)
        v = torch.randn(sum(p.numel() for p in params), generator=torch.Generator().manual_seed(1))
        acc = torch.zeros(2 * v.numel() + 1)
        for cx, ct in mine:  # local sums of per-chunk SUM-reduced quantities, scaled by the global count ("mean")
            out = model(cx)
            loss = loss_fn(out, ct)
            g = O.flatten(torch.autograd.grad(loss, params, create_graph=True)).detach()
            acc += torch.cat([g, O.Gv(loss, out, params, v), loss.detach().reshape(1)]) / n_total
        all_reduce_sum(acc, dist.group.WORLD)
        q.put((rank, n_total, acc))
    finally:
        dist.destroy_process_group()


def test_shard_chunks_round_robin():
    chunks = list(range(7))
    assert shard_chunks(chunks, 0, 2) == [0, 2, 4, 6] and shard_chunks(chunks, 1, 2) == [1, 3, 5]
    assert sorted(sum((shard_chunks(chunks, r, 3) for r in range(3)), [])) == chunks
    assert shard_chunks(chunks, 0, 1) == chunks
    with pytest.raises(ValueError):
        shard_chunks(chunks, 2, 2)
    assert world() == (0, 1) and global_count(5) == 5


def test_two_rank_sharding_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = SPECS["mlp_ce"]
    torch.manual_seed(0)
    model, loss_fn = build_model(spec), build_loss(spec, "mean")
    params = list(model.parameters())
    x, t = make_data(spec, 23, 5)
    v = torch.randn(sum(p.numel() for p in params), generator=torch.Generator().manual_seed(1))
    out = model(x)
    loss = loss_fn(out, t)
    want = torch.cat([O.flatten(torch.autograd.grad(loss, params, create_graph=True)).detach(),
                      O.Gv(loss, out, params, v), loss.detach().reshape(1)])
    assert got[0][1] == got[1][1] == 23
    assert torch.equal(got[0][2], got[1][2]), "replicas must hold bit-identical vectors after the all-reduce"
    assert torch.allclose(got[0][2], want, rtol=1e-5, atol=1e-7)
