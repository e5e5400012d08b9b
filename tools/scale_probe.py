"""Where does an iteration of the sharded solve go?  Launch under torchrun (one rank per GPU), bench.py's workload:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 tools/scale_probe.py

Times, in ONE process group (so that the variants see the same boards):
  * the exchange alone: ncclAllReduce vs hf_allreduce_multimem, by barrier flavour, CTA count and message size
    (one quantum = latency floor: launch + two rendezvous; the whole vector);
  * the 50-iteration solve with the exchange on NCCL / through the switch, overlapped with the first layer's weight
    gradient or not;
  * the solve with NO exchange at all (each rank on its shard: the compute floor of an iteration at this N).
Rank 0 prints one line per measurement."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

from pytorchhessianfree_b200 import DiagonalPreconditioner, _lib, pcg_device  # noqa: E402
from pytorchhessianfree_b200.dist import SymmetricVector, all_reduce_sum  # noqa: E402
from pytorchhessianfree_b200.lowering import lower_module  # noqa: E402
from pytorchhessianfree_b200.native import NativeNet  # noqa: E402
from pytorchhessianfree_b200.problem import NativeProblem  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
    lib = _lib.load()

    def say(*a):
        if rank == 0:
            print(*a, flush=True)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    model = B.build_ae(0).to(dev)
    loss_fn = torch.nn.BCEWithLogitsLoss()
    params = list(model.parameters())
    prog = lower_module(model, loss_fn, params)
    theta = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
    P = theta.numel()

    # ---- the exchange alone ----
    buf = torch.zeros(P, device=dev)
    say(f"world {world}  P {P}")
    say(f"nccl whole vector: {1e3 * timed(lambda: all_reduce_sum(buf, group), 50):.1f} us")
    sv = SymmetricVector.try_create(P, dev, group, force=True)
    cap = 0
    if sv is None:
        say("no multicast mapping: nothing else to compare")
    else:
        say(f"signal pad {sv.hdl.signal_pad_size} B -> at most {sv.max_blocks} CTAs")
        cap = max(1, min(64, sv.hdl.signal_pad_size // (4 * world)))
        for variant in (0, 1, 2, 3):
            lib.hf_debug_allreduce_variant(variant)
            floor = 1e3 * timed(lambda: sv.all_reduce_(0, sv.quantum), 50)
            row = [f"variant {variant}: one quantum {floor:.1f} us; whole vector by CTAs:"]
            for blocks in (8, 16, 32, 64):
                if blocks > cap:
                    continue
                sv.max_blocks = blocks
                row.append(f"{blocks}: {1e3 * timed(lambda: sv.all_reduce_(), 50):.1f} us")
            sv.max_blocks = cap
            say("  ".join(row))
        lib.hf_debug_allreduce_variant(3)

    # ---- the solve ----
    net = NativeNet(prog.layers, prog.loss, prog.reduction, prog.n_params, engine="tc")
    mine = [B.ae_chunk(c).to(dev) for c in range(B.CHUNKS) if c % world == rank]

    def problem(grp):
        prob = NativeProblem(net, theta, "ggn", [(x, x) for x in mine], group=grp)
        prob.linearize()
        g = prob.gradient()
        return prob, g, DiagonalPreconditioner(prob.fisher_diag(), B.DAMPING)

    def solve(prob, g, M):
        return pcg_device(prob.matvec, -g, minv=M.minv, damping=B.DAMPING, max_iter=B.K_CG, tol=0.0, martens_conv_crit=False,
                          store_x_at_iters=None, poll=B.K_CG, out_buffer=prob.out_buffer())

    prob, g, M = problem(group)
    symm = sv  # same (group, length, device): the cached vector
    cases = [("nccl, overlapped", None, True, 3, 64), ("nccl, in line", None, False, 3, 64)]
    if symm is not None:
        for variant in (3, 0):
            for blocks in (64, 16):
                cases.append((f"switch v{variant} {blocks} CTAs, overlapped", symm, True, variant, blocks))
                cases.append((f"switch v{variant} {blocks} CTAs, in line", symm, False, variant, blocks))
    for name, s, overlap, variant, blocks in cases:
        prob._symm, prob.overlap_allreduce = s, overlap
        if s is not None:
            s.max_blocks = min(blocks, cap)
        lib.hf_debug_allreduce_variant(variant)
        ms = timed(lambda: solve(prob, g, M), 3)
        say(f"solve, {name}: {ms:.2f} ms = {1e3 * ms / B.K_CG:.0f} us per iteration, {B.K_CG / ms * 1e3:.1f} products/s")
    lib.hf_debug_allreduce_variant(3)
    local_prob, lg, lM = problem(None)
    ms = timed(lambda: solve(local_prob, lg, lM), 3)
    say(f"solve, no exchange (compute floor of this rank count): {ms:.2f} ms = {1e3 * ms / B.K_CG:.0f} us per iteration")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
