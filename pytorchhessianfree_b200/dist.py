"""Host-side plumbing of the data-parallel path (one process per GPU, ``torch.distributed``).

The Newton-step solve shards over the ``acc_step`` data dimension only: chunk ``c`` of a
global chunk list goes to rank ``c mod world``; parameters and all CG vectors are replicated; the single exchange
step is an all-reduce(sum) of a flat vector.  "mean" reductions divide by the GLOBAL sample count, which is why
every rank needs :func:`global_count` before it scales its local sums (reference semantics:
``optimizer.py:678-684`` computes ``sum_c N_c q_c / sum_c N_c`` over all chunks).
"""
import torch

from . import nccl_direct as _nccl_direct


def world(group=None):
    """(rank, world_size) of ``group``; (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_chunks(datalist, rank, world_size):
    """This rank's share of a global list of ``(inputs, targets)`` chunks: round-robin, order preserved."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return [chunk for c, chunk in enumerate(datalist) if c % world_size == rank]


def all_reduce_sum(t, group=None):
    """In-place sum over the ranks of ``group`` (no-op without a group)."""
    if group is not None:
        import torch.distributed as dist

        comm = None
        if t.is_cuda and t.is_contiguous() and t.dtype in _nccl_direct._DTYPES:
            comm = _nccl_direct.comm_for(group)
        if comm is not None:
            comm.all_reduce_sum(t)  # same collective, enqueued on the launching stream (see nccl_direct.py)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def global_count(n_local, group=None, device=None):
    """Sum of the ranks' local sample counts."""
    if group is None:
        return int(n_local)
    n = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    all_reduce_sum(n, group)
    return int(n.item())
