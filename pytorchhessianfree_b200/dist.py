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


# Whether the per-iteration exchange of the solve goes through hf_allreduce_multimem when the fabric allows it (else, and
# with HF_NVLS=0, ncclAllReduce): 8 B200, 11.3 MB vector: 41 us against 86 us alone, 786 against 754 products/s in the solve
# (tools/scale_probe.py, profiles/r2_summary.md).
NVLS_DEFAULT = True


def padded_range(begin, end, quantum, limit):
    """[begin, end) widened to multiples of ``quantum`` (= 4 floats x world: every rank reduces a whole number of 16-byte
    groups), clipped to ``limit`` (itself a multiple of ``quantum``).  Two ranges that share an end point which is a
    multiple of ``quantum`` stay disjoint -- what the sliced (overlapped) form of the exchange relies on."""
    return begin // quantum * quantum, min(limit, (end + quantum - 1) // quantum * quantum)


class SymmetricVector:
    """A flat FP32 vector in symmetric memory (same buffer on every rank of ``group``, peer-mapped, behind one NVSwitch
    multicast address) with an in-place all-reduce(sum) through the switch: ``hf_allreduce_multimem`` (csrc/collective.cu,
    hand-written ``multimem.ld_reduce`` / ``multimem.st``).  ``torch.distributed._symmetric_memory`` allocates the buffer
    and exchanges the handles -- plumbing; the collective itself is this repository's kernel.  ``try_create`` returns None
    when the process group, the driver or the fabric cannot provide a multicast mapping (the caller then keeps NCCL)."""

    def __init__(self, numel, device, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        self.lib = _lib.load()
        self.world = dist.get_world_size(group)
        self.quantum = 4 * self.world  # float4 slices per rank
        self.numel = int(numel)
        padded = (self.numel + self.quantum - 1) // self.quantum * self.quantum
        self.buf = symm_mem.empty(padded, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        if not self.hdl.multicast_ptr:
            raise RuntimeError("no multicast mapping for the symmetric buffer")
        self.rank = self.hdl.rank
        # 8-16 CTAs of 512 threads keep the links busy; more only add contention (8: 40.8 us, 64: 45.8 us at 8 ranks)
        self.max_blocks = max(1, min(16, self.hdl.signal_pad_size // (4 * self.world)))
        self.vec = self.buf[: self.numel]
        self.hdl.barrier()  # every rank has zeroed its buffer before anybody reduces

    _cache = {}

    @staticmethod
    def try_create(numel, device, group, force=False):
        """One symmetric vector per (group, length, device), created collectively on first use and shared by every
        problem of the process afterwards (a rendezvous per optimizer step would cost more than it saves).
        ``HF_NVLS=1`` / ``0`` switches the solver's use of it on / off (default: see NVLS_DEFAULT); ``force`` is for
        callers that want the collective itself (tests, tools)."""
        import os

        import torch.distributed as dist

        want = force or os.environ.get("HF_NVLS", "1" if NVLS_DEFAULT else "0") == "1"
        if group is None or not want or dist.get_backend(group) != "nccl":
            return None
        key = (getattr(group, "group_name", id(group)), int(numel), str(device))
        if key not in SymmetricVector._cache:
            try:
                SymmetricVector._cache[key] = SymmetricVector(numel, device, group)
            except Exception:  # noqa: BLE001 -- no symmetric memory / multicast here: NCCL stays
                SymmetricVector._cache[key] = None
        return SymmetricVector._cache[key]

    def padded_range(self, begin, end):
        """[begin, end) of the vector widened to the collective's granularity (the padding holds zeros on every rank)."""
        return padded_range(begin, end, self.quantum, self.buf.numel())

    def all_reduce_(self, begin=0, end=None, skip_ptr=None, stream=None):
        """In-place sum over the ranks of elements [begin, end) (widened with ``padded_range``), on ``stream``."""
        from . import _lib

        lo, hi = self.padded_range(begin, self.numel if end is None else end)
        _lib.check(self.lib.hf_allreduce_multimem(self.hdl.multicast_ptr, self.hdl.signal_pad_ptrs_dev, self.rank, self.world, lo, hi - lo,
                                                  self.max_blocks, skip_ptr, stream if stream is not None else _lib.stream()))
