"""All-reduce on the CALLER'S stream through the NCCL library that torch already has loaded.

``torch.distributed.all_reduce`` runs the collective on ProcessGroupNCCL's own stream, bracketed by two event
hand-overs with the current stream.  On the data-parallel CG path the collective sits between two kernels of the
same stream once per iteration (curvature product -> all-reduce -> fused vector update), and those hand-overs cost
more than the 2.7 MB collective itself (measured on B200: ~60 us per iteration, of which ~25 us NCCL).  Here the
same ``ncclAllReduce`` is enqueued directly on the launching stream through a communicator of our own; the unique
id travels over the existing process group.  Plumbing only: no arithmetic of the path lives here.
"""
import atexit
import ctypes as C
import os

import torch

_DTYPES = {torch.float32: 7, torch.float64: 8, torch.int64: 4, torch.int32: 2, torch.uint8: 1}
_NCCL_SUM = 0


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


class DirectComm:
    def __init__(self, group):
        import torch.distributed as dist

        self.lib = C.CDLL("libnccl.so.2")  # resolves to the copy torch has already mapped
        self.lib.ncclGetErrorString.restype = C.c_char_p
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        self.lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = _UniqueId()
        if rank == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)))
        wire = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).cuda()
        dist.broadcast(wire, src=dist.get_global_rank(group, 0), group=group)
        uid = _UniqueId.from_buffer_copy(wire.cpu().numpy().tobytes())
        self.comm = C.c_void_p()
        self._check(self.lib.ncclCommInitRank(C.byref(self.comm), world, uid, rank))
        atexit.register(self.close)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"NCCL error {rc}: {self.lib.ncclGetErrorString(rc).decode()}")

    def all_reduce_sum(self, t):
        if not t.is_contiguous():
            raise ValueError("all_reduce_sum needs a contiguous tensor")
        self._check(self.lib.ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), _DTYPES[t.dtype], _NCCL_SUM, self.comm,
                                           torch.cuda.current_stream().cuda_stream))

    def close(self):
        comm, self.comm = self.comm, None
        if comm:
            try:
                self.lib.ncclCommDestroy(comm)
            except Exception:  # noqa: BLE001 -- interpreter shutdown
                pass


_comms = {}


def comm_for(group):
    """The direct communicator of an NCCL process group (created collectively on first use), or None."""
    import torch.distributed as dist

    if os.environ.get("HF_B200_TORCH_ALLREDUCE") == "1" or dist.get_backend(group) != "nccl":
        return None
    key = id(group)
    if key not in _comms:
        try:
            _comms[key] = DirectComm(group)
        except (OSError, AttributeError):  # library not loadable: keep torch's collective
            _comms[key] = None
    return _comms[key]
