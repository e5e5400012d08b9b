"""The linear system of one Newton step, resident on the device.

A :class:`NativeProblem` bundles what the reference spreads over three closures
(``forward``, ``grad``, ``mvp``; ``optimizer.py:584-597``): the lowered net, the linearisations of
this rank's data chunks, the flat parameter vector, and -- when a process group is given -- the one
exchange step of the data-parallel path: an all-reduce(sum) of the flat FP32 vector after the local
chunk sum (gradient once per step, curvature product once per CG iteration, candidate losses once
per batch of candidates).  Every rank then runs the identical fused vector update, so the replicas
stay in lock-step without any scalar collective.
"""
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .dist import SymmetricVector
from .dist import all_reduce_sum as _all_reduce
from .dist import global_count
from .native import Linearization, NativeNet


class NativeProblem:
    def __init__(self, net: NativeNet, theta: torch.Tensor, curvature_opt: str,
                 mvp_data: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                 grad_data: Optional[Sequence] = None, loss_data: Optional[Sequence] = None, group=None):
        self.net, self.theta, self.curvature_opt, self.group = net, theta, curvature_opt, group
        self.lib = _lib.load()
        self.device = theta.device
        hess = curvature_opt == "hessian"
        # Chunk products in flight at a time (see _local_products); HF_CHUNK_LANES=1 runs them one after the other.
        self.chunk_lanes = max(1, int(os.environ.get("HF_CHUNK_LANES", "4"))) if theta.is_cuda else 1
        self.mvp_lins: List[Linearization] = [net.linearize(x, t, hessian=hess) for x, t in mvp_data]
        self.grad_lins = self.mvp_lins if grad_data is None else [net.linearize(x, t) for x, t in grad_data]
        # candidate losses (line search, backtracking, LM ratio) run on their own loss-only linearisations: two
        # ping-pong buffers per chunk, tensor-core tiles, and the curvature linearisation stays intact
        self.loss_lins = [net.linearize(l.x, l.targets, loss_only=True) for l in self.mvp_lins] if loss_data is None \
            else [net.linearize(x, t, loss_only=True) for x, t in loss_data]
        self.n_mvp, self.n_grad, self.n_loss = (self._count(l) for l in (self.mvp_lins, self.grad_lins, self.loss_lins))
        self._cand = torch.empty_like(theta)
        # Overlapping the all-reduce of the upper layers' slices with the first layer's weight gradient (see matvec) is an
        # option, off by default: timed in ONE process group on 8 B200 (tools/scale_probe.py, P = 2.8 M autoencoder) the
        # in-line exchange won both on NCCL (754 vs 739 products/s) and through the switch (786 vs 774): the collective
        # takes SMs and L2 bandwidth from the contraction it hides behind, and is two launches instead of one.
        # HF_OVERLAP_ALLREDUCE=1 turns it on.
        self.overlap_allreduce = os.environ.get("HF_OVERLAP_ALLREDUCE") == "1"
        off, cnt = net.first_layer_span()
        # the overlap split needs the first layer's slice to be a 16-byte aligned prefix of the flat vector
        self._split_at = cnt if (off == 0 and 0 < cnt < theta.numel() and cnt % 4 == 0) else 0
        self._linearized = False
        # The per-iteration exchange goes through the NVSwitch (hf_allreduce_multimem) when the curvature product can be
        # written straight into a symmetric buffer (`out_buffer`); everything else (gradient, Fisher diagonal, candidate
        # losses: once per step) stays on NCCL.
        self._symm = SymmetricVector.try_create(theta.numel(), self.device, group) if theta.is_cuda else None
        self._side = None
        self._lanes = []  # (stream, buffer) of lanes 1, 2, ...; lane 0 is the launching stream and `out`

    def out_buffer(self):
        """Where the solver should let :meth:`matvec` write its products so that their all-reduce can run through the
        switch: a view of the symmetric vector, or None (any tensor works then, over NCCL)."""
        return self._symm.vec if self._symm is not None else None

    def _count(self, lins):
        return global_count(sum(l.n for l in lins), self.group, self.device)

    # ---- once per step -----------------------------------------------------------------------------
    def linearize(self):
        """Forward passes at ``theta`` for the curvature (and gradient) chunks; returns the loss over
        the curvature data as a float64 device scalar."""
        acc = torch.zeros(1, dtype=torch.float64, device=self.device)
        for lin in self.mvp_lins:
            lin.forward(self.theta, self.n_mvp, acc)
        if self.grad_lins is not self.mvp_lins:
            for lin in self.grad_lins:
                lin.forward(self.theta, self.n_grad, None)
        _all_reduce(acc, self.group)
        self._linearized = True
        return acc

    def gradient(self):
        """Flat gradient over the gradient chunks (``optimizer.py:725-765``); also primes the Hessian path."""
        assert self._linearized
        # zeros, not empty: a rank whose shard is empty (fewer chunks than ranks) still joins the all-reduce
        g = torch.empty_like(self.theta) if self.grad_lins else torch.zeros_like(self.theta)
        for i, lin in enumerate(self.grad_lins):
            lin.gradient(self.theta, g, accumulate=i > 0)
        if self.curvature_opt == "hessian" and self.grad_lins is not self.mvp_lins:
            scratch = torch.empty_like(self.theta)
            for lin in self.mvp_lins:  # the Hessian product needs delta_l of its own chunks
                lin.gradient(self.theta, scratch, accumulate=False)
        _all_reduce(g, self.group)
        return g

    def fisher_diag(self):
        """Empirical-Fisher diagonal over the curvature chunks."""
        assert self._linearized
        d = torch.empty_like(self.theta) if self.mvp_lins else torch.zeros_like(self.theta)
        for i, lin in enumerate(self.mvp_lins):
            lin.fisher(self.theta, d, accumulate=i > 0)
        _all_reduce(d, self.group)
        return d

    # ---- once per CG iteration ---------------------------------------------------------------------
    def matvec(self, v, out, skip_ptr=None):
        """``out = B v`` (no damping), enqueued without host synchronisation.

        Data-parallel: one all-reduce(sum) of the flat vector after the local chunk sum -- through the NVSwitch with this
        library's own kernel (hf_allreduce_multimem) when ``out`` is :meth:`out_buffer`, else ncclAllReduce.  With
        ``overlap_allreduce`` (opt-in) and one chunk per rank the sweep runs in two phases and the all-reduce of the
        upper layers' slices (final after phase 0) overlaps with the first layer's weight gradient, the last and
        largest contraction of the sweep.  Every form sums the same per-rank vectors once per element and hands every
        rank the same bits, so replicas stay bit-identical."""
        nvls = self._symm is not None and out.data_ptr() == self._symm.vec.data_ptr()
        if nvls and self.overlap_allreduce and len(self.mvp_lins) == 1 and self._split_at > 0:
            # switch all-reduce of the upper layers' slices on a side stream, under the first layer's weight gradient
            symm, lin = self._symm, self.mvp_lins[0]
            cut = symm.padded_range(0, self._split_at)[1]  # >= the first layer's span: final after phase 0 from here on
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            lin.matvec_phase(self.curvature_opt, self.theta, v, out, 0, skip_ptr=skip_ptr)
            self._side.wait_stream(main)
            symm.all_reduce_(cut, symm.numel, skip_ptr, stream=self._side.cuda_stream)
            lin.matvec_phase(self.curvature_opt, self.theta, v, out, 1, skip_ptr=skip_ptr)
            main.wait_stream(self._side)
            symm.all_reduce_(0, cut, skip_ptr)
            return
        if nvls:
            self._local_products(v, out, skip_ptr)
            self._symm.all_reduce_(0, self._symm.numel, skip_ptr)
            return
        if self.overlap_allreduce and self.group is not None and len(self.mvp_lins) == 1 and self._split_at > 0:
            import torch.distributed as dist

            lin, n0 = self.mvp_lins[0], self._split_at
            lin.matvec_phase(self.curvature_opt, self.theta, v, out, 0, skip_ptr=skip_ptr)
            upper = dist.all_reduce(out[n0:], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            lin.matvec_phase(self.curvature_opt, self.theta, v, out, 1, skip_ptr=skip_ptr)
            first = dist.all_reduce(out[:n0], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            upper.wait()
            first.wait()
            return
        self._local_products(v, out, skip_ptr)
        _all_reduce(out, self.group)

    def _local_products(self, v, out, skip_ptr):
        """``out`` = sum over this rank's chunks of the chunk's curvature product.

        With two chunks or more, up to ``chunk_lanes`` (4) products are in flight: lane 0 accumulates its chunks into ``out``
        on the launching stream, every other lane into a buffer of its own on its own stream, and the lanes are added
        in a fixed order (deterministic).
        The contractions of a 7 500-row chunk are 120 pair tiles on 74 SM pairs: the second wave of every launch leaves
        28 pairs idle, and the chain of one product is sequential layer by layer -- the other lane's launches fill them."""
        lins = self.mvp_lins
        if not lins:
            out.zero_()  # empty shard: contribute nothing to the sum over ranks
            return
        hess = self.curvature_opt == "hessian"

        def product(lin, dst, accumulate):
            if hess:
                lin.hessian(self.theta, v, dst, accumulate=accumulate, skip_ptr=skip_ptr)
            else:
                lin.ggn(self.theta, v, dst, accumulate=accumulate, skip_ptr=skip_ptr)

        n_lanes = min(self.chunk_lanes, len(lins))
        if n_lanes < 2:
            for i, lin in enumerate(lins):
                product(lin, out, i > 0)
            return
        main = torch.cuda.current_stream()
        while len(self._lanes) < n_lanes - 1:
            self._lanes.append((torch.cuda.Stream(), torch.empty_like(self.theta)))
        for stream, _ in self._lanes[:n_lanes - 1]:
            stream.wait_stream(main)  # the direction is ready; the lane's buffer is no longer being read
        for i, lin in enumerate(lins):
            lane = i % n_lanes
            if lane == 0:
                product(lin, out, i >= n_lanes)
            else:
                stream, buf = self._lanes[lane - 1]
                with torch.cuda.stream(stream):
                    product(lin, buf, i >= n_lanes)
        for stream, buf in self._lanes[:n_lanes - 1]:
            main.wait_stream(stream)
            out.add_(buf)  # fixed order; (after the solver's skip flag is up nothing reads `out` any more)

    def mvp(self, v):
        """Tensor-in / tensor-out form of :meth:`matvec` (the reference's ``mvp`` plug-in signature)."""
        out = torch.empty_like(self.theta)
        self.matvec(_lib.vec(v.detach().to(torch.float32)), out)
        return out

    # ---- step selection ----------------------------------------------------------------------------
    def losses_at(self, steps):
        """Loss over the loss chunks at ``theta + s`` for every candidate ``s``: one device pass, one
        all-reduce, one device-to-host copy for the whole batch of candidates."""
        acc = torch.zeros(len(steps), dtype=torch.float64, device=self.device)
        for k, s in enumerate(steps):
            s = _lib.vec(s.detach().to(torch.float32))
            _lib.check(self.lib.hf_axpy_out(_lib.HF_F32, self.theta.numel(), self.theta.data_ptr(), 1.0, s.data_ptr(),
                                            self._cand.data_ptr(), _lib.stream()))
            for lin in self.loss_lins:
                lin.forward(self._cand, self.n_loss, acc[k:])
        _all_reduce(acc, self.group)
        return acc.to(torch.float32).tolist()

    def target_function(self):
        """``tfunc`` of ``optimizer.py:290-294`` with a batched variant attached as ``.many``.

        Values are remembered per candidate tensor (address + version) for the life of the function object: damping
        adaptation, backtracking and the line search ask for the same few candidates again (the last iterate three
        times), and every question is a device pass plus a host synchronisation.  ``.prime(steps)`` evaluates a
        list in one pass ahead of those questions."""
        memo = {}  # key -> (tensor, value): holding the tensor keeps its address from being recycled under the key

        def key(s):
            return (s.data_ptr(), s._version, s.numel())

        def many(steps):
            missing = {}
            for s in steps:
                if key(s) not in memo:
                    missing.setdefault(key(s), s)
            if missing:
                todo = list(missing.values())
                for s, v in zip(todo, self.losses_at(todo)):
                    memo[key(s)] = (s, v)
            return [memo[key(s)][1] for s in steps]

        def f(step):
            return many([step])[0]

        f.many = many
        f.prime = many
        return f
