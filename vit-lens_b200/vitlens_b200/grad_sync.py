"""Bucketed gradient all-reduce overlapped with the backward pass (data-parallel training of the hot path).

The reference wraps the model in torch DDP (training/point_cloud/pc_tri_main.py:378-380, training/main.py:318-324): gradients
are averaged over ranks in ~25 MB buckets while backward is still running.  Here the same exchange is written against the
engine's autograd Functions: parameters are grouped into buckets in reverse registration order (the order backward produces
their gradients); a post-accumulate hook counts a bucket down and, when its last gradient has landed, packs the bucket into
one flat fp32 buffer and issues ONE asynchronous NCCL all-reduce for it on a side stream, so the transfer of block i's
gradients overlaps the compute of blocks i-1, i-2, ...  `finish()` waits for the outstanding reductions and re-points each
`p.grad` at its slice of the flat buffer (no copy back); the mean's 1/world is folded into the optimizer's grad_scale.

Two transports.  NCCL: one asynchronous all-reduce per bucket on a side stream (its channel CTAs share the SMs with the
backward GEMMs).  Peer memory (vitlens_b200.comm, the default on one NVLink box): the bucket buffers live in this rank's slot of
a [world, total] gradient region of the peer arena; when a bucket is ready the COPY ENGINES push it into the same slot of every
peer's arena (vl_allreduce_grads: no SM-resident collective kernel next to the persistent GEMMs), and the reduction itself is
folded into the fused AdamW, which adds the world slots in rank order (`n_src` / `src_stride`, bit-identical on every rank).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("params", "flat", "offsets", "pending", "work", "ready", "start", "index")

    def __init__(self, params, device, flat=None):
        self.params = params
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, device=device, dtype=torch.float32) if flat is None else flat
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += p.numel()
        self.pending = len(params)
        self.work = None
        self.ready = None


class GradReducer:
    """Overlapped, bucketed SUM all-reduce of the gradients of `params` (divide by world size in the optimizer).

    Usage per step:  loss.backward()  ->  reducer.finish()  ->  optimizer.step(grad_scale=1/world)  ->  zero_grad().
    Gradient accumulation (several backward passes per optimizer step, training/train.py:154-210): run every backward except
    the last under `with reducer.no_sync():` -- like DDP's no_sync the hooks then leave the buckets alone, gradients pile up in
    p.grad, and the last backward reduces the accumulated values.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 96 << 20, group=None, arena=None):
        """arena: a vitlens_b200.comm.PeerArena with room for world x (total gradient bytes) -> copy-engine transport; None -> NCCL."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        plist = [p for p in params if p.requires_grad]
        self.device = plist[0].device
        self.buckets: List[_Bucket] = []
        groups, cur, cur_bytes = [], [], 0
        for p in reversed(plist):  # backward order
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                groups.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            groups.append(cur)
        self.total = sum(p.numel() for p in plist)
        self.arena = arena if (arena is not None and self.world > 1) else None
        self.n_src, self.src_stride, self.step = 1, 0, 0
        if self.arena is not None:
            from . import comm

            if len(groups) > comm.MAX_GRAD_FLAGS:
                raise ValueError(f"{len(groups)} gradient buckets exceed the {comm.MAX_GRAD_FLAGS} peer flags: raise bucket_bytes")
            self.stride_elems = (self.total + 63) // 64 * 64
            self.region = self.arena.alloc(self.world * self.stride_elems * 4)
            self.n_src, self.src_stride = self.world, self.stride_elems
        start = 0
        for i, g in enumerate(groups):
            n = sum(p.numel() for p in g)
            flat = None
            if self.arena is not None:  # this rank's slot of the region; slot q of every arena holds rank q's gradients
                flat = self.arena.tensor(self.region + (self.arena.rank * self.stride_elems + start) * 4, (n,), torch.float32)
            b = _Bucket(g, self.device, flat)
            b.start, b.index = start, i
            self.buckets.append(b)
            start += n
        self._of = {}
        for b in self.buckets:
            for p in b.params:
                self._of[p] = b
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._armed = True
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in plist]

    def no_sync(self):
        """Context manager: backward passes inside it only accumulate into p.grad (no bucket is counted down or reduced)."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old, self._armed = self._armed, False
            try:
                yield
            finally:
                self._armed = old

        return ctx()

    # ------------------------------------------------------------------ hooks
    def _on_grad(self, p):
        if not self._armed:
            return
        b = self._of[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        if self.world == 1:
            return
        # pack on the compute stream (one kernel), reduce on the side stream once the pack has finished
        torch.cat([p.grad.reshape(-1) for p in b.params], out=b.flat)
        if self.arena is not None:
            from . import comm
            from . import lib as L

            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                L.allreduce_grads(self.region + (self.arena.rank * self.stride_elems + b.start) * 4, b.flat.numel() * 4,
                                  comm.FLAG_GRAD0 + b.index, self.step + 1)
            b.work = True
            return
        if self.stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                b.work = dist.all_reduce(b.flat, group=self.group, async_op=True)
        else:
            b.work = dist.all_reduce(b.flat, group=self.group, async_op=True)

    # ------------------------------------------------------------------ step boundary
    def finish(self, materialize: bool = False):
        """Wait for the outstanding exchanges.  NCCL transport: afterwards every p.grad is the cross-rank SUM (a view of its
        bucket).  Peer-memory transport: every p.grad is a view of slot 0 of the gradient region and the sum over the world slots
        (`self.n_src` copies `self.src_stride` elements apart) is taken inside the optimizer -- pass both to AdamW.step();
        `materialize=True` adds the slots up into fresh p.grad tensors instead (for optimizers that cannot, and for checks)."""
        for b in self.buckets:
            if b.pending != 0:
                raise RuntimeError("GradReducer.finish: a bucket is missing gradients (a parameter did not take part in backward)")
        if self.arena is not None:
            from . import comm
            from . import lib as L

            self.step += 1
            for b in self.buckets:
                L.comm_wait(comm.FLAG_GRAD0 + b.index, self.step)  # every peer's copy of this bucket has landed in this arena
                if materialize:
                    n = b.flat.numel()
                    total = self.arena.tensor(self.region + b.start * 4, (n,), torch.float32).clone()
                    for q in range(1, self.world):  # rank order, like the optimizer kernel
                        total += self.arena.tensor(self.region + (q * self.stride_elems + b.start) * 4, (n,), torch.float32)
                for p, off in zip(b.params, b.offsets):
                    if materialize:
                        p.grad = total[off:off + p.numel()].view_as(p)
                    else:
                        p.grad = self.arena.tensor(self.region + (b.start + off) * 4, tuple(p.shape), torch.float32)
                b.pending = len(b.params)
                b.work = None
            return
        for b in self.buckets:
            if self.world > 1:
                b.work.wait()  # makes the current stream wait for the NCCL stream
                for p, off in zip(b.params, b.offsets):
                    p.grad = b.flat[off:off + p.numel()].view_as(p)
            b.pending = len(b.params)
            b.work = None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
