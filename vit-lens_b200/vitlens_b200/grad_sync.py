"""Bucketed gradient all-reduce overlapped with the backward pass (data-parallel training of the hot path).

The reference wraps the model in torch DDP (training/point_cloud/pc_tri_main.py:378-380, training/main.py:318-324): gradients
are averaged over ranks in ~25 MB buckets while backward is still running.  Here the same exchange is written against the
engine's autograd Functions: parameters are grouped into buckets in reverse registration order (the order backward produces
their gradients); a post-accumulate hook counts a bucket down and, when its last gradient has landed, packs the bucket into
one flat fp32 buffer and issues ONE asynchronous NCCL all-reduce for it on a side stream, so the transfer of block i's
gradients overlaps the compute of blocks i-1, i-2, ...  `finish()` waits for the outstanding reductions and re-points each
`p.grad` at its slice of the flat buffer (no copy back); the mean's 1/world is folded into the optimizer's grad_scale.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("params", "flat", "offsets", "pending", "work", "ready")

    def __init__(self, params, device):
        self.params = params
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, device=device, dtype=torch.float32)
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

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 96 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        plist = [p for p in params if p.requires_grad]
        self.device = plist[0].device
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        for p in reversed(plist):  # backward order
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self.buckets.append(_Bucket(cur, self.device))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(_Bucket(cur, self.device))
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
        if self.stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                b.work = dist.all_reduce(b.flat, group=self.group, async_op=True)
        else:
            b.work = dist.all_reduce(b.flat, group=self.group, async_op=True)

    # ------------------------------------------------------------------ step boundary
    def finish(self):
        """Wait for the outstanding reductions; afterwards every p.grad is the cross-rank SUM (a view of its bucket)."""
        for b in self.buckets:
            if b.pending != 0:
                raise RuntimeError("GradReducer.finish: a bucket is missing gradients (a parameter did not take part in backward)")
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
