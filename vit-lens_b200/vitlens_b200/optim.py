"""AdamW over fp32 master weights with the fused vl_adamw_step kernel, param groups as in the reference's
launch scripts (training/point_cloud/pc_tri_main.py:394-419: no weight decay for ndim < 2 / bn / ln / bias /
logit_scale), plus the logit_scale clamp of training/train.py:248-249."""
from __future__ import annotations

import math
from typing import Iterable

import torch

from . import engine
from . import lib as L


def split_decay(named_parameters):
    def exclude(n, p):
        return p.ndim < 2 or "bn" in n or "ln" in n or "bias" in n or "logit_scale" in n

    named = [(n, p) for n, p in named_parameters if p.requires_grad]
    return [p for n, p in named if exclude(n, p)], [p for n, p in named if not exclude(n, p)]


class AdamW:
    """All parameters are updated by ONE vl_adamw_multi launch per step; the same pass rewrites the cached bf16 operand
    copies of the weights (engine.WEIGHTS) so no separate cast kernels run in the next forward."""

    def __init__(self, named_parameters: Iterable, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2):
        no_decay, decay = split_decay(list(named_parameters))
        self.groups = [dict(params=no_decay, weight_decay=0.0), dict(params=decay, weight_decay=weight_decay)]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.t = 0
        self.grad_norm = None
        self._plan = None
        self._copied = None

    def zero_grad(self):
        for g in self.groups:
            for p in g["params"]:
                p.grad = None

    def _build_plan(self):
        params, wds = [], []
        for g in self.groups:
            for p in g["params"]:
                params.append(p)
                wds.append(g["weight_decay"])
        dev = params[0].device
        self.params = params
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        tab = []
        for i, p in enumerate(params):
            for c in range((p.numel() + L.ADAM_CHUNK - 1) // L.ADAM_CHUNK):
                tab.append((i, c))
        self.chunk_tab = torch.tensor(tab, dtype=torch.int32, device=dev).contiguous()
        self.sizes = torch.tensor([p.numel() for p in params], dtype=torch.int64, device=dev)
        self.wds = torch.tensor(wds, dtype=torch.float32, device=dev)
        self.ptr_host = torch.zeros((len(params), 5), dtype=torch.int64).pin_memory()
        self.ptr_host[:, 0] = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64)
        self.ptr_host[:, 2] = torch.tensor([t.data_ptr() for t in self.m], dtype=torch.int64)
        self.ptr_host[:, 3] = torch.tensor([t.data_ptr() for t in self.v], dtype=torch.int64)
        self.ptr_dev = torch.zeros((len(params), 5), dtype=torch.int64, device=dev)
        self._plan = True

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, clip_norm: float = None):
        """One update.  grad_scale multiplies every gradient (1 / world_size after a summing all-reduce).  clip_norm: the
        reference's --grad-clip-norm (torch.nn.utils.clip_grad_norm_ over all parameters, train.py:212-240), applied to the
        scaled gradients inside the fused kernel; the total norm lands in self.grad_norm (a device scalar, no host sync)."""
        if self._plan is None:
            self._build_plan()
        self.t += 1
        keep, gp, wp = [], [], []
        rows = self.ptr_host
        for p in self.params:
            g = p.grad
            if g is None:
                raise RuntimeError("AdamW.step: a parameter has no gradient (frozen parameters must not be passed to the optimizer)")
            if not g.is_contiguous() or g.dtype != torch.float32:
                g = g.contiguous().float()
                keep.append(g)
            w16 = engine.WEIGHTS.plain_copy(p)  # bf16 operand copy to refresh in place (None if this weight has none)
            gp.append(g.data_ptr())
            wp.append(0 if w16 is None else w16.data_ptr())
        if self._copied is not None:
            self._copied.synchronize()  # the previous step's async table upload must have left the pinned buffer
        rows[:, 1] = torch.tensor(gp, dtype=torch.int64)
        rows[:, 4] = torch.tensor(wp, dtype=torch.int64)
        self.ptr_dev.copy_(rows, non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record()
        sumsq = None
        if clip_norm is not None:
            sumsq = torch.empty((1,), dtype=torch.float32, device=self.ptr_dev.device)
            L.multi_sqnorm(self.ptr_dev, self.sizes, self.chunk_tab, sumsq, n_chunks=self.chunk_tab.shape[0])
            self.grad_norm = sumsq.sqrt() * grad_scale
        L.adamw_multi(self.ptr_dev, self.sizes, self.wds, self.chunk_tab, n_chunks=self.chunk_tab.shape[0], lr=self.lr, beta1=self.betas[0],
                      beta2=self.betas[1], eps=self.eps, step=self.t, grad_scale=grad_scale, sumsq=sumsq, max_norm=clip_norm)
        engine.WEIGHTS.clear_derived()  # concatenated / padded / folded copies are rebuilt lazily; plain copies were refreshed above


def clamp_logit_scale(model, max_val=math.log(100)):
    with torch.no_grad():
        model.logit_scale.clamp_(0, max_val)
