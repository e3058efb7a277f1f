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
    def __init__(self, named_parameters: Iterable, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2):
        no_decay, decay = split_decay(list(named_parameters))
        self.groups = [dict(params=no_decay, weight_decay=0.0), dict(params=decay, weight_decay=weight_decay)]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.state = {}
        self.t = 0

    def zero_grad(self):
        for g in self.groups:
            for p in g["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        self.t += 1
        for g in self.groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self.state.get(id(p))
                if st is None:
                    st = self.state[id(p)] = (torch.zeros_like(p), torch.zeros_like(p))
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                L.adamw_step(p, grad, st[0], st[1], lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                             weight_decay=g["weight_decay"], step=self.t, grad_scale=grad_scale)
        engine.WEIGHTS.clear()  # master weights moved: bf16 operand copies are stale


def clamp_logit_scale(model, max_val=math.log(100)):
    with torch.no_grad():
        model.logit_scale.clamp_(0, max_val)
