"""AdamW over fp32 master weights in ONE fused launch per step (vl_adamw_multi), shaped like torch.optim.AdamW so that the
reference's training loop can use it unchanged: `param_groups` (the scheduler's assign_learning_rate writes
param_group["lr"], training/scheduler.py), `state_dict()` / `load_state_dict()` in torch's own layout (the reference
checkpoints optimizer.state_dict(), training/point_cloud/pc_tri_main.py:580-611), parameters without a gradient are
skipped.  Param groups as in the reference's launch scripts (pc_tri_main.py:394-419: no weight decay for ndim < 2 / bn /
ln / bias / logit_scale); the logit_scale clamp of training/train.py:248-249 is `clamp_logit_scale`."""
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
    copies of the weights (engine.WEIGHTS) so no separate cast kernels run in the next forward.

    Construct either from named parameters (the reference's decay / no-decay split is applied) or, like torch.optim.AdamW,
    from a list of parameters / of param-group dicts ({"params": [...], "weight_decay": ..., "lr": ...})."""

    def __init__(self, params: Iterable, lr=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2):
        params = list(params)
        if params and isinstance(params[0], dict):
            groups = [dict(g) for g in params]
        elif params and isinstance(params[0], (tuple, list)):
            no_decay, decay = split_decay(params)
            groups = [dict(params=no_decay, weight_decay=0.0), dict(params=decay, weight_decay=weight_decay)]
        else:
            groups = [dict(params=params)]
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = []
        for g in groups:
            g = dict(g)
            g["params"] = [p for p in g["params"]]
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            self.param_groups.append(g)
        self.t = 0
        self.grad_norm = None
        self._plan_key = None
        self._copied = None
        self._hyper_key = None
        self.m, self.v = {}, {}  # id(param) -> moment tensors (kept across plan rebuilds)

    # ---- torch.optim compatibility ------------------------------------------------------------------------------
    @property
    def groups(self):  # round-1 name
        return self.param_groups

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        for g in self.param_groups:
            g["lr"] = value

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none or p.grad is None:
                    p.grad = None
                else:
                    p.grad.zero_()

    def _all_params(self):
        return [p for g in self.param_groups for p in g["params"]]

    def state_dict(self):
        """torch.optim.AdamW's layout: {"state": {index: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]} with
        parameters replaced by their running index."""
        idx, state, groups = 0, {}, []
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                if id(p) in self.m:
                    state[idx] = {"step": torch.tensor(float(self.t)), "exp_avg": self.m[id(p)].clone(), "exp_avg_sq": self.v[id(p)].clone()}
                ids.append(idx)
                idx += 1
            groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        params = self._all_params()
        if sum(len(g["params"]) for g in sd["param_groups"]) != len(params):
            raise ValueError("loaded state dict has a different number of parameters")
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            if len(g["params"]) != len(saved["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            for k, v in saved.items():
                if k != "params":
                    g[k] = tuple(v) if k == "betas" else v
        steps = set()
        for i, st in sd["state"].items():
            p = params[int(i)]
            self.m[id(p)] = st["exp_avg"].to(device=p.device, dtype=torch.float32).clone()
            self.v[id(p)] = st["exp_avg_sq"].to(device=p.device, dtype=torch.float32).clone()
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the fused update keeps one step count")
        self.t = steps.pop() if steps else 0
        self._plan_key = None

    # ---- the fused step -----------------------------------------------------------------------------------------
    def _build_plan(self, params, key):
        dev = params[0].device
        self.params = params
        for p in params:
            if id(p) not in self.m or self.m[id(p)].device != p.device or self.m[id(p)].shape != p.shape:
                self.m[id(p)] = torch.zeros_like(p, dtype=torch.float32)
                self.v[id(p)] = torch.zeros_like(p, dtype=torch.float32)
        tab = []
        for i, p in enumerate(params):
            for c in range((p.numel() + L.ADAM_CHUNK - 1) // L.ADAM_CHUNK):
                tab.append((i, c))
        self.chunk_tab = torch.tensor(tab, dtype=torch.int32, device=dev).contiguous()
        self.sizes = torch.tensor([p.numel() for p in params], dtype=torch.int64, device=dev)
        self.ptr_host = torch.zeros((len(params), 5), dtype=torch.int64).pin_memory()
        self.ptr_host[:, 0] = torch.tensor([p.data_ptr() for p in params], dtype=torch.int64)
        self.ptr_host[:, 2] = torch.tensor([self.m[id(p)].data_ptr() for p in params], dtype=torch.int64)
        self.ptr_host[:, 3] = torch.tensor([self.v[id(p)].data_ptr() for p in params], dtype=torch.int64)
        self.ptr_dev = torch.zeros((len(params), 5), dtype=torch.int64, device=dev)
        self._plan_key = key
        self._hyper_key = None

    def _hyper(self, dev):
        """Per-tensor weight decay / learning rate tables; re-uploaded only when a param group changed (the scheduler)."""
        key = tuple((g["lr"], g["weight_decay"], len(g["params"])) for g in self.param_groups)
        if key != self._hyper_key:
            wds = [g["weight_decay"] for g in self.param_groups for _ in g["params"]]
            lrs = [g["lr"] for g in self.param_groups for _ in g["params"]]
            self.wds = torch.tensor(wds, dtype=torch.float32, device=dev)
            self.lrs = torch.tensor(lrs, dtype=torch.float32, device=dev)
            self._hyper_key = key
        return self.wds, self.lrs

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, clip_norm: float = None, n_src: int = 1, src_stride: int = 0):
        """One update.  grad_scale multiplies every gradient (1 / world_size after a summing all-reduce).  n_src / src_stride:
        every p.grad is the first of n_src copies src_stride elements apart that the kernel adds up in order (the peer-memory
        gradient exchange, grad_sync.GradReducer.n_src / .src_stride).  clip_norm: the
        reference's --grad-clip-norm (torch.nn.utils.clip_grad_norm_ over all parameters, train.py:212-240), applied to the
        scaled gradients inside the fused kernel; the total norm lands in self.grad_norm (a device scalar, no host sync)."""
        params = self._all_params()
        if not params:
            return
        # the plan holds raw device pointers: rebuild it when a parameter moved (model.to(), re-allocation) or the groups changed
        key = tuple(p.data_ptr() for p in params)
        if key != self._plan_key:
            self._build_plan(params, key)
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            if tuple(g["betas"]) != tuple(g0["betas"]) or g["eps"] != g0["eps"]:
                raise NotImplementedError("the fused update shares betas / eps between param groups (as every reference recipe does)")
        self.t += 1
        keep, gp, wp = [], [], []
        rows = self.ptr_host
        for p in self.params:
            g = p.grad
            if g is None:  # torch.optim skips such parameters (e.g. a positional embedding the forward did not use)
                gp.append(0)
                wp.append(0)
                continue
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
        wds, lrs = self._hyper(self.ptr_dev.device)
        sumsq = None
        if clip_norm is not None:
            sumsq = torch.empty((1,), dtype=torch.float32, device=self.ptr_dev.device)
            L.multi_sqnorm(self.ptr_dev, self.sizes, self.chunk_tab, sumsq, n_chunks=self.chunk_tab.shape[0], n_src=n_src, src_stride=src_stride)
            self.grad_norm = sumsq.sqrt() * grad_scale
        L.adamw_multi(self.ptr_dev, self.sizes, wds, self.chunk_tab, n_chunks=self.chunk_tab.shape[0], lr=g0["lr"], beta1=g0["betas"][0],
                      beta2=g0["betas"][1], eps=g0["eps"], step=self.t, grad_scale=grad_scale, sumsq=sumsq, max_norm=clip_norm, lrs=lrs,
                      n_src=n_src, src_stride=src_stride)
        engine.WEIGHTS.clear_derived()  # concatenated / padded / folded copies are rebuilt lazily; plain copies were refreshed above


def clamp_logit_scale(model, max_val=math.log(100)):
    with torch.no_grad():
        model.logit_scale.clamp_(0, max_val)
