"""One optimisation step with the reference's semantics (training/train.py:131-249; the point-cloud trainer
training/point_cloud/pc_tri_main.py:470-560 has the same body): forward -> contrastive loss -> backward -> optional
gradient-norm clip -> optimizer step -> logit_scale clamp, and the `accum_freq > 1` scheme in which the features of every
micro-batch are first cached WITHOUT gradients and each micro-batch is then re-forwarded with gradients against the cached
features of the others (train.py:154-210) -- the audio recipe relies on it to reach its global batch.

Not a trainer: data loading, LR schedule, logging and checkpointing stay with the caller (out of scope)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

FEATURE_KEYS = ("image_features", "text_features", "visual_features")


def _unwrap(model):
    return getattr(model, "module", model)


class TrainStep:
    """`model(*inputs)` must return the reference's output dict ({image,text[,visual]}_features + logit_scale; build the model
    with output_dict=True); `loss(**features, logit_scale=..., output_dict=True)` a dict of loss terms (ClipLoss / TriClipLoss).

    optimizer: anything with step() / zero_grad(); vitlens_b200.optim.AdamW additionally takes the clip and the 1 / world
    scale inside its fused kernel.  reducer: vitlens_b200.grad_sync.GradReducer (or None when DDP / a single process)."""

    def __init__(self, model, loss, optimizer, *, accum_freq: int = 1, grad_clip_norm: Optional[float] = None, reducer=None,
                 world_size: int = 1, logit_scale_max: float = math.log(100)):
        assert accum_freq >= 1
        self.model, self.loss, self.optimizer = model, loss, optimizer
        self.accum_freq, self.grad_clip_norm = accum_freq, grad_clip_norm
        self.reducer, self.world_size, self.logit_scale_max = reducer, world_size, logit_scale_max
        self._inputs: List[Sequence[torch.Tensor]] = []
        self._feats: Dict[str, List[torch.Tensor]] = {}
        self.last_losses: Optional[Dict[str, torch.Tensor]] = None

    # -- pieces -----------------------------------------------------------------------------------------------------
    def _loss(self, feats: Dict[str, torch.Tensor], logit_scale) -> torch.Tensor:
        losses = self.loss(**feats, logit_scale=logit_scale, output_dict=True)
        total = sum(losses.values())
        self.last_losses = {k: v.detach() for k, v in losses.items()}
        self.last_losses["loss"] = total.detach()
        return total

    def _finish(self):
        from . import optim as _optim

        if self.reducer is not None:
            self.reducer.finish(materialize=not isinstance(self.optimizer, _optim.AdamW))
        scale = 1.0 / self.world_size if self.reducer is not None else 1.0
        if isinstance(self.optimizer, _optim.AdamW):
            src = dict(n_src=self.reducer.n_src, src_stride=self.reducer.src_stride) if self.reducer is not None else {}
            self.optimizer.step(grad_scale=scale, clip_norm=self.grad_clip_norm, **src)
        else:
            params = [p for p in self.model.parameters() if p.grad is not None]
            if scale != 1.0:
                for p in params:
                    p.grad.mul_(scale)
            if self.grad_clip_norm is not None:
                torch.nn.utils.clip_grad_norm_(params, self.grad_clip_norm, norm_type=2.0)
            self.optimizer.step()
        with torch.no_grad():  # "we clamp to 4.6052 = ln(100), as in the original paper" (train.py:247-249)
            _unwrap(self.model).logit_scale.clamp_(0, self.logit_scale_max)

    # -- the step ---------------------------------------------------------------------------------------------------
    def __call__(self, *inputs: torch.Tensor) -> bool:
        """Feed one (micro-)batch.  Returns True when the optimizer stepped (every accum_freq-th call)."""
        if self.accum_freq == 1:
            self.optimizer.zero_grad()
            out = dict(self.model(*inputs))
            logit_scale = out.pop("logit_scale")
            self._loss({k: v for k, v in out.items() if k in FEATURE_KEYS}, logit_scale).backward()
            self._finish()
            return True
        # accumulate: cache this micro-batch's features without gradient tracking (train.py:154-170)
        with torch.no_grad():
            out = dict(self.model(*inputs))
            out.pop("logit_scale")
            for k, v in out.items():
                if k in FEATURE_KEYS:
                    self._feats.setdefault(k, []).append(v)
        self._inputs.append(inputs)
        if len(self._inputs) < self.accum_freq:
            return False
        # re-forward every micro-batch with gradients; the others' cached features are the negatives (train.py:176-210)
        self.optimizer.zero_grad()
        import contextlib

        for j, inp in enumerate(self._inputs):
            out = dict(self.model(*inp))
            logit_scale = out.pop("logit_scale")
            feats = {k: torch.cat(acc[:j] + [out[k]] + acc[j + 1:]) for k, acc in self._feats.items()}
            # only the LAST micro-batch's backward may hand the (by then fully accumulated) gradients to the reducer
            last = j == len(self._inputs) - 1
            with (contextlib.nullcontext() if (last or self.reducer is None) else self.reducer.no_sync()):
                self._loss(feats, logit_scale).backward()
        self._finish()
        self._inputs, self._feats = [], {}
        return True
