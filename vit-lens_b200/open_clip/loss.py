"""Contrastive losses of the ViT-Lens hot path with the reference's constructors / forward signatures
(reference open_clip/loss.py:20-165, 234-385), computed by the fused tcgen05 row-LSE / gradient
epilogues (vitlens_b200.engine.ContrastiveFn): the [B x B_all] logits never reach HBM in forward.

Distributed semantics (loss.py:20-78) are reproduced for all four (local_loss, gather_with_grad)
combinations, but with ONE packed all-gather of the feature block per step instead of one per
tensor, and without every rank recomputing the full [Bg x Bg] matrix: each rank evaluates only its
[B_loc x Bg] row blocks and the ranks exchange the [B_loc] row-LSE vectors (see engine.ContrastiveFn).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from vitlens_b200 import engine as E

try:
    import torch.distributed.nn  # noqa: F401
    from torch import distributed as dist

    has_distributed = True
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1, use_horovod=False):
    """API-compatible feature all-gather (loss.py:20-78).  The loss modules below do not use it on their
    hot path (they gather one packed block); it is kept for callers that want the gathered tensors."""
    assert has_distributed, "torch.distributed did not import correctly"
    if use_horovod:
        raise NotImplementedError("horovod is not part of the B200 path; use torch.distributed (NCCL)")
    if gather_with_grad:
        all_image_features = torch.cat(torch.distributed.nn.all_gather(image_features), dim=0)
        all_text_features = torch.cat(torch.distributed.nn.all_gather(text_features), dim=0)
    else:
        gi = [torch.zeros_like(image_features) for _ in range(world_size)]
        gt = [torch.zeros_like(text_features) for _ in range(world_size)]
        dist.all_gather(gi, image_features)
        dist.all_gather(gt, text_features)
        if not local_loss:
            gi[rank] = image_features
            gt[rank] = text_features
        all_image_features = torch.cat(gi, dim=0)
        all_text_features = torch.cat(gt, dim=0)
    return all_image_features, all_text_features


class _ContrastiveBase(nn.Module):
    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1, use_horovod=False):
        super().__init__()
        if use_horovod:
            raise NotImplementedError("horovod is not part of the B200 path; use torch.distributed (NCCL)")
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.prev_num_logits = 0
        self.labels = {}
        self._arena_in_use = None

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        labels = torch.arange(num_logits, device=device, dtype=torch.long)
        if self.world_size > 1 and self.local_loss:
            labels = labels + num_logits * self.rank
        return labels

    # ---- distributed plumbing
    def _arena(self, Bl):
        """The peer-memory arena when this exchange can run over it (vitlens_b200.comm; all ranks on one NVLink box, row blocks
        the peer GEMMs can address in place), else None -> NCCL collectives."""
        from vitlens_b200 import comm, ops

        if not (has_distributed and dist.is_initialized()) or not ops.peer_rows_ok(Bl):
            return None
        return comm.init_arena()

    def _gather_tensors(self, feats):
        """NCCL / gloo gather of the packed block as plain tensors (the mask variants also need the gathered rows on the host side
        of the kernels to build their masks)."""
        W = self.world_size
        packed = torch.stack([f.detach().float() for f in feats], 0).contiguous()
        k, Bl, Ed = packed.shape
        out = torch.empty((W * k, Bl, Ed), device=packed.device, dtype=packed.dtype)
        dist.all_gather_into_tensor(out, packed)
        out = out.view(W, k, Bl, Ed)
        return [out[:, i].reshape(W * Bl, Ed) for i in range(k)]

    def _gather_packed(self, feats):
        """The packed [k, B_loc, E] feature block of every rank -> k matrices [Bg, E] (rank-major rows).  Over the peer arena
        nothing is gathered: each matrix is an ops.PeerRows whose row blocks the loss GEMMs read in place over NVLink; otherwise
        ONE all_gather_into_tensor of the packed block (the reference issues one all-gather per tensor, loss.py:55-76)."""
        W = self.world_size
        A = self._arena(feats[0].shape[0])
        if A is not None:
            return A.publish_features(feats)
        packed = torch.stack([f.detach().float() for f in feats], 0).contiguous()  # [k, B_loc, E]
        k, Bl, Ed = packed.shape
        out = torch.empty((W * k, Bl, Ed), device=packed.device, dtype=packed.dtype)
        dist.all_gather_into_tensor(out, packed)  # rank-major concat along dim 0
        out = out.view(W, k, Bl, Ed)
        return [out[:, i].reshape(W * Bl, Ed) for i in range(k)]

    def _gather_vec(self, v):
        A = self._arena_in_use
        if A is not None:
            return A.all_gather_vec(v)
        out = torch.empty((self.world_size * v.numel(),), device=v.device, dtype=v.dtype)
        dist.all_gather_into_tensor(out, v.contiguous())
        return out

    def _sum_ranks(self, t):
        """Cross-rank sum of a small tensor (loss value / d(scale) partial sums)."""
        A = self._arena_in_use
        if A is not None:
            return A.all_reduce_sum(t)
        t = t.clone()
        dist.all_reduce(t)
        return t

    def _pair(self, x, y, all_x, all_y, logit_scale, mask=None):
        """Loss of one (x, y) feature pair for this rank's rows; value follows the reference's definition.  `mask`: the full
        bool [B_all, B_all] keep-matrix of the mask variants (`logits * mask`), sliced here to this rank's rows."""
        from vitlens_b200 import ops

        W = self.world_size
        Bl = x.shape[0]
        masks = None
        if mask is not None:
            r0 = self.rank * Bl if W > 1 else 0
            masks = (mask[r0:r0 + Bl].to(torch.uint8).contiguous(), mask.t()[r0:r0 + Bl].to(torch.uint8).contiguous())
        if W == 1:
            return E.ContrastiveFn.apply(x, y, None, None, logit_scale, 0, Bl, Bl, True, None, None, False, masks)
        self._arena_in_use = all_x.arena if isinstance(all_x, ops.PeerRows) else None
        Bg = W * Bl
        off = self.rank * Bl
        if self.local_loss:
            col = bool(self.gather_with_grad)
            return E.ContrastiveFn.apply(x, y, all_x, all_y, logit_scale, off, Bl, Bl, col, self._gather_vec if col else None, None, col, masks)
        # full-matrix loss on every rank in the reference: value / d(scale) need cross-rank sums
        grad_rows = Bl if self.gather_with_grad else Bg

        def ds_post(ds):
            ds = self._sum_ranks(ds)
            return ds / W if self.gather_with_grad else ds

        part = E.ContrastiveFn.apply(x, y, all_x, all_y, logit_scale, off, Bg, grad_rows, True, self._gather_vec, ds_post, False, masks)
        total = self._sum_ranks(part.detach())
        return part + (total - part.detach())


class ClipLoss(_ContrastiveBase):
    """loss.py:311-385."""

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        all_i = all_t = None
        if self.world_size > 1:
            all_i, all_t = self._gather_packed([image_features, text_features])
        total_loss = self._pair(image_features, text_features, all_i, all_t, logit_scale)
        return {"contrastive_loss": total_loss} if output_dict else total_loss


class ClipLossGeneral(_ContrastiveBase):
    """loss.py:234-308."""

    def forward(self, x, y, logit_scale, output_dict=False, key="contrastive_loss"):
        all_x = all_y = None
        if self.world_size > 1:
            all_x, all_y = self._gather_packed([x, y])
        total_loss = self._pair(x, y, all_x, all_y, logit_scale)
        return {key: total_loss} if output_dict else total_loss


class TriClipLoss(_ContrastiveBase):
    """loss.py:81-165: pairs (image, visual) and (text, visual); four CE terms / 2."""

    def forward(self, image_features, text_features, visual_features, logit_scale, output_dict=False):
        all_i = all_t = all_v = None
        if self.world_size > 1:
            all_i, all_t, all_v = self._gather_packed([image_features, text_features, visual_features])
        total_loss = self._pair(image_features, visual_features, all_i, all_v, logit_scale) + \
            self._pair(text_features, visual_features, all_t, all_v, logit_scale)
        return {"contrastive_loss": total_loss} if output_dict else total_loss


# ----------------------------------------------------------------------------- mask variants (loss.py:485-903)
def _eye(n, device):
    return torch.eye(n, device=device, dtype=torch.bool)


class ClipLossSimMask(_ContrastiveBase):
    """loss.py:485-598: pairs whose TEACHER features x are more similar than `sim_thres` are taken out of the contrast --
    `logits * mask` with mask = not (x x^T >= sim_thres) or I.  The similarity matrix comes from the tcgen05 GEMM, the masking
    itself is applied inside the loss epilogues."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1, sim_thres=0.9, use_horovod=False):
        super().__init__(local_loss, gather_with_grad, cache_labels, rank, world_size, use_horovod)
        self.sim_thres = sim_thres

    def forward(self, x_features, y_features, logit_scale, output_dict=False, key="contrastive loss[with sim mask]"):
        all_x = all_y = None
        teacher = x_features
        if self.world_size > 1:
            all_x, all_y = self._gather_tensors([x_features, y_features])
            teacher = all_x
        t = teacher.detach().float().contiguous()
        sim = E._ops.similarity(t, gallery=t)
        mask = torch.logical_or(torch.logical_not(sim >= self.sim_thres), _eye(t.shape[0], t.device))
        total_loss = self._pair(x_features, y_features, all_x, all_y, logit_scale, mask)
        return {key: total_loss} if output_dict else total_loss


def _label_mask(x_labels, y_labels, n, device):
    if x_labels.ndim == 1:
        x_labels, y_labels = x_labels.unsqueeze(0), y_labels.unsqueeze(0)
    return torch.logical_or(torch.logical_not(x_labels.T == y_labels), _eye(n, device))


class ClipLossLabelMask(_ContrastiveBase):
    """loss.py:601-746: samples of the same class are not each other's negatives -- mask = (label_x != label_y^T) or I."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1, use_mask=False, use_horovod=False):
        super().__init__(local_loss, gather_with_grad, cache_labels, rank, world_size, use_horovod)
        self.use_mask = use_mask

    def _all_labels(self, lab):
        if self.world_size == 1:
            return lab
        out = [torch.zeros_like(lab) for _ in range(self.world_size)]
        dist.all_gather(out, lab.contiguous())
        return torch.cat(out, dim=0)

    def _masked_pair(self, x, y, logit_scale, x_labels, y_labels):
        all_x = all_y = None
        if self.world_size > 1:
            all_x, all_y = self._gather_tensors([x, y])
        mask = None
        if x_labels is not None and y_labels is not None and self.use_mask:
            assert x_labels.shape == y_labels.shape
            ax, ay = self._all_labels(x_labels), self._all_labels(y_labels)
            mask = _label_mask(ax, ay, ax.shape[0], x.device)
        return self._pair(x, y, all_x, all_y, logit_scale, mask)

    def forward(self, x_features, y_features, logit_scale, x_labels=None, y_labels=None, output_dict=False, key="image-text"):
        total_loss = self._masked_pair(x_features, y_features, logit_scale, x_labels, y_labels)
        return {key: total_loss} if output_dict else total_loss


class TriClipLossLabelMask(ClipLossLabelMask):
    """loss.py:749-903: the (image, visual) and (text, visual) pairs of TriClipLoss, each with its label mask."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1, use_horovod=False):
        super().__init__(local_loss, gather_with_grad, cache_labels, rank, world_size, True, use_horovod)

    def forward(self, image_features, text_features, visual_features, logit_scale, image_labels=None, text_labels=None, visual_labels=None,
                output_dict=False):
        total_loss = self._masked_pair(image_features, visual_features, logit_scale, image_labels, visual_labels) + \
            self._masked_pair(text_features, visual_features, logit_scale, text_labels, visual_labels)
        return {"tri_contrastive_loss": total_loss} if output_dict else total_loss
