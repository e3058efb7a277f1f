"""Shared pieces of the mask-loss parity tests (ClipLossSimMask / ClipLossLabelMask / TriClipLossLabelMask, loss.py:485-903): the
seeded inputs of oracle/make_golden_maskloss.py, the REAL reference's committed results (tests/golden/mask_loss.pt: world size 1
and per rank under two gloo processes), this repo's loss modules on the same inputs, and the oracle's statement."""
import os

import torch

from tests.common import ROOT  # noqa: F401  (sets sys.path)

GOLDEN = os.path.join(ROOT, "tests", "golden", "mask_loss.pt")
KINDS = ("sim", "label", "trilabel")
FLAGS = [(ll, gwg) for ll in (False, True) for gwg in (False, True)]


def load_golden():
    return torch.load(GOLDEN, map_location="cpu", weights_only=True)


def feature_blocks(seed, world, bl, e, dup=False):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(world, bl, e, generator=g)
    if dup:
        for r in range(world):
            t[r, 3] = t[0, 0] + 0.25 * torch.randn(e, generator=g)
            t[r, 5] = t[r, 1] + 0.25 * torch.randn(e, generator=g)
    return torch.nn.functional.normalize(t, dim=-1)


def label_blocks(seed, world, bl):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 5, (world, bl), generator=g)


def inputs(gold):
    W, bl, e = int(gold["world"]), int(gold["bl"]), int(gold["e"])
    return (feature_blocks(11, W, bl, e, dup=True), feature_blocks(12, W, bl, e), feature_blocks(13, W, bl, e),
            label_blocks(21, W, bl), label_blocks(22, W, bl), label_blocks(23, W, bl))


def run_ours(open_clip, gold, kind, kw, x, y, v, lx, ly, lv, device="cpu"):
    """This repo's loss module on one rank's blocks -> dict(loss, dx, dy, dv, ds) on the CPU."""
    x, y, v = (t.clone().to(device).requires_grad_(True) for t in (x, y, v))
    lx, ly, lv = (t.to(device) for t in (lx, ly, lv))
    s = torch.tensor(float(gold["scale_log"]), device=device, requires_grad=True)
    if kind == "sim":
        loss = open_clip.ClipLossSimMask(sim_thres=float(gold["sim_thres"]), **kw)(x, y, s.exp())
    elif kind == "label":
        loss = open_clip.ClipLossLabelMask(use_mask=True, **kw)(x, y, s.exp(), x_labels=lx, y_labels=ly)
    else:
        loss = open_clip.TriClipLossLabelMask(**kw)(x, y, v, s.exp(), image_labels=lx, text_labels=ly, visual_labels=lv)
    loss.backward()
    return dict(loss=loss.detach().cpu(), dx=x.grad.cpu(), dy=y.grad.cpu(), dv=v.grad.cpu() if kind == "trilabel" else None, ds=s.grad.cpu())


def flat(t):
    return t.reshape(-1, *t.shape[2:])


def compare(got, gold, prefix, tol_loss=1e-2, tol_grad=3e-2):
    """Max relative error per quantity; tolerance: bf16-rounded operands in the logits GEMMs."""
    worst = {}
    for k in ("loss", "ds", "dx", "dy", "dv"):
        if got.get(k) is None:
            continue
        ref = gold[f"{prefix}/{k}"]
        err = float((got[k] - ref).abs().max()) / float(ref.abs().max())
        worst[k] = err
        assert err <= (tol_loss if k == "loss" else tol_grad), (prefix, k, err)
    return worst


def oracle_per_rank(gold, kind, local_loss, gwg, world):
    """The oracle's restatement (oracle.clip_loss_sharded with mask=) of what every rank ends up with; world == 1 puts all rows on
    one rank."""
    from oracle import vitlens_oracle as O

    X, Y, V, LX, LY, LV = inputs(gold)
    if world == 1:
        X, Y, V, LX, LY, LV = (flat(t).unsqueeze(0) for t in (X, Y, V, LX, LY, LV))
    Xs, Ys, Vs = ([t.clone().requires_grad_(True) for t in T] for T in (X, Y, V))
    s = torch.tensor(float(gold["scale_log"]), requires_grad=True)
    ax = torch.cat([t.detach() for t in Xs], 0)
    if kind == "sim":
        pairs = [(Xs, Ys, O.sim_mask(ax, float(gold["sim_thres"])))]
    elif kind == "label":
        pairs = [(Xs, Ys, O.label_mask(flat(LX), flat(LY)))]
    else:
        pairs = [(Xs, Vs, O.label_mask(flat(LX), flat(LV))), (Ys, Vs, O.label_mask(flat(LY), flat(LV)))]
    losses = [sum(O.clip_loss_sharded(a, b, s.exp(), r, local_loss, gwg, mask=m) for a, b, m in pairs) for r in range(world)]
    out = []
    for r in range(world):
        ds = torch.autograd.grad(losses[r], s, retain_graph=True)[0]
        src = sum(losses) if gwg else losses[r]
        gx, gy, gv = torch.autograd.grad(src, [Xs[r], Ys[r], Vs[r]], retain_graph=True, allow_unused=True)
        out.append(dict(loss=losses[r].detach(), dx=gx, dy=gy, dv=gv if kind == "trilabel" else None, ds=ds))
    return out
