"""Shared pieces of the distributed-loss parity tests: the seeded feature blocks of oracle/make_golden_dist.py, the
reference's committed per-rank results (tests/golden/dist_loss_w2.pt) and the oracle's statement of the same quantities."""
import os

import torch

from tests.common import ROOT  # noqa: F401  (sets sys.path)

GOLDEN = os.path.join(ROOT, "tests", "golden", "dist_loss_w2.pt")
COMBOS = [(tri, ll, gwg) for tri in (False, True) for ll in (False, True) for gwg in (False, True)]


def load_dist_golden():
    return torch.load(GOLDEN, map_location="cpu", weights_only=True)


def combo_name(tri, local_loss, gwg):
    return f"{'tri' if tri else 'clip'}_local{int(local_loss)}_gwg{int(gwg)}"


def feature_blocks(seed, world, bl, e):
    """[world, bl, e] unit-norm rows; block r belongs to rank r (same recipe as oracle/make_golden_dist.py)."""
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(world, bl, e, generator=g), dim=-1)


def oracle_per_rank(tri, local_loss, gwg, world, bl, e, scale_log, seeds=(1, 2, 3)):
    """What every rank ends up with after loss.backward() according to the oracle's restatement of loss.py:20-78,116-165:
    its loss value, the gradients of ITS feature blocks (all_gather's backward sums the other ranks' contributions when
    gather_with_grad) and d(its loss)/d(log logit_scale)."""
    from oracle import vitlens_oracle as O

    X = [t.clone().requires_grad_(True) for t in feature_blocks(seeds[0], world, bl, e)]
    Y = [t.clone().requires_grad_(True) for t in feature_blocks(seeds[1], world, bl, e)]
    V = [t.clone().requires_grad_(True) for t in feature_blocks(seeds[2], world, bl, e)]
    s = torch.tensor(float(scale_log), requires_grad=True)
    losses = []
    for r in range(world):
        if tri:
            losses.append(O.clip_loss_sharded(X, V, s.exp(), r, local_loss, gwg) + O.clip_loss_sharded(Y, V, s.exp(), r, local_loss, gwg))
        else:
            losses.append(O.clip_loss_sharded(X, Y, s.exp(), r, local_loss, gwg))
    out = []
    for r in range(world):
        ds = torch.autograd.grad(losses[r], s, retain_graph=True)[0]
        # gather_with_grad: rank r's block receives gradient from every rank's loss; otherwise only from its own
        src = sum(losses) if gwg else losses[r]
        gx, gy, gv = torch.autograd.grad(src, [X[r], Y[r], V[r]], retain_graph=True, allow_unused=True)
        out.append(dict(loss=losses[r].detach(), dx=gx, dy=gy, dv=gv if tri else None, ds=ds))
    return out
