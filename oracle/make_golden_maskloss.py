"""Pin the MASK variants of the contrastive loss against the real reference  --  TEST INFRASTRUCTURE (build container only).

    python oracle/make_golden_maskloss.py

Runs the unmodified reference's `ClipLossSimMask`, `ClipLossLabelMask` and `TriClipLossLabelMask` (open_clip/loss.py:485-903)
on seeded feature blocks: once in-process at world_size 1 (all rows on one rank) and once as WORLD gloo processes for the four
(local_loss, gather_with_grad) combinations.  Per rank: loss, d(loss)/d(local features), d(loss)/d(log logit_scale)  ->
tests/golden/mask_loss.pt.  tests/test_oracle_golden.py pins oracle.clip_loss_sharded(mask=...) to it; the host-logic (CPU,
gloo) and GPU tests compare this repo's loss modules with it.
"""
from __future__ import annotations

import os
import socket
import sys

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle.make_golden_dist import _gloo_all_to_all  # noqa: E402  (the one collective gloo lacks)

WORLD, BL, E = 2, 8, 32
SCALE_LOG = 2.5
SIM_THRES = 0.8
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mask_loss.pt")
KINDS = ("sim", "label", "trilabel")


def feature_blocks(seed: int, world: int = WORLD, bl: int = BL, e: int = E, dup: bool = False) -> torch.Tensor:
    """[world, bl, e] unit-norm rows; block r belongs to rank r.  dup: rows 3, 5 of every block are near-copies of row 0 of block
    0 / row 1 of block r, so the similarity mask has off-diagonal zeros inside and across ranks."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(world, bl, e, generator=g)
    if dup:
        for r in range(world):
            t[r, 3] = t[0, 0] + 0.25 * torch.randn(e, generator=g)
            t[r, 5] = t[r, 1] + 0.25 * torch.randn(e, generator=g)
    return torch.nn.functional.normalize(t, dim=-1)


def label_blocks(seed: int, world: int = WORLD, bl: int = BL) -> torch.Tensor:
    """[world, bl] int64 class ids out of 5 classes (plenty of same-class pairs)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 5, (world, bl), generator=g)


def inputs():
    X, Y, V = feature_blocks(11, dup=True), feature_blocks(12), feature_blocks(13)
    LX, LY, LV = label_blocks(21), label_blocks(22), label_blocks(23)
    return X, Y, V, LX, LY, LV


def run_reference(open_clip, kind, kw, x, y, v, s, lx, ly, lv):
    L = open_clip.loss
    if kind == "sim":
        return L.ClipLossSimMask(sim_thres=SIM_THRES, **kw)(x, y, s.exp())
    if kind == "label":
        return L.ClipLossLabelMask(use_mask=True, **kw)(x, y, s.exp(), x_labels=lx, y_labels=ly)
    return L.TriClipLossLabelMask(use_mask=True, **kw)(x, y, v, s.exp(), image_labels=lx, text_labels=ly, visual_labels=lv)


def _one(open_clip, kind, kw, x, y, v, lx, ly, lv):
    x, y, v = (t.clone().requires_grad_(True) for t in (x, y, v))
    s = torch.tensor(SCALE_LOG, requires_grad=True)
    loss = run_reference(open_clip, kind, kw, x, y, v, s, lx, ly, lv)
    loss.backward()
    return dict(loss=loss.detach().clone(), dx=x.grad.clone(), dy=y.grad.clone(), dv=v.grad.clone() if kind == "trilabel" else None,
                ds=s.grad.clone())


def _worker(rank, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from oracle import ref_import

    open_clip, _, _ = ref_import.import_reference()
    dist.all_to_all = _gloo_all_to_all
    X, Y, V, LX, LY, LV = inputs()
    res = {}
    for kind in KINDS:
        for local_loss in (False, True):
            for gwg in (False, True):
                kw = dict(local_loss=local_loss, gather_with_grad=gwg, rank=rank, world_size=WORLD)
                res[(kind, local_loss, gwg)] = _one(open_clip, kind, kw, X[rank], Y[rank], V[rank], LX[rank], LY[rank], LV[rank])
    torch.save(res, out.format(rank))
    dist.destroy_process_group()


def main():
    from oracle import ref_import

    open_clip, _, _ = ref_import.import_reference()
    X, Y, V, LX, LY, LV = inputs()
    sim = X.reshape(-1, E) @ X.reshape(-1, E).t()
    off = sim[~torch.eye(sim.shape[0], dtype=torch.bool)]
    assert int((off >= SIM_THRES).sum()) >= 6, "the similarity mask must remove some pairs"
    assert float((off - SIM_THRES).abs().min()) > 0.03, "no similarity may sit at the threshold (bf16 operands on the device)"
    fx = {"world": WORLD, "bl": BL, "e": E, "scale_log": SCALE_LOG, "sim_thres": SIM_THRES}
    flat = lambda t: t.reshape(-1, *t.shape[2:])  # noqa: E731
    for kind in KINDS:
        r = _one(open_clip, kind, dict(world_size=1), flat(X), flat(Y), flat(V), flat(LX), flat(LY), flat(LV))
        for k, v in r.items():
            if v is not None:
                fx[f"{kind}_w1/{k}"] = v
        print(kind, "world 1", float(r["loss"]), float(r["ds"]))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    tmp = "/tmp/_mask_golden_r{}.pt"
    mp.spawn(_worker, args=(port, tmp), nprocs=WORLD, join=True)
    per_rank = [torch.load(tmp.format(r), weights_only=False) for r in range(WORLD)]
    for key in per_rank[0]:
        kind, ll, gwg = key
        name = f"{kind}_local{int(ll)}_gwg{int(gwg)}"
        for r in range(WORLD):
            for k, v in per_rank[r][key].items():
                if v is not None:
                    fx[f"{name}/rank{r}/{k}"] = v
        print(name, [float(per_rank[r][key]["loss"]) for r in range(WORLD)], [float(per_rank[r][key]["ds"]) for r in range(WORLD)])
    torch.save(fx, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
