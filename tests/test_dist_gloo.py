"""World-size-2 (gloo, CPU) check of the distributed contrastive-loss logic: one packed feature all-gather,
row-LSE exchange, cross-rank d(logit_scale) -- against the oracle's statement of the reference's
gather_features/ClipLoss/TriClipLoss semantics for all four (local_loss, gather_with_grad) combinations.
The kernels are emulated (tests/emu_ops.py); what is under test is the host logic in open_clip/loss.py."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from tests.common import ROOT  # noqa: F401  (sets sys.path)

W, BL, E = 2, 6, 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _feats(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(W, BL, E, generator=g), dim=-1)


def _worker(rank, port, tri, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=W)
    from tests import emu_ops
    from vitlens_b200 import engine

    engine._ops = emu_ops
    import open_clip

    res = {}
    X, Y, V = _feats(1), _feats(2), _feats(3)
    for local_loss in (False, True):
        for gwg in (False, True):
            x = X[rank].clone().requires_grad_(True)
            y = Y[rank].clone().requires_grad_(True)
            v = V[rank].clone().requires_grad_(True)
            s = torch.tensor(2.5, requires_grad=True)
            if tri:
                loss = open_clip.TriClipLoss(local_loss=local_loss, gather_with_grad=gwg, rank=rank, world_size=W)(x, y, v, s.exp())
            else:
                loss = open_clip.ClipLoss(local_loss=local_loss, gather_with_grad=gwg, rank=rank, world_size=W)(x, y, s.exp())
            loss.backward()
            res[(local_loss, gwg)] = dict(loss=loss.detach(), dx=x.grad, dy=y.grad, dv=v.grad, ds=s.grad)
    torch.save(res, out.format(rank))
    dist.destroy_process_group()


def _reference(tri):
    """Per-rank losses and the gradients each rank ends up with, from the oracle's restatement of loss.py."""
    from oracle import vitlens_oracle as O

    res = {}
    for local_loss in (False, True):
        for gwg in (False, True):
            X = [t.clone().requires_grad_(True) for t in _feats(1)]
            Y = [t.clone().requires_grad_(True) for t in _feats(2)]
            V = [t.clone().requires_grad_(True) for t in _feats(3)]
            s = torch.tensor(2.5, requires_grad=True)
            losses = []
            for r in range(W):
                if tri:
                    lr = O.clip_loss_sharded(X, V, s.exp(), r, local_loss, gwg) + O.clip_loss_sharded(Y, V, s.exp(), r, local_loss, gwg)
                else:
                    lr = O.clip_loss_sharded(X, Y, s.exp(), r, local_loss, gwg)
                losses.append(lr)
            # all_gather's backward sums over ranks; DDP then averages parameter grads over ranks
            torch.stack(losses).sum().backward()
            res[(local_loss, gwg)] = dict(losses=[float(l) for l in losses], dX=[t.grad for t in X], dY=[t.grad for t in Y],
                                          dV=[t.grad for t in V], ds_mean=float(s.grad) / W)
    return res


@pytest.mark.parametrize("tri", [False, True])
def test_sharded_contrastive_loss_matches_reference_semantics(tri, tmp_path):
    port = _free_port()
    out = str(tmp_path / "r{}.pt")
    mp.spawn(_worker, args=(port, tri, out), nprocs=W, join=True)
    got = [torch.load(out.format(r), weights_only=False) for r in range(W)]
    ref = _reference(tri)
    for key, rr in ref.items():
        ds_mean = sum(float(got[r][key]["ds"]) for r in range(W)) / W
        assert abs(ds_mean - rr["ds_mean"]) < 2e-2 * abs(rr["ds_mean"]) + 1e-4, (key, ds_mean, rr["ds_mean"])
        for r in range(W):
            g = got[r][key]
            assert abs(float(g["loss"]) - rr["losses"][r]) < 1e-2 * abs(rr["losses"][r]), (key, r, float(g["loss"]), rr["losses"][r])
            pairs = [(g["dx"], rr["dX"][r]), (g["dv"] if tri else g["dy"], rr["dV"][r] if tri else rr["dY"][r])]
            if tri:
                pairs.append((g["dy"], rr["dY"][r]))
            for a, b in pairs:
                err = float((a - b).abs().max())
                assert err < 3e-2 * float(b.abs().max()) + 1e-5, (key, r, err, float(b.abs().max()))


def _reducer_worker(rank, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=W)
    from vitlens_b200.grad_sync import GradReducer

    torch.manual_seed(0)  # identical weights on both ranks
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(), torch.nn.Linear(64, 8))
    frozen = net[2].bias
    frozen.requires_grad_(False)
    params = [p for p in net.parameters()]
    red = GradReducer(params, bucket_bytes=4096)  # several buckets
    res = []
    for step in range(2):
        g = torch.Generator().manual_seed(10 * step + rank)
        x = torch.randn(5, 16, generator=g)
        net(x).square().sum().backward()
        local = [p.grad.clone() for p in params if p.requires_grad]
        red.finish()
        res.append(dict(local=local, reduced=[p.grad.clone() for p in params if p.requires_grad]))
        for p in params:
            p.grad = None
    torch.save(dict(res=res, n_buckets=len(red.buckets)), out.format(rank))
    dist.destroy_process_group()


def test_bucketed_gradient_all_reduce(tmp_path):
    """GradReducer (the overlapped DDP-style exchange, pc_tri_main.py:378-380): every p.grad ends up as the cross-rank sum,
    buckets are re-armed for the next step, frozen parameters are skipped."""
    out = str(tmp_path / "r{}.pt")
    mp.spawn(_reducer_worker, args=(_free_port(), out), nprocs=W, join=True)
    r = [torch.load(out.format(i)) for i in range(W)]
    assert r[0]["n_buckets"] > 1
    for step in range(2):
        for k in range(len(r[0]["res"][step]["local"])):
            want = r[0]["res"][step]["local"][k] + r[1]["res"][step]["local"][k]
            for rank in range(W):
                assert torch.allclose(r[rank]["res"][step]["reduced"][k], want, rtol=1e-6, atol=1e-6)


# ----------------------------------------------------------------------------- SyncBatchNorm in the point tokenizer
def _pc_tokenizer(sync):
    """The tiny point tokenizer of oracle case tiny_tri_pc (16 groups of 8, 64 encoder channels), deterministic weights."""
    from types import SimpleNamespace

    from open_clip.modal_3d.models.pointbert.point_encoder import PointTokenizer

    torch.manual_seed(0)
    tk = PointTokenizer(SimpleNamespace(trans_dim=96, group_size=8, num_group=16, encoder_dims=64))
    with torch.no_grad():  # non-trivial BatchNorm affines / running statistics
        for bn in (tk.encoder.first_conv[1], tk.encoder.second_conv[1]):
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.2, 0.2)
    if sync:
        tk = torch.nn.SyncBatchNorm.convert_sync_batchnorm(tk)
    return tk.train()


def _pc_inputs():
    g = torch.Generator().manual_seed(7)
    pts = torch.randn(4, 128, 3, generator=g)
    start = torch.randint(0, 128, (4,), generator=g)
    return pts, start


def _pc_run(tk, pts, start):
    out = tk(pts, fps_start=start)
    x, pos = out["x"].t.float(), out["pos"].t.float()
    w = torch.linspace(-1, 1, x.numel()).reshape(x.shape)
    (x * w).sum().add((pos * w.flip(0)).sum()).backward()
    grads = {k: p.grad.clone() for k, p in tk.named_parameters()}
    stats = {k: v.clone() for k, v in tk.state_dict().items() if "running" in k}
    return x.detach(), grads, stats


def _syncbn_worker(rank, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=W)
    from tests import emu_ops
    from vitlens_b200 import engine

    engine._ops = emu_ops
    tk = _pc_tokenizer(sync=True)
    pts, start = _pc_inputs()
    sl = slice(2 * rank, 2 * rank + 2)
    x, grads, stats = _pc_run(tk, pts[sl], start[sl])
    torch.save(dict(x=x, grads=grads, stats=stats), out.format(rank))
    dist.destroy_process_group()


def test_point_tokenizer_sync_batchnorm(tmp_path):
    """--use-bn-sync (pc_tri_main.py:372-373): with SyncBatchNorm layers two ranks holding half the clouds each produce the
    tokens of one process holding all of them, the same running statistics, and parameter gradients that SUM to its
    gradients (the weights of the loss below are per global row; DDP's averaging is a separate step)."""
    from tests import emu_ops
    from vitlens_b200 import engine

    out = str(tmp_path / "s{}.pt")
    mp.spawn(_syncbn_worker, args=(_free_port(), out), nprocs=W, join=True)
    r = [torch.load(out.format(i)) for i in range(W)]
    old = engine._ops
    engine._ops = emu_ops
    try:
        tk = _pc_tokenizer(sync=False)
        pts, start = _pc_inputs()
        # single process, all four clouds; per-rank losses used the same weight pattern on their own rows, so rebuild that
        outs, gsum = [], None
        full = tk(pts, fps_start=start)
        x = full["x"].t.float()
        pos = full["pos"].t.float()
        half = x.shape[0] // 2
        loss = 0
        for rk in range(W):
            xs, ps = x[rk * half:(rk + 1) * half], pos[rk * half:(rk + 1) * half]
            w = torch.linspace(-1, 1, xs.numel()).reshape(xs.shape)
            loss = loss + (xs * w).sum() + (ps * w.flip(0)).sum()
        loss.backward()
        ref_g = {k: p.grad for k, p in tk.named_parameters()}
        ref_s = {k: v for k, v in tk.state_dict().items() if "running" in k}
    finally:
        engine._ops = old
        engine.WEIGHTS.clear()
    got_x = torch.cat([r[0]["x"], r[1]["x"]])
    assert float((got_x - x.detach()).abs().max()) < 3e-2 * float(x.detach().abs().max())
    for k, v in ref_s.items():
        for rk in range(W):
            assert torch.allclose(r[rk]["stats"][k], v, rtol=2e-2, atol=1e-3), k
    for k, g in ref_g.items():
        tot = r[0]["grads"][k] + r[1]["grads"][k]
        if float(g.abs().max()) < 1e-4:
            continue
        if k.endswith("_conv.0.bias") or k.endswith("first_conv.3.bias"):
            continue  # shifts a batch-statistics BatchNorm removes again: zero gradient, rounding noise on both sides
        err = float((tot - g).abs().max())
        cos = float(torch.nn.functional.cosine_similarity(tot.flatten(), g.flatten(), dim=0))
        assert err < 8e-2 * float(g.abs().max()) + 1e-4 and cos > 0.998, (k, err, float(g.abs().max()), cos)  # bf16 activations

