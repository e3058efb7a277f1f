"""Pin the DISTRIBUTED loss semantics against the real reference  --  TEST INFRASTRUCTURE (build container only).

    python oracle/make_golden_dist.py

Spawns WORLD processes (gloo, CPU); each imports the unmodified reference from /root/reference and runs its own
`gather_features` + `ClipLoss` / `TriClipLoss` (open_clip/loss.py:20-165,311-385) with world_size = WORLD for the four
(local_loss, gather_with_grad) combinations on seeded L2-normalised feature blocks, then backward.  What each rank ends up
with -- its loss value, d(loss)/d(local features), d(loss)/d(logit_scale) -- is committed as tests/golden/dist_loss_w2.pt.
tests/test_oracle_golden.py pins oracle.clip_loss_sharded to it, tests/test_dist_gloo.py and the multi-GPU checks
(bench.py --verify, tests/test_gpu_multi.py) compare this repo's loss modules with it.

gloo has no all_to_all, which torch.distributed.nn.all_gather's backward uses on non-NCCL backends; the harness supplies
that one collective (built from all_gather) -- the reference's code is not touched.
"""
from __future__ import annotations

import os
import socket
import sys

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

WORLD, BL, E = 2, 6, 32
SCALE_LOG = 2.5
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dist_loss_w2.pt")


def feature_blocks(seed: int, world: int = WORLD, bl: int = BL, e: int = E) -> torch.Tensor:
    """[world, bl, e] unit-norm rows; block r belongs to rank r (the recipe the tests regenerate)."""
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(world, bl, e, generator=g), dim=-1)


def _gloo_all_to_all(output_tensor_list, input_tensor_list, group=None, async_op=False):
    import torch.distributed as dist

    w = dist.get_world_size(group)
    r = dist.get_rank(group)
    for src in range(w):  # rank r receives input_tensor_list[r] of every source rank
        bucket = [torch.empty_like(t) for t in input_tensor_list]
        # every rank contributes the whole list; gather slot `dst` of rank `src`
        for dst in range(w):
            got = [torch.empty_like(input_tensor_list[dst]) for _ in range(w)]
            dist.all_gather(got, input_tensor_list[dst].contiguous(), group=group)
            bucket[dst] = got[src]
        output_tensor_list[src].copy_(bucket[r])


def _worker(rank, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from oracle import ref_import

    open_clip, _, _ = ref_import.import_reference()
    dist.all_to_all = _gloo_all_to_all  # see module docstring
    res = {}
    X, Y, V = feature_blocks(1), feature_blocks(2), feature_blocks(3)
    for tri in (False, True):
        for local_loss in (False, True):
            for gwg in (False, True):
                x = X[rank].clone().requires_grad_(True)
                y = Y[rank].clone().requires_grad_(True)
                v = V[rank].clone().requires_grad_(True)
                s = torch.tensor(SCALE_LOG, requires_grad=True)
                kw = dict(local_loss=local_loss, gather_with_grad=gwg, rank=rank, world_size=WORLD)
                if tri:
                    loss = open_clip.loss.TriClipLoss(**kw)(x, y, v, s.exp())
                else:
                    loss = open_clip.loss.ClipLoss(**kw)(x, y, s.exp())
                loss.backward()
                res[(tri, local_loss, gwg)] = dict(loss=loss.detach().clone(), dx=x.grad.clone(), dy=y.grad.clone(),
                                                   dv=v.grad.clone() if tri else None, ds=s.grad.clone())
    torch.save(res, out.format(rank))
    dist.destroy_process_group()


def main():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    tmp = "/tmp/_dist_golden_r{}.pt"
    mp.spawn(_worker, args=(port, tmp), nprocs=WORLD, join=True)
    per_rank = [torch.load(tmp.format(r), weights_only=False) for r in range(WORLD)]
    fx = {"world": WORLD, "bl": BL, "e": E, "scale_log": SCALE_LOG, "seeds": (1, 2, 3)}
    for key in per_rank[0]:
        tri, ll, gwg = key
        name = f"{'tri' if tri else 'clip'}_local{int(ll)}_gwg{int(gwg)}"
        for r in range(WORLD):
            for k, v in per_rank[r][key].items():
                if v is not None:
                    fx[f"{name}/rank{r}/{k}"] = v
        print(name, [float(per_rank[r][key]["loss"]) for r in range(WORLD)], [float(per_rank[r][key]["ds"]) for r in range(WORLD)])
    torch.save(fx, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
