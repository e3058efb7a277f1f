"""Deterministic synthetic weights and inputs (no checkpoints / datasets are reachable offline).

Every tensor is generated from ``(seed, key name)`` alone, independent of module
construction order, so the reference (in the build container), the CPU oracle
and this package can all materialise *identical* weights for a given
``state_dict`` schema without shipping them.  Input recipes follow SURVEY.md 8(d).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    return g


def synth_tensor(name: str, like: torch.Tensor, seed: int = 0) -> torch.Tensor:
    shape = tuple(like.shape)
    if name.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.long)
    g = _gen(seed, name)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "logit_scale":
        return torch.full(shape, math.log(1 / 0.07))
    if leaf == "running_var":
        return 1.0 + 0.1 * r.abs()
    if leaf == "running_mean":
        return 0.1 * r
    if like.ndim == 1:
        if leaf == "weight":  # LayerNorm / BatchNorm scale
            return 1.0 + 0.1 * r
        if leaf == "bias":
            return 0.02 * r
        return r * shape[0] ** -0.5  # class_embedding
    if leaf in ("positional_embedding", "pos_emb") or name.endswith("token_embedding.weight"):
        return 0.02 * r
    if leaf == "latents":
        return r
    if leaf in ("proj", "text_projection"):  # [in, out]
        return r * shape[0] ** -0.5
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return r * fan_in ** -0.5


def synth_state_dict(like: Mapping[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """A full state_dict with the schema (keys, shapes) of ``like``."""
    return {k: synth_tensor(k, v, seed) for k, v in like.items()}


def synth_text(batch: int, context: int = 77, vocab: int = 49408, seed: int = 1) -> torch.Tensor:
    """Token ids per SURVEY.md 8(d).1: SOT, random body, EOT (= vocab-1, the arg-max), zero padding."""
    g = _gen(seed, "text")
    sot, eot = vocab - 2, vocab - 1
    lo, hi = min(1000, vocab // 8), min(40000, vocab - 2)
    t = torch.zeros(batch, context, dtype=torch.long)
    lens = torch.randint(min(5, context - 3), context - 1, (batch,), generator=g)
    for i in range(batch):
        n = int(lens[i])
        t[i, 0] = sot
        t[i, 1 : n + 1] = torch.randint(lo, hi, (n,), generator=g)
        t[i, n + 1] = eot
    return t


def synth_normal(name: str, shape, seed: int = 1) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=_gen(seed, name), dtype=torch.float32)


def synth_points(batch: int, npoints: int, seed: int = 1):
    """Points uniform in the unit ball + FPS start indices (the reference draws them with
    torch.randint at misc.py:60; they are an explicit input here)."""
    g = _gen(seed, "pc")
    d = torch.randn(batch, npoints, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    rad = torch.rand(batch, npoints, 1, generator=g) ** (1.0 / 3.0)
    start = torch.randint(0, npoints, (batch,), generator=g)
    return d * rad, start


def synth_shapes(batch: int, npoints: int, seed: int = 1):
    """Point clouds of DISTINCT shapes (solid ball, sphere shell, cube surface, cylinder, torus, two blobs, disc, ellipsoid
    rod; cycled over the batch), each randomly stretched and rotated, then centred and scaled into the unit sphere the way the
    reference's loader does (modal_3d/processors/pc_processor.py:32-38, pc_norm).  Uniform balls (synth_points) all look alike
    to the tokenizer, which leaves a contrastive batch with near-identical visual features and ill-conditioned gradients;
    these do not.  Returns (points [B, N, 3], FPS start indices [B])."""
    g = _gen(seed, "pc_shapes")
    out = []
    for i in range(batch):
        u = torch.rand(npoints, 3, generator=g)
        n = torch.randn(npoints, 3, generator=g)
        unit = n / n.norm(dim=-1, keepdim=True).clamp_min(1e-6)
        kind = i % 8
        if kind == 0:      # solid ball
            p = unit * u[:, :1] ** (1.0 / 3.0)
        elif kind == 1:    # thin sphere shell
            p = unit * (1.0 + 0.02 * n[:, :1])
        elif kind == 2:    # cube surface: one coordinate pushed to a face
            p = 2 * u - 1
            face = torch.randint(0, 3, (npoints,), generator=g)
            sign = torch.where(torch.rand(npoints, generator=g) < 0.5, -1.0, 1.0)
            p[torch.arange(npoints), face] = sign
        elif kind == 3:    # cylinder surface
            a = 2 * math.pi * u[:, 0]
            p = torch.stack([torch.cos(a), torch.sin(a), 3 * (u[:, 1] - 0.5)], -1)
        elif kind == 4:    # torus
            a, b = 2 * math.pi * u[:, 0], 2 * math.pi * u[:, 1]
            p = torch.stack([(1 + 0.3 * torch.cos(b)) * torch.cos(a), (1 + 0.3 * torch.cos(b)) * torch.sin(a), 0.3 * torch.sin(b)], -1)
        elif kind == 5:    # two gaussian blobs
            p = 0.25 * n + torch.where(u[:, :1] < 0.5, -1.0, 1.0) * torch.tensor([0.8, 0.3, 0.0])
        elif kind == 6:    # flat disc
            a = 2 * math.pi * u[:, 0]
            p = torch.stack([torch.sqrt(u[:, 1]) * torch.cos(a), torch.sqrt(u[:, 1]) * torch.sin(a), 0.03 * n[:, 2]], -1)
        else:              # long ellipsoid
            p = unit * u[:, :1] ** (1.0 / 3.0) * torch.tensor([1.0, 0.25, 0.15])
        stretch = 0.6 + 0.8 * torch.rand(3, generator=g)
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        p = (p * stretch) @ q.t()
        p = p - p.mean(0, keepdim=True)
        p = p / p.norm(dim=-1).max().clamp_min(1e-6)
        out.append(p)
    start = torch.randint(0, npoints, (batch,), generator=g)
    return torch.stack(out).contiguous(), start
