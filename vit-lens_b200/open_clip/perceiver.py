"""The Lens: a lucidrains-style Perceiver (reference open_clip/perceiver.py:67-332) with the same
parameter tree (latents, layers.{d}.{0,1,2}.…) whose stages run engine.LensAttnFn / LensFFFn."""
from __future__ import annotations

from functools import wraps

import torch
from torch import nn

from vitlens_b200 import engine as E

from . import transformer
from .transformer import TokenMat


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


def cache_fn(f):
    cache = dict()

    @wraps(f)
    def cached_fn(*args, _cache=True, key=None, **kwargs):
        if not _cache:
            return f(*args, **kwargs)
        nonlocal cache
        if key in cache:
            return cache[key]
        result = f(*args, **kwargs)
        cache[key] = result
        return result

    return cached_fn


class PreNorm(nn.Module):
    """perceiver.py:67-82 (parameter container: norm, norm_context, fn)."""

    def __init__(self, dim, fn, context_dim=None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if exists(context_dim) else None


class GEGLU(nn.Module):
    """perceiver.py:85-88 marker (evaluated by vl_geglu_fwd)."""


class FeedForward(nn.Module):
    """perceiver.py:91-102: net.0 = Linear(dim, 8 dim), net.1 = GEGLU, net.2 = Linear(4 dim, dim), net.3 = Dropout."""

    def __init__(self, dim, mult=4, dropout=0.0):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("ff_dropout > 0 is not used by any ViT-Lens config")
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), GEGLU(), nn.Linear(dim * mult, dim), nn.Dropout(dropout))


class Attention(nn.Module):
    """perceiver.py:105-154 (parameter container: to_q, to_kv bias-free; to_out with bias)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("Lens attention head_dim must be 64")
        if dropout != 0.0:
            raise NotImplementedError("attn_dropout > 0 is not used by any ViT-Lens config")
        inner_dim = dim_head * heads
        context_dim = default(context_dim, query_dim)
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(context_dim, inner_dim * 2, bias=False)
        self.dropout = nn.Dropout(dropout)
        self.to_out = nn.Linear(inner_dim, query_dim)


class Perceiver(nn.Module):
    """perceiver.py:157-332 with fourier_encode_data=False (every ViT-Lens config, model_cfg.py:85-182)."""

    def __init__(self, *, num_freq_bands, depth, max_freq, input_channels=3, input_axis=2, num_latents=512, latent_dim=512,
                 cross_heads=1, latent_heads=8, cross_dim_head=64, latent_dim_head=64, num_classes=1000, attn_dropout=0.0,
                 ff_dropout=0.0, weight_tie_layers=False, fourier_encode_data=True, self_per_cross_attn=1,
                 final_classifier_head=True):
        super().__init__()
        if fourier_encode_data:
            raise NotImplementedError("fourier_encode_data=True is not used by any ViT-Lens config")
        if final_classifier_head:
            raise NotImplementedError("the classifier head is never built on the ViT-Lens path (get_perceiver passes False)")
        self.input_axis = input_axis
        self.max_freq = max_freq
        self.num_freq_bands = num_freq_bands
        self.fourier_encode_data = False
        input_dim = input_channels
        self.latents = nn.Parameter(torch.randn(num_latents, latent_dim))

        get_cross_attn = lambda: PreNorm(latent_dim, Attention(latent_dim, input_dim, heads=cross_heads, dim_head=cross_dim_head, dropout=attn_dropout), context_dim=input_dim)
        get_cross_ff = lambda: PreNorm(latent_dim, FeedForward(latent_dim, dropout=ff_dropout))
        get_latent_attn = lambda: PreNorm(latent_dim, Attention(latent_dim, heads=latent_heads, dim_head=latent_dim_head, dropout=attn_dropout))
        get_latent_ff = lambda: PreNorm(latent_dim, FeedForward(latent_dim, dropout=ff_dropout))
        get_cross_attn, get_cross_ff, get_latent_attn, get_latent_ff = map(cache_fn, (get_cross_attn, get_cross_ff, get_latent_attn, get_latent_ff))

        self.layers = nn.ModuleList([])
        for i in range(depth):
            should_cache = i > 0 and weight_tie_layers
            cache_args = {"_cache": should_cache}
            self_attns = nn.ModuleList([])
            for block_ind in range(self_per_cross_attn):
                self_attns.append(nn.ModuleList([get_latent_attn(**cache_args, key=block_ind), get_latent_ff(**cache_args, key=block_ind)]))
            self.layers.append(nn.ModuleList([get_cross_attn(**cache_args), get_cross_ff(**cache_args), self_attns]))
        self.to_logits = nn.Identity()

    @staticmethod
    def _attn(pre: PreNorm, x: TokenMat, data):
        a = pre.fn
        nc = pre.norm_context
        y = E.LensAttnFn.apply(
            x.t, None if data is None else data.t, pre.norm.weight, pre.norm.bias,
            None if nc is None else nc.weight, None if nc is None else nc.bias,
            a.to_q.weight, a.to_kv.weight, a.to_out.weight, a.to_out.bias,
            x.B, x.N, 0 if data is None else data.N, a.heads)
        return TokenMat(y, x.B, x.N)

    @staticmethod
    def _ff(pre: PreNorm, x: TokenMat):
        n = pre.fn.net
        return TokenMat(E.LensFFFn.apply(x.t, pre.norm.weight, pre.norm.bias, n[0].weight, n[0].bias, n[2].weight, n[2].bias), x.B, x.N)

    def forward(self, data, mask=None, return_embeddings=False):
        if mask is not None:
            raise NotImplementedError("masked Lens cross-attention is never used on the ViT-Lens path")
        if not return_embeddings:
            raise NotImplementedError("Perceiver is only used with return_embeddings=True (transformer.py:753)")
        public = not isinstance(data, TokenMat)
        d = TokenMat.from_bnd(data) if public else data
        x = TokenMat(E.BroadcastRowsFn.apply(self.latents, d.B), d.B, self.latents.shape[0])
        for cross_attn, cross_ff, self_attns in self.layers:
            x = self._attn(cross_attn, x, d)
            x = self._ff(cross_ff, x)
            for self_attn, self_ff in self_attns:
                x = self._attn(self_attn, x, None)
                x = self._ff(self_ff, x)
        return x.to_bnd().to(data.dtype) if public else x


class _TokenIdentity(nn.Identity):
    pass


def get_perceiver(cfg, args, **kwargs):
    """perceiver.py:369-401: Identity / Transformer / Perceiver."""
    if args.perceiver_as_identity or (not args.use_perceiver):
        return _TokenIdentity()
    elif args.perceiver_as_transformer:
        return transformer.Transformer(
            kwargs["transformer_width"], cfg["depth"], kwargs["transformer_heads"], kwargs["transformer_mlp_ratio"],
            ls_init_value=kwargs["transformer_ls_init_value"], act_layer=kwargs["transformer_act_layer"],
            norm_layer=kwargs["transformer_norm_layer"])
    return Perceiver(
        input_channels=cfg["input_chan"], input_axis=cfg["input_axis"], num_freq_bands=cfg["num_freq_bands"],
        max_freq=cfg["max_freq"], depth=cfg["depth"], num_latents=cfg["num_latents"], latent_dim=cfg["latent_dim"],
        cross_heads=cfg["cross_heads"], latent_heads=cfg["latent_heads"], cross_dim_head=cfg["cross_dim_head"],
        latent_dim_head=cfg["latent_dim_head"], num_classes=cfg["num_classes"], attn_dropout=cfg["attn_dropout"],
        ff_dropout=cfg["ff_dropout"], weight_tie_layers=cfg["weight_tie_layers"],
        fourier_encode_data=cfg["fourier_encode_data"], self_per_cross_attn=cfg["self_per_cross_attn"],
        final_classifier_head=False)
