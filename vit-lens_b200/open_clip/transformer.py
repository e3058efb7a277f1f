"""ViT / text Transformer modules with the reference's constructor signatures, attribute names
and state_dict keys (reference open_clip/transformer.py), whose forward passes run the
hand-written sm_100a kernels through vitlens_b200.engine.

Layout: the reference keeps activations as [N, B, D] ("LND") fp32/fp16 tensors.  Here the
residual stream between modules is one contiguous bf16 token matrix [B*N, D] (`TokenMat`);
public entry points still accept / return the reference's tensors.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Optional, Sequence, Tuple

import torch
from torch import nn

from vitlens_b200 import engine as E

from .module_cfg import AttrDict


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class TokenMat:
    """[B*N, D] bf16 token matrix + its (B, N) factorisation."""

    __slots__ = ("t", "B", "N")

    def __init__(self, t: torch.Tensor, B: int, N: int):
        assert t.dim() == 2 and t.shape[0] == B * N
        self.t, self.B, self.N = t, B, N

    @property
    def D(self):
        return self.t.shape[1]

    @staticmethod
    def from_bnd(x: torch.Tensor) -> "TokenMat":
        B, N, D = x.shape
        return TokenMat(x.reshape(B * N, D).to(torch.bfloat16).contiguous(), B, N)

    def to_bnd(self) -> torch.Tensor:
        return self.t.reshape(self.B, self.N, self.D)


class LayerNorm(nn.LayerNorm):
    """transformer.py:28-34.  fp32 statistics always (also covers LayerNormFp32, :17-25)."""

    def forward(self, x):
        if isinstance(x, TokenMat):
            return TokenMat(E.LayerNormFn.apply(x.t, self.weight, self.bias, None), x.B, x.N)
        shp = x.shape
        y = E.LayerNormFn.apply(x.reshape(-1, shp[-1]).to(torch.bfloat16).contiguous(), self.weight, self.bias, None)
        return y.reshape(shp).to(x.dtype)


LayerNormFp32 = LayerNorm


class QuickGELU(nn.Module):
    """Marker module (x * sigmoid(1.702 x), transformer.py:37-40); evaluated inside the GEMM epilogue."""

    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


def _is_causal_mask(attn_mask) -> bool:
    if attn_mask is None:
        return False
    m = attn_mask
    if m.dim() != 2 or m.shape[0] != m.shape[1]:
        raise NotImplementedError("only the causal additive attn_mask of the CLIP text tower is supported")
    return True


class ResidualAttentionBlock(nn.Module):
    """transformer.py:201-272.  nn.MultiheadAttention / nn.Sequential are kept as *parameter
    containers* so state_dict keys (attn.in_proj_weight, attn.out_proj.*, mlp.c_fc.*, mlp.c_proj.*)
    and initialisation match the reference; the arithmetic is engine.VitBlockFn."""

    def __init__(self, d_model: int, n_head: int, mlp_ratio: float = 4.0, ls_init_value: float = None,
                 act_layer: Callable = nn.GELU, norm_layer: Callable = LayerNorm, is_cross_attention: bool = False):
        super().__init__()
        if ls_init_value is not None:
            raise NotImplementedError("LayerScale (ls_init_value) is unused by every shipped ViT-Lens config")
        if is_cross_attention:
            raise NotImplementedError("cross-attention residual blocks belong to CoCa, outside the ViT-Lens hot path")
        if d_model % n_head or d_model // n_head != 64:
            raise NotImplementedError(f"head_dim must be 64 (got d_model={d_model}, heads={n_head})")
        self.n_head = n_head
        self.ln_1 = norm_layer(d_model)
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ls_1 = nn.Identity()
        self.ln_2 = norm_layer(d_model)
        mlp_width = int(d_model * mlp_ratio)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, mlp_width)),
            ("gelu", act_layer()),
            ("c_proj", nn.Linear(mlp_width, d_model)),
        ]))
        self.ls_2 = nn.Identity()
        self.quick = isinstance(self.mlp.gelu, QuickGELU)

    def forward(self, q_x, k_x=None, v_x=None, attn_mask=None):
        assert k_x is None and v_x is None
        public = not isinstance(q_x, TokenMat)
        x = TokenMat.from_bnd(q_x.permute(1, 0, 2)) if public else q_x  # public entry is LND like the reference
        y = E.VitBlockFn.apply(
            x.t, self.ln_1.weight, self.ln_1.bias, self.attn.in_proj_weight, self.attn.in_proj_bias,
            self.attn.out_proj.weight, self.attn.out_proj.bias, self.ln_2.weight, self.ln_2.bias,
            self.mlp.c_fc.weight, self.mlp.c_fc.bias, self.mlp.c_proj.weight, self.mlp.c_proj.bias,
            x.B, x.N, self.n_head, _is_causal_mask(attn_mask), self.quick)
        out = TokenMat(y, x.B, x.N)
        return out.to_bnd().permute(1, 0, 2).to(q_x.dtype) if public else out


class Transformer(nn.Module):
    """transformer.py:329-375."""

    def __init__(self, width: int, layers: int, heads: int, mlp_ratio: float = 4.0, ls_init_value: float = None,
                 act_layer: Callable = nn.GELU, norm_layer: Callable = LayerNorm):
        super().__init__()
        self.width = width
        self.layers = layers
        self.grad_checkpointing = False
        self.resblocks = nn.ModuleList([
            ResidualAttentionBlock(width, heads, mlp_ratio, ls_init_value=ls_init_value, act_layer=act_layer, norm_layer=norm_layer)
            for _ in range(layers)])

    def get_cast_dtype(self) -> torch.dtype:
        return self.resblocks[0].mlp.c_fc.weight.dtype

    def forward(self, x, attn_mask=None):
        public = not isinstance(x, TokenMat)
        t = TokenMat.from_bnd(x.permute(1, 0, 2)) if public else x
        for r in self.resblocks:
            if self.grad_checkpointing and torch.is_grad_enabled() and (t.t.requires_grad or any(p.requires_grad for p in r.parameters())):
                # transformer.py:366-368: `x = checkpoint(r, x, None, None, attn_mask)` -- keep only the block's input (one
                # [T, D] bf16 matrix instead of ~17 T D bytes of saved activations) and re-run its kernels inside backward
                B, N = t.B, t.N
                y = torch.utils.checkpoint.checkpoint(lambda tt, r=r, B=B, N=N: r(TokenMat(tt, B, N), attn_mask=attn_mask).t, t.t,
                                                      use_reentrant=False)
                t = TokenMat(y, B, N)
            else:
                t = r(t, attn_mask=attn_mask)
        return t.to_bnd().permute(1, 0, 2).to(x.dtype) if public else t

    def lock(self, *args, **kwargs):
        for p in self.parameters():
            p.requires_grad = False


class VisionTransformer(nn.Module):
    """transformer.py:378-792: visual adapter -> Lens (Perceiver) -> cls + positional -> ln_pre ->
    ViT blocks -> ln_post(cls) @ proj.  Returns fp32 [B, output_dim]."""

    def __init__(self, image_size: int, patch_size: int, width: int, layers: int, heads: int, mlp_ratio: float,
                 ls_init_value: float = None, global_average_pool: bool = False, attentional_pool: bool = False,
                 n_queries: int = 256, attn_pooler_heads: int = 8, output_dim: int = 512, patch_dropout: float = 0.0,
                 input_patchnorm: bool = False, act_layer: Callable = nn.GELU, norm_layer: Callable = LayerNorm,
                 output_tokens: bool = False, vision_cfg=None):
        super().__init__()
        if global_average_pool or attentional_pool or input_patchnorm or patch_dropout > 0.0:
            raise NotImplementedError("global_average_pool / attentional_pool / input_patchnorm / patch_dropout are outside the ViT-Lens hot path")
        self.output_tokens = output_tokens
        image_height, image_width = self.image_size = to_2tuple(image_size)
        patch_height, patch_width = self.patch_size = to_2tuple(patch_size)
        self.grid_size = (image_height // patch_height, image_width // patch_width)
        self.width = width
        self.output_dim = output_dim
        if vision_cfg is None:
            vision_cfg = AttrDict(perceiver_cfg=None, visual_adapter_cfg=None, visual_modality_type="image", exp_args=None)
        self.vision_cfg = vision_cfg

        self.perceiver = None
        perceiver_cfg = vision_cfg.perceiver_cfg
        self.use_perceiver = perceiver_cfg.use_perceiver if perceiver_cfg else False
        if self.use_perceiver:
            from .perceiver import get_perceiver

            self.perceiver = get_perceiver(
                perceiver_cfg, args=self.vision_cfg.exp_args, transformer_width=width, transformer_heads=heads,
                transformer_mlp_ratio=mlp_ratio, transformer_ls_init_value=ls_init_value,
                transformer_act_layer=act_layer, transformer_norm_layer=norm_layer)

        self.visual_adapter = None
        self.patchnorm_pre_ln, self.conv1 = None, None
        visual_adapter_cfg = vision_cfg.visual_adapter_cfg
        self.use_visual_adapter = visual_adapter_cfg.use_visual_adapter if visual_adapter_cfg else False
        if self.use_visual_adapter:
            from .visual_adapter import get_visual_adapter

            self.visual_adapter = get_visual_adapter(
                visual_adapter_cfg, grid_size=self.grid_size, patch_size=self.patch_size, width=width,
                input_patchnorm=input_patchnorm, exp_args=self.vision_cfg.exp_args)

        if self.vision_cfg.visual_modality_type in ("image", "video", "tactile"):
            if self.vision_cfg.visual_modality_type == "video":
                raise NotImplementedError("video towers: the reference's own video path is broken (transformer.py:478, visual_adapter.py:65-66)")
            self.input_patchnorm = False
            self.patchnorm_pre_ln = nn.Identity()
            self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)

        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        if self.use_perceiver:
            n_lat = self.vision_cfg.exp_args.perceiver_num_latents
            self.positional_embedding = nn.Parameter(scale * torch.randn(n_lat + 1, width))
        else:
            self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid_size[0] * self.grid_size[1] + 1, width))

        self.patch_dropout = nn.Identity()
        self.ln_pre = norm_layer(width)
        self.transformer = Transformer(width, layers, heads, mlp_ratio, ls_init_value=ls_init_value, act_layer=act_layer, norm_layer=norm_layer)
        self.global_average_pool = False
        self.attn_pool = None
        self.ln_post = norm_layer(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

        self.use_orig_pos = True
        if visual_adapter_cfg and visual_adapter_cfg.disable_orig_pos:
            self.use_orig_pos = False

    # ------------------------------------------------------------------ freezing (transformer.py:553-627)
    def lock(self, unlocked_groups=0, freeze_bn_stats=False, unlock_cls=False, unlock_pos_emb=False, unlock_trans_first_n_layers=None):
        for param in self.parameters():
            param.requires_grad = False

        def _unlock(x):
            if x is None:
                return
            if isinstance(x, Sequence) and not isinstance(x, nn.Module):
                for g in x:
                    _unlock(g)
            elif isinstance(x, torch.nn.Parameter):
                x.requires_grad = True
            else:
                for p in x.parameters():
                    p.requires_grad = True

        exp = self.vision_cfg.exp_args
        if unlocked_groups != 0:
            groups = [
                [self.conv1, self.class_embedding, self.positional_embedding, self.ln_pre],
                *self.transformer.resblocks[:-1],
                [self.transformer.resblocks[-1], self.ln_post],
                self.proj,
            ]
            if exp is not None and getattr(exp, "unlock_from_head", False):
                _unlock(groups[:unlocked_groups])
            else:
                _unlock(groups[-unlocked_groups:])
        groups = []
        if self.perceiver is not None:
            groups.append(self.perceiver)
        if self.visual_adapter is not None:
            groups.append(self.visual_adapter)
        if unlock_cls:
            groups.append(self.class_embedding)
        if unlock_pos_emb:
            groups.append(self.positional_embedding)
        if unlock_trans_first_n_layers is not None:
            for i in range(unlock_trans_first_n_layers):
                groups.append(self.transformer.resblocks[i])
        _unlock(groups)

    def init_parameters(self):
        pass

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.transformer.grad_checkpointing = enable  # Transformer.forward re-runs each block's kernels in backward

    # ------------------------------------------------------------------ forward pieces
    def img_adapter_forawrd(self, x: torch.Tensor) -> TokenMat:  # (sic) name kept from transformer.py:659
        B, C, H, W = x.shape
        kh, kw = self.patch_size
        gh, gw = H // kh, W // kw
        geom = dict(B=B, C=C, OH=gh, OW=gw, kh=kh, kw=kw, stride_h=kh, stride_w=kw,
                    sb=x.stride(0), sc=x.stride(1), sh=x.stride(2), sw=x.stride(3))
        x = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        tok = E.PatchEmbedFn.apply(x, self.conv1.weight, geom)
        return TokenMat(tok, B, gh * gw)

    def _adapter_tokens(self, x, **kwargs) -> TokenMat:
        mt = self.vision_cfg.visual_modality_type
        if mt in ("image", "tactile"):
            return self.img_adapter_forawrd(x)
        assert self.visual_adapter is not None, "Please use visual adapter for this modality type."
        assert self.use_perceiver, f"Other modalities shall use perceiver, got {mt} did not set perceiver configuration."
        x_vada = self.visual_adapter(x, **kwargs)
        if isinstance(x_vada, TokenMat):
            return x_vada
        if isinstance(x_vada, torch.Tensor):
            return TokenMat.from_bnd(x_vada)
        tok = x_vada["x"]
        tok = tok if isinstance(tok, TokenMat) else TokenMat.from_bnd(tok)
        if "pos" in x_vada and x_vada["pos"] is not None:
            pos = x_vada["pos"]
            if self.vision_cfg.exp_args.disable_visual_adapter_pos:
                return tok
            if isinstance(pos, TokenMat):  # per-sample positional tokens (point clouds)
                return TokenMat(E.AddFn.apply(tok.t, pos.t), tok.B, tok.N)
            return TokenMat(E.AssembleFn.apply(tok.t, None, pos, tok.B, tok.N), tok.B, tok.N)
        return tok

    def forward(self, x: torch.Tensor, fwd_output_tokens: bool = False, **kwargs):
        t = self._adapter_tokens(x, **kwargs)
        if self.use_perceiver:
            exp = self.vision_cfg.exp_args
            if exp.perceiver_as_identity:
                t = self.perceiver(t)
            elif exp.perceiver_as_transformer:
                # The reference hands its BATCH-FIRST tokens [B, M, C] to transformer.Transformer (transformer.py:751), whose
                # nn.MultiheadAttention is sequence-first: attention runs over dim 0, i.e. across the samples of the batch,
                # separately per token position.  Same arithmetic here: M "samples" of B "tokens" each.
                sw = t.to_bnd().transpose(0, 1).contiguous()
                sw = self.perceiver(TokenMat(sw.reshape(t.N * t.B, t.D), t.N, t.B))
                t = TokenMat(sw.to_bnd().transpose(0, 1).contiguous().reshape(t.B * t.N, t.D), t.B, t.N)
            else:
                t = self.perceiver(t, return_embeddings=True)
        B, L = t.B, t.N
        pos = self.positional_embedding if self.use_orig_pos else None
        x0 = TokenMat(E.AssembleFn.apply(t.t, self.class_embedding, pos, B, L), B, L + 1)
        xt = self.ln_pre(x0)
        xt = self.transformer(xt)
        rows = torch.arange(B, device=xt.t.device, dtype=torch.long) * xt.N
        pooled = E.LayerNormFn.apply(xt.t, self.ln_post.weight, self.ln_post.bias, rows)
        feats = E.ProjFn.apply(pooled, self.proj)
        if self.output_tokens or fwd_output_tokens:
            return feats, xt.to_bnd()[:, 1:]
        return feats


class TextTransformer(nn.Module):
    """transformer.py:795-930 (no embed_cls variant): used as the parameter factory for CLIP / TriCLIP."""

    def __init__(self, context_length: int = 77, vocab_size: int = 49408, width: int = 512, heads: int = 8, layers: int = 12,
                 ls_init_value: float = None, output_dim: int = 512, act_layer: Callable = nn.GELU, norm_layer: Callable = LayerNorm,
                 embed_cls: bool = False, pad_id: int = 0, output_tokens: bool = False):
        super().__init__()
        if embed_cls:
            raise NotImplementedError("embed_cls text towers belong to CoCa, outside the ViT-Lens hot path")
        self.output_tokens = output_tokens
        self.num_pos = self.context_length = context_length
        self.vocab_size = vocab_size
        self.width = width
        self.output_dim = output_dim
        self.heads = heads
        self.pad_id = pad_id
        self.text_projection = nn.Parameter(torch.empty(width, output_dim))
        self.cls_emb = None
        self.token_embedding = nn.Embedding(vocab_size, width)
        self.positional_embedding = nn.Parameter(torch.empty(self.num_pos, width))
        self.transformer = Transformer(width=width, layers=layers, heads=heads, ls_init_value=ls_init_value, act_layer=act_layer, norm_layer=norm_layer)
        self.ln_final = norm_layer(width)
        self.register_buffer("attn_mask", self.build_attention_mask(), persistent=False)
        self.init_parameters()

    def init_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.num_pos, self.num_pos)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    def lock(self, *args, **kwargs):
        for p in self.parameters():
            p.requires_grad = False

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.transformer.grad_checkpointing = enable

    def forward(self, text, fwd_output_tokens=False):
        return encode_text_tokens(text, self.token_embedding, self.positional_embedding, self.transformer, self.ln_final,
                                  self.text_projection, self.attn_mask)


def encode_text_tokens(text, token_embedding, positional_embedding, transformer, ln_final, text_projection, attn_mask):
    """The shared body of CLIP.encode_text / TriCLIP.encode_text / TextTransformer.forward
    (model.py:297-307, 528-540): embedding + pos -> causal blocks -> ln_final on the EOT row -> projection."""
    B, ctx = text.shape
    x = E.TextEmbedFn.apply(text, token_embedding.weight, positional_embedding[:ctx])
    t = transformer(TokenMat(x, B, ctx), attn_mask=attn_mask)
    rows = torch.arange(B, device=text.device, dtype=torch.long) * ctx + text.argmax(dim=-1)
    pooled = E.LayerNormFn.apply(t.t, ln_final.weight, ln_final.bias, rows)
    return E.ProjFn.apply(pooled, text_projection)
