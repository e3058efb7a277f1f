"""Tower assembly with the reference's classes / attributes / state_dict keys (reference open_clip/model.py):
CLIPVisionCfg, CLIPTextCfg, CLIP, TriCLIP, _build_vision_tower, _build_text_tower."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from vitlens_b200 import engine as E

from .module_cfg import set_default_image_cfg
from .transformer import LayerNorm, LayerNormFp32, QuickGELU, TextTransformer, VisionTransformer, encode_text_tokens


@dataclass
class CLIPVisionCfg:
    layers: Union[Tuple[int, int, int, int], int] = 12
    width: int = 768
    head_width: int = 64
    mlp_ratio: float = 4.0
    patch_size: int = 16
    image_size: Union[Tuple[int, int], int] = 224
    ls_init_value: Optional[float] = None
    patch_dropout: float = 0.0
    input_patchnorm: bool = False
    global_average_pool: bool = False
    attentional_pool: bool = False
    n_queries: int = 256
    attn_pooler_heads: int = 8
    output_tokens: bool = False
    timm_model_name: str = None
    timm_model_pretrained: bool = False
    timm_pool: str = "avg"
    timm_proj: str = "linear"
    timm_proj_bias: bool = False
    timm_drop: float = 0.0
    timm_drop_path: Optional[float] = None
    # ViT-Lens
    visual_modality_type: str = "image"
    use_perceiver: bool = False
    perceiver_cfg: Optional[dict] = None
    use_visual_adapter: bool = False
    visual_adapter_cfg: Optional[dict] = None
    visual_arch: str = "perceiver_vit"
    exp_args: Optional[Any] = None


@dataclass
class CLIPTextCfg:
    context_length: int = 77
    vocab_size: int = 49408
    width: int = 512
    heads: int = 8
    layers: int = 12
    ls_init_value: Optional[float] = None
    hf_model_name: str = None
    hf_tokenizer_name: str = None
    hf_model_pretrained: bool = True
    proj: str = "mlp"
    pooler_type: str = "mean_pooler"
    embed_cls: bool = False
    pad_id: int = 0
    output_tokens: bool = False


def get_cast_dtype(precision: str):
    return {"bf16": torch.bfloat16, "fp16": torch.float16}.get(precision)


def get_input_dtype(precision: str):
    if precision in ("bf16", "pure_bf16"):
        return torch.bfloat16
    if precision in ("fp16", "pure_fp16"):
        return torch.float16
    return None


def _build_vision_tower(embed_dim: int, vision_cfg: CLIPVisionCfg, quick_gelu: bool = False, cast_dtype: Optional[torch.dtype] = None):
    if isinstance(vision_cfg, dict):
        vision_cfg = CLIPVisionCfg(**vision_cfg)
    if vision_cfg.timm_model_name or isinstance(vision_cfg.layers, (tuple, list)):
        raise NotImplementedError("timm / ModifiedResNet towers are other open_clip model families, outside the ViT-Lens hot path")
    act_layer = QuickGELU if quick_gelu else nn.GELU
    vision_heads = vision_cfg.width // vision_cfg.head_width
    return VisionTransformer(
        image_size=vision_cfg.image_size, patch_size=vision_cfg.patch_size, width=vision_cfg.width, layers=vision_cfg.layers,
        heads=vision_heads, mlp_ratio=vision_cfg.mlp_ratio, ls_init_value=vision_cfg.ls_init_value,
        patch_dropout=vision_cfg.patch_dropout, input_patchnorm=vision_cfg.input_patchnorm,
        global_average_pool=vision_cfg.global_average_pool, attentional_pool=vision_cfg.attentional_pool,
        n_queries=vision_cfg.n_queries, attn_pooler_heads=vision_cfg.attn_pooler_heads, output_tokens=vision_cfg.output_tokens,
        output_dim=embed_dim, act_layer=act_layer, norm_layer=LayerNorm, vision_cfg=vision_cfg)


def _build_text_tower(embed_dim: int, text_cfg: CLIPTextCfg, quick_gelu: bool = False, cast_dtype: Optional[torch.dtype] = None):
    if isinstance(text_cfg, dict):
        text_cfg = CLIPTextCfg(**text_cfg)
    if text_cfg.hf_model_name:
        raise NotImplementedError("HF text encoders are another open_clip model family, outside the ViT-Lens hot path")
    act_layer = QuickGELU if quick_gelu else nn.GELU
    return TextTransformer(
        context_length=text_cfg.context_length, vocab_size=text_cfg.vocab_size, width=text_cfg.width, heads=text_cfg.heads,
        layers=text_cfg.layers, ls_init_value=text_cfg.ls_init_value, output_dim=embed_dim, embed_cls=text_cfg.embed_cls,
        output_tokens=text_cfg.output_tokens, pad_id=text_cfg.pad_id, act_layer=act_layer, norm_layer=LayerNorm)


def _normalize(features):
    return E.L2NormFn.apply(features)


class _TextMixin:
    def _adopt_text(self, text):
        self.transformer = text.transformer
        self.context_length = text.context_length
        self.vocab_size = text.vocab_size
        self.token_embedding = text.token_embedding
        self.positional_embedding = text.positional_embedding
        self.ln_final = text.ln_final
        self.text_projection = text.text_projection
        self.register_buffer("attn_mask", text.attn_mask, persistent=False)

    def lock_text_tower(self, unlocked_layers: int = 0, freeze_layer_norm: bool = True):
        self.transformer.lock(unlocked_layers, freeze_layer_norm)
        for x in (self.token_embedding, self.positional_embedding, self.ln_final, self.text_projection):
            if isinstance(x, torch.nn.Parameter):
                x.requires_grad = False
            else:
                for p in x.parameters():
                    p.requires_grad = False

    def encode_text(self, text, normalize: bool = False):
        x = encode_text_tokens(text, self.token_embedding, self.positional_embedding, self.transformer, self.ln_final,
                               self.text_projection, self.attn_mask)
        return _normalize(x) if normalize else x


class CLIP(nn.Module, _TextMixin):
    """model.py:229-326."""

    def __init__(self, embed_dim: int, vision_cfg: CLIPVisionCfg, text_cfg: CLIPTextCfg, quick_gelu: bool = False,
                 cast_dtype: Optional[torch.dtype] = None, output_dict: bool = False):
        super().__init__()
        self.output_dict = output_dict
        self.visual = _build_vision_tower(embed_dim, vision_cfg, quick_gelu, cast_dtype)
        self._adopt_text(_build_text_tower(embed_dim, text_cfg, quick_gelu, cast_dtype))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def lock_image_tower(self, unlocked_groups=0, freeze_bn_stats=False):
        self.visual.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats)

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.visual.set_grad_checkpointing(enable)
        self.transformer.grad_checkpointing = enable

    def encode_image(self, image, normalize: bool = False):
        features = self.visual(image)
        return _normalize(features) if normalize else features

    def forward(self, image: Optional[torch.Tensor] = None, text: Optional[torch.Tensor] = None):
        image_features = self.encode_image(image, normalize=True) if image is not None else None
        text_features = self.encode_text(text, normalize=True) if text is not None else None
        if self.output_dict:
            return {"image_features": image_features, "text_features": text_features, "logit_scale": self.logit_scale.exp()}
        return image_features, text_features, self.logit_scale.exp()


class TriCLIP(nn.Module, _TextMixin):
    """model.py:391-621: .image (plain CLIP ViT), .visual (adapter + Lens + ViT), text pieces, logit_scale."""

    def __init__(self, embed_dim: int, vision_cfg: CLIPVisionCfg, text_cfg: CLIPTextCfg, quick_gelu: bool = False,
                 cast_dtype: Optional[torch.dtype] = None, output_dict: bool = False):
        super().__init__()
        vision_cfg = CLIPVisionCfg(**vision_cfg) if isinstance(vision_cfg, dict) else vision_cfg
        self.exp_args = vision_cfg.exp_args
        self.output_dict = output_dict
        self.visual_arch = vision_cfg.visual_arch
        if self.visual_arch != "perceiver_vit":
            raise NotImplementedError(f"visual_arch={self.visual_arch!r}: only 'perceiver_vit' is on the covered path")
        self.image = _build_vision_tower(embed_dim, set_default_image_cfg(vision_cfg), quick_gelu, cast_dtype)
        self.visual = _build_vision_tower(embed_dim, vision_cfg, quick_gelu, cast_dtype)
        self._adopt_text(_build_text_tower(embed_dim, text_cfg, quick_gelu, cast_dtype))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))

    def lock_image_tower(self, unlocked_groups=0, freeze_bn_stats=False, unlock_cls=False, unlock_pos_emb=False):
        self.image.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats, unlock_cls=unlock_cls, unlock_pos_emb=unlock_pos_emb)

    def lock_visual_tower(self, unlocked_groups=0, freeze_bn_stats=False, unlock_cls=False, unlock_pos_emb=False, unlock_trans_first_n_layers=None):
        self.visual.lock(unlocked_groups=unlocked_groups, freeze_bn_stats=freeze_bn_stats, unlock_cls=unlock_cls,
                         unlock_pos_emb=unlock_pos_emb, unlock_trans_first_n_layers=unlock_trans_first_n_layers)

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.visual.set_grad_checkpointing(enable)
        self.transformer.grad_checkpointing = enable
        self.image.set_grad_checkpointing(enable)

    def encode_image(self, image, normalize: bool = False):
        n_img = None
        if image.ndim == 5:  # per-frame encode + mean (model.py:511-520)
            n_img = image.size(1)
            image = image.reshape((-1,) + tuple(image.shape[2:]))
        features = self.image(image)
        if n_img is not None:
            features = features.reshape(-1, n_img, features.shape[-1]).mean(1)
        return _normalize(features) if normalize else features

    def encode_visual(self, visual_x, normalize: bool = False, **kwargs):
        features = self.visual(visual_x, **kwargs)
        return _normalize(features) if normalize else features

    def forward(self, image: Optional[torch.Tensor] = None, text: Optional[torch.Tensor] = None, visual_x: Optional[torch.Tensor] = None):
        if self.exp_args is not None and getattr(self.exp_args, "visual_modality_type", None) == "video" and getattr(self.exp_args, "vid_distill_tokens", False):
            raise NotImplementedError("video token distillation is outside the ViT-Lens hot path")
        # 5-D image input [b, t, c, h, w] (model.py:591-604): every frame is encoded AND L2-normalised on its own, the
        # normalised frame features are averaged over t, and the mean is normalised again
        n_img = None
        if image is not None and image.ndim == 5:
            n_img = image.size(1)
            image = image.reshape((-1,) + tuple(image.shape[2:]))
        image_features = self.encode_image(image, normalize=True) if image is not None else None
        if n_img is not None:
            image_features = _normalize(image_features.reshape(-1, n_img, image_features.shape[-1]).mean(1))
        text_features = self.encode_text(text, normalize=True) if text is not None else None
        visual_features = self.encode_visual(visual_x, normalize=True) if visual_x is not None else None
        if self.output_dict:
            return {"image_features": image_features, "text_features": text_features, "visual_features": visual_features,
                    "logit_scale": self.logit_scale.exp()}
        return image_features, text_features, visual_features, self.logit_scale.exp()


def resize_pos_embed(state_dict, model, interpolation: str = "bicubic", antialias: bool = True):
    """model.py:1079-1146: when a checkpoint's `visual.positional_embedding` has another token count than the model, the
    grid part (everything after the class token) is resampled as a 2-D image to the model's grid -- for a Lens tower to
    the floor(sqrt(num_latents))^2 grid and then nearest-neighbour stretched to num_latents rows -- and written back into
    `state_dict`.  Host-side, runs once at load time."""
    import math

    import torch.nn.functional as F

    old = state_dict.get("visual.positional_embedding", None)
    visual = getattr(model, "visual", None)
    if old is None or visual is None or not hasattr(visual, "grid_size"):
        return
    gh, gw = visual.grid_size if isinstance(visual.grid_size, (tuple, list)) else (visual.grid_size, visual.grid_size)
    extra = 1  # the class token
    lens = bool(getattr(visual, "use_perceiver", False))
    n_lat = visual.vision_cfg.exp_args.perceiver_num_latents if lens else None
    new_len = (n_lat if lens else gh * gw) + extra
    if new_len == old.shape[0]:
        return
    tok, img = old[:extra], old[extra:]
    og = int(math.sqrt(len(img)))
    target = (int(math.sqrt(n_lat)),) * 2 if lens else (gh, gw)
    img = img.reshape(1, og, og, -1).permute(0, 3, 1, 2)
    img = F.interpolate(img, size=target, mode=interpolation, antialias=antialias, align_corners=False)
    img = img.permute(0, 2, 3, 1).reshape(target[0] * target[1], -1)
    if lens and target[0] * target[1] != n_lat:
        img = F.interpolate(img.unsqueeze(0).transpose(1, 2), size=n_lat, mode="nearest").transpose(1, 2).squeeze(0)
    state_dict["visual.positional_embedding"] = torch.cat([tok, img], dim=0)


def convert_weights_to_lp(model: nn.Module, dtype=torch.float16):
    """No-op: master weights stay fp32; the kernels always consume cached bf16 copies (model.py:795-827)."""
    return model


def trace_model(model, batch_size=256, device=torch.device("cpu")):
    raise NotImplementedError("torch.jit tracing is not supported: the towers call custom sm_100a kernels")
