"""Public constructors with the reference's signatures (reference open_clip/factory.py:55-116, 119-160,
164-365, 368-422, 468-705, 750-851): create_model(_and_transforms), tri_create_model(_and_transforms),
create_loss, get_tokenizer, load_checkpoint, model-config registry.

Scope notes: no pretrained download (offline); `precision` keeps its meaning for the *master* weights
(always fp32 here) while arithmetic runs in bf16 on the tensor cores with fp32 accumulation/statistics;
`device` must be a CUDA device for forward passes (construction / state_dict I/O work anywhere).
"""
from __future__ import annotations

import json
import logging
import os
import re
from copy import deepcopy
from pathlib import Path
from typing import Optional, Tuple, Union

import torch

from .constants import CKPT_CACHE_DIR, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD
from .loss import ClipLoss, ClipLossGeneral, ClipLossLabelMask, ClipLossSimMask, TriClipLoss, TriClipLossLabelMask
from .model import CLIP, TriCLIP, get_cast_dtype, resize_pos_embed
from .module_cfg import get_input_adapter_cfg, get_perceiver_cfg

HF_HUB_PREFIX = "hf-hub:"
_MODEL_CONFIG_PATHS = [Path(__file__).parent / "model_configs/"]
_MODEL_CONFIGS = {}


def _natural_key(string_):
    return [int(s) if s.isdigit() else s for s in re.split(r"(\d+)", string_.lower())]


def _rescan_model_configs():
    global _MODEL_CONFIGS
    files = []
    for config_path in _MODEL_CONFIG_PATHS:
        if config_path.is_file() and config_path.suffix == ".json":
            files.append(config_path)
        elif config_path.is_dir():
            files.extend(config_path.glob("*.json"))
    for cf in files:
        with open(cf, "r") as f:
            model_cfg = json.load(f)
            if all(a in model_cfg for a in ("embed_dim", "vision_cfg", "text_cfg")):
                _MODEL_CONFIGS[cf.stem] = model_cfg
    _MODEL_CONFIGS = {k: v for k, v in sorted(_MODEL_CONFIGS.items(), key=lambda x: _natural_key(x[0]))}


_rescan_model_configs()


def list_models():
    return list(_MODEL_CONFIGS.keys())


def add_model_config(path):
    if not isinstance(path, Path):
        path = Path(path)
    _MODEL_CONFIG_PATHS.append(path)
    _rescan_model_configs()


def get_model_config(model_name):
    return deepcopy(_MODEL_CONFIGS[model_name]) if model_name in _MODEL_CONFIGS else None


def get_tokenizer(model_name):
    from .tokenizer import tokenize

    return tokenize


def load_state_dict(checkpoint_path: str, map_location="cpu"):
    checkpoint = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    state_dict = checkpoint["state_dict"] if isinstance(checkpoint, dict) and "state_dict" in checkpoint else checkpoint
    if next(iter(state_dict.items()))[0].startswith("module"):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    return state_dict


def load_checkpoint(model, checkpoint_path, strict=True, args=None):
    """factory.py:130-160 incl. the `visual.` -> `image.` key duplication for tri-models."""
    state_dict = load_state_dict(checkpoint_path)
    do_pop = args is not None and (args.visual_arch != "perceiver_vit" or args.disable_pt_vit)
    if hasattr(model, "image") and hasattr(model, "visual"):
        orig_sd = dict(state_dict)
        for k, v in orig_sd.items():
            if "visual." in k:
                state_dict[k.replace("visual.", "image.")] = v
                if do_pop:
                    state_dict.pop(k)
    resize_pos_embed(state_dict, model)
    incompatible_keys = model.load_state_dict(state_dict, strict=strict)
    if len(incompatible_keys.missing_keys) or len(incompatible_keys.unexpected_keys):
        logging.info(msg=incompatible_keys)
    return incompatible_keys


def _resolve_cfg(model_name, force_quick_gelu, force_patch_dropout, force_image_size, pretrained_image):
    if model_name.startswith(HF_HUB_PREFIX):
        raise NotImplementedError("hf-hub models need network access")
    model_name = model_name.replace("/", "-")
    model_cfg = get_model_config(model_name)
    if model_cfg is None:
        raise RuntimeError(f"Model config for {model_name} not found; available models {list_models()}.")
    if force_quick_gelu:
        model_cfg["quick_gelu"] = True
    if force_patch_dropout is not None:
        model_cfg["vision_cfg"]["patch_dropout"] = force_patch_dropout
    if force_image_size is not None:
        model_cfg["vision_cfg"]["image_size"] = force_image_size
    if pretrained_image:
        raise NotImplementedError("pretrained image towers are only supported for timm models in the reference")
    return model_name, model_cfg


def _finish(model, model_name, pretrained, precision, device, strict, args, output_dict, require_pretrained):
    if isinstance(device, str):
        device = torch.device(device)
    if precision in ("pure_fp16", "pure_bf16", "fp16"):
        logging.warning("precision=%s: master weights stay fp32; arithmetic runs in bf16 on the tensor cores", precision)
    model.to(device=device)
    pretrained_loaded = False
    if pretrained:
        if os.path.exists(pretrained):
            load_checkpoint(model, pretrained, strict, args)
            pretrained_loaded = True
        else:
            raise RuntimeError(f"Pretrained weights ({pretrained}) not found for model {model_name} (no network access: pass a local path).")
    if require_pretrained and not pretrained_loaded:
        raise RuntimeError(f"Pretrained weights were required for (model: {model_name}, pretrained: {pretrained}) but not loaded.")
    for tower in (getattr(model, "visual", None), getattr(model, "image", None)):
        if tower is not None:
            tower.image_mean = OPENAI_DATASET_MEAN
            tower.image_std = OPENAI_DATASET_STD
    if output_dict and hasattr(model, "output_dict"):
        model.output_dict = True
    return model


def create_model(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32", device: Union[str, torch.device] = "cpu",
                 jit: bool = False, force_quick_gelu: bool = False, force_custom_text: bool = False,
                 force_patch_dropout: Optional[float] = None, force_image_size: Optional[Union[int, Tuple[int, int]]] = None,
                 pretrained_image: bool = False, pretrained_hf: bool = True, cache_dir: Optional[str] = CKPT_CACHE_DIR,
                 output_dict: Optional[bool] = None, require_pretrained: bool = False, strict: bool = False, args=None):
    """factory.py:468-648 (native CLIP towers only)."""
    if jit:
        raise NotImplementedError("torch.jit is not supported: the towers call custom sm_100a kernels")
    model_name, model_cfg = _resolve_cfg(model_name, force_quick_gelu, force_patch_dropout, force_image_size, pretrained_image)
    if model_cfg.pop("custom_text", False) or force_custom_text or "hf_model_name" in model_cfg.get("text_cfg", {}):
        raise NotImplementedError("CustomTextCLIP / HF text towers are outside the ViT-Lens hot path")
    model = CLIP(**model_cfg, cast_dtype=get_cast_dtype(precision))
    return _finish(model, model_name, pretrained, precision, device, strict, args, output_dict, require_pretrained)


def tri_create_model(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32", device: Union[str, torch.device] = "cpu",
                     jit: bool = False, force_quick_gelu: bool = False, force_custom_text: bool = False,
                     force_patch_dropout: Optional[float] = None, force_image_size: Optional[Union[int, Tuple[int, int]]] = None,
                     pretrained_image: bool = False, pretrained_hf: bool = True, cache_dir: Optional[str] = CKPT_CACHE_DIR,
                     output_dict: Optional[bool] = None, require_pretrained: bool = False, strict: bool = False, args=None):
    """factory.py:164-365: TriCLIP = image tower + Lens tower + text tower; `args` carries the Lens / adapter config."""
    if jit:
        raise NotImplementedError("torch.jit is not supported: the towers call custom sm_100a kernels")
    if args is None:
        raise ValueError("tri_create_model needs `args` (the reference reads args.skip_trans_first_n_layers unconditionally, factory.py:348)")
    model_name, model_cfg = _resolve_cfg(model_name, force_quick_gelu, force_patch_dropout, force_image_size, pretrained_image)
    vc = model_cfg["vision_cfg"]
    vc["use_perceiver"] = args.use_perceiver
    vc["visual_modality_type"] = args.visual_modality_type
    vc["perceiver_cfg"] = get_perceiver_cfg(args)
    vc["visual_adapter_cfg"] = get_input_adapter_cfg(args)
    vc["visual_arch"] = args.visual_arch
    vc["exp_args"] = args
    if model_cfg.pop("custom_text", False) or force_custom_text or "hf_model_name" in model_cfg.get("text_cfg", {}):
        raise NotImplementedError("TriCustomTextCLIP / HF text towers are outside the ViT-Lens hot path")
    model = TriCLIP(**model_cfg, cast_dtype=get_cast_dtype(precision))
    # `pretrained` goes through unchanged: a path that does not exist raises (factory.py:318-337 does the same for unknown
    # tags) instead of silently leaving a randomly initialised frozen ViT under the Lens
    model = _finish(model, model_name, pretrained, precision, device, strict, args, output_dict, require_pretrained)
    skip = getattr(args, "skip_trans_first_n_layers", None)
    if skip is not None:
        n_blocks = len(model.visual.transformer.resblocks)
        assert skip < n_blocks
        model.visual.transformer.resblocks = model.visual.transformer.resblocks[-(n_blocks - skip):]
    return model


def _transforms(model, image_mean=None, image_std=None, aug_cfg=None):
    from .transform import image_transform

    tower = getattr(model, "visual", None)
    image_mean = image_mean or getattr(tower, "image_mean", None)
    image_std = image_std or getattr(tower, "image_std", None)
    size = tower.image_size
    return (image_transform(size, is_train=True, mean=image_mean, std=image_std, aug_cfg=aug_cfg),
            image_transform(size, is_train=False, mean=image_mean, std=image_std))


def create_model_and_transforms(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32",
                                device: Union[str, torch.device] = "cpu", jit: bool = False, force_quick_gelu: bool = False,
                                force_custom_text: bool = False, force_patch_dropout: Optional[float] = None,
                                force_image_size=None, pretrained_image: bool = False, pretrained_hf: bool = True,
                                image_mean=None, image_std=None, aug_cfg=None, cache_dir: Optional[str] = CKPT_CACHE_DIR,
                                output_dict: Optional[bool] = None, strict: bool = False, args=None):
    model = create_model(model_name, pretrained, precision=precision, device=device, jit=jit, force_quick_gelu=force_quick_gelu,
                         force_custom_text=force_custom_text, force_patch_dropout=force_patch_dropout, force_image_size=force_image_size,
                         pretrained_image=pretrained_image, pretrained_hf=pretrained_hf, cache_dir=cache_dir, output_dict=output_dict,
                         strict=strict, args=args)
    return (model, *_transforms(model, image_mean, image_std, aug_cfg))


def tri_create_model_and_transforms(model_name: str, pretrained: Optional[str] = None, precision: str = "fp32",
                                    device: Union[str, torch.device] = "cpu", jit: bool = False, force_quick_gelu: bool = False,
                                    force_custom_text: bool = False, force_patch_dropout: Optional[float] = None,
                                    force_image_size=None, pretrained_image: bool = False, pretrained_hf: bool = True,
                                    image_mean=None, image_std=None, aug_cfg=None, cache_dir: Optional[str] = CKPT_CACHE_DIR,
                                    output_dict: Optional[bool] = None, strict: bool = False, args=None):
    model = tri_create_model(model_name, pretrained, precision=precision, device=device, jit=jit, force_quick_gelu=force_quick_gelu,
                             force_custom_text=force_custom_text, force_patch_dropout=force_patch_dropout,
                             force_image_size=force_image_size, pretrained_image=pretrained_image, pretrained_hf=pretrained_hf,
                             cache_dir=cache_dir, output_dict=output_dict, strict=strict, args=args)
    return (model, *_transforms(model, image_mean, image_std, aug_cfg))


def create_loss(args):
    """factory.py:750-851: the contrastive loss classes incl. the label / similarity mask variants; distillation, CoCa and
    video-token distillation (other model families) raise NotImplementedError."""
    kw = dict(local_loss=args.local_loss, gather_with_grad=args.gather_with_grad, cache_labels=True, rank=args.rank,
              world_size=args.world_size, use_horovod=getattr(args, "horovod", False))
    if getattr(args, "distill", False) or "coca" in str(getattr(args, "model", "")).lower() or getattr(args, "vid_distill_tokens", False):
        raise NotImplementedError("distillation / CoCa / video-token losses are outside the ViT-Lens hot path")
    if getattr(args, "n_tower", 2) == 3:
        loss_type = getattr(args, "contra_loss_type", "general")
        if getattr(args, "use_dual_loss", False):
            if loss_type == "general":
                return ClipLossGeneral(**kw)
            if loss_type == "label_mask":
                return ClipLossLabelMask(use_mask=True, **kw)
            if loss_type == "sim_mask":
                return ClipLossSimMask(sim_thres=args.sim_thres, **kw)
            raise NotImplementedError(loss_type)
        if loss_type == "general":
            return TriClipLoss(**kw)
        if loss_type == "label_mask":
            return TriClipLossLabelMask(**kw)
    return ClipLoss(**kw)
