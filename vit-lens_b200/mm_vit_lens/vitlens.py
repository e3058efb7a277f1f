"""One-stop inference API (reference mm_vit_lens/vitlens.py:21-189): an nn.ModuleDict with one tower per
modality and `encode(input_dict, normalize=True)`.

Differences that follow from the scope: file -> tensor processors (PIL / torchaudio / numpy
loaders, reference mm_vit_lens/data_processors.py) are host-side I/O and not part of this package, so
`encode` takes already-processed tensors (what the reference's processors return) unless a callable
processor is registered in `self.processors[modality]`; no checkpoint download (offline)."""
from __future__ import annotations

import logging
import os
from functools import partial

import torch
import torch.nn as nn

from open_clip.factory import tri_create_model_and_transforms
from open_clip.model import _normalize
from open_clip.transformer import encode_text_tokens

from .model_cfg import fetch_model_cfg


class ViTLens(nn.Module):
    def __init__(self, model_var="vitlensL", modality_loaded=("image", "text", "pc", "depth", "audio", "tactile", "eeg"),
                 load_from_ckpt=None, device=None):
        super().__init__()
        self.model_var = model_var
        self.modality_loaded = list(modality_loaded)
        self.processors = dict()
        self.vitlens = nn.ModuleDict()
        self._device_hint = device
        self.init_processors_and_model(load_from_ckpt=load_from_ckpt)

    def _init_modality_module(self, modality, load_from_pt_flag=False):
        cfg = fetch_model_cfg(modality=modality, model_option=self.model_var)
        for k in ("unlock_from_head", "vid_use_fpos", "vid_use_ltpos", "vid_distill_tokens"):
            setattr(cfg, k, False)
        model, _, image_process_val = tri_create_model_and_transforms(
            cfg.model, None, precision=cfg.precision, device=self._device_hint or cfg.device, jit=False,
            force_quick_gelu=cfg.force_quick_gelu, force_custom_text=cfg.force_custom_text, force_patch_dropout=None,
            force_image_size=cfg.force_image_size, pretrained_image=cfg.pretrained_image, output_dict=True,
            cache_dir=cfg.cache_dir, args=cfg)
        if modality == "image":
            self.vitlens.add_module("image", model.image)
            self.processors.setdefault("image", image_process_val)
        elif modality == "text":
            text = nn.ModuleDict()
            self.vitlens.add_module("text", text)
            text.add_module("transformer", model.transformer)
            text.context_length = model.context_length
            text.vocab_size = model.vocab_size
            text.add_module("token_embedding", model.token_embedding)
            text.positional_embedding = model.positional_embedding
            text.add_module("ln_final", model.ln_final)
            text.text_projection = model.text_projection
            text.register_buffer("attn_mask", model.attn_mask, persistent=False)

            def encode_text(module, tokens):
                return encode_text_tokens(tokens, module.token_embedding, module.positional_embedding, module.transformer,
                                          module.ln_final, module.text_projection, module.attn_mask)

            text.forward = partial(encode_text, text)
        else:
            self.vitlens.add_module(modality, model.visual)
            if load_from_pt_flag:
                self.load_modality_from_pt_ckpt(modality=modality, pt_ckpt_path=cfg.ckpt_pth)
        del model

    def init_processors_and_model(self, load_from_ckpt=None):
        for m in self.modality_loaded:
            self._init_modality_module(m)
        if load_from_ckpt:
            ckpt_path = os.path.join(load_from_ckpt, f"{self.model_var}.pt")
            if not os.path.exists(ckpt_path):
                raise FileNotFoundError(f"{ckpt_path} not found (no network access: place the released checkpoint there)")
            ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            msg = self.load_state_dict(ckpt["state_dict"], strict=False)
            logging.info(msg)

    def load_modality_from_pt_ckpt(self, modality, pt_ckpt_path):
        checkpoint = torch.load(pt_ckpt_path, map_location="cpu", weights_only=False)
        sd = checkpoint["state_dict"]
        if next(iter(sd.items()))[0].startswith("module."):
            sd = {k[len("module."):]: v for k, v in sd.items()}
        sd = {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
        msg = self.vitlens[modality].load_state_dict(sd, strict=False)
        print(f"[Load ViT-Lens from `{modality}` Pretrained ckpt] : {msg}.")

    def export_checkpoint(self, save_path="model_release/vitlens.pt"):
        torch.save(dict(model_var=self.model_var, modality_loaded=self.modality_loaded, state_dict=self.state_dict()), save_path)

    @property
    def device(self):
        return list(self.parameters())[0].device

    def reduce_list(self, modality):
        return modality in ["audio"]

    def encode(self, input_dict, normalize=True):
        """{modality: tensor | raw input for a registered processor} -> {modality: [B, embed_dim] features}.
        Audio arrives as [B, S clips, T, F] and is averaged over the clips after encoding (vitlens.py:175-183)."""
        output_dict = dict()
        for m, x in input_dict.items():
            proc = self.processors.get(m)
            if proc is not None and not torch.is_tensor(x):
                x = proc(x, device=self.device)
            x = x.to(self.device)
            B = S = None
            if self.reduce_list(m):
                B, S = x.shape[:2]
                x = x.reshape((B * S,) + tuple(x.shape[2:]))
            features = self.vitlens[m](x)
            if self.reduce_list(m):
                features = features.reshape(B, S, -1).mean(dim=1)
            output_dict[m] = _normalize(features) if normalize else features
        return output_dict
