"""Parity cases shared by oracle/make_golden.py (reference side) and tests/ (oracle + CUDA
side)  --  TEST INFRASTRUCTURE.  Inputs and weights are regenerated from the deterministic
recipes in vitlens_b200.synth; only the reference's *outputs* are committed as fixtures."""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SRC = os.path.join(_ROOT, "vit-lens_b200")
GOLDEN_DIR = os.path.join(_ROOT, "tests", "golden")
MODEL_CONFIG_DIR = os.path.join(_SRC, "open_clip", "model_configs")


def _synth():
    # load the recipe module by path: this file must stay importable in a process that has the
    # *reference's* open_clip on sys.path (make_golden.py), so the product package is not imported.
    import importlib.util

    name = "_vitlens_synth_recipe"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_SRC, "vitlens_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


@dataclass
class Case:
    name: str
    model: str  # model_configs/<model>.json
    kind: str  # "clip" | "tri"
    batch: int
    modality: Optional[str] = None  # tri: audio | depth | pc
    overrides: Dict = field(default_factory=dict)  # fields of the reference's `args`
    lock: Dict = field(default_factory=dict)  # kwargs of lock_visual_tower for tri models
    full_grads: bool = True  # store every gradient (tiny) or only norms + a few small ones
    bn_train: bool = False  # point tokenizer: BatchNorm layers in training mode (batch statistics), as under model.train()
    image_frames: int = 0  # > 0: the image input is [B, t, 3, H, W] (per-frame encode + mean aggregation, model.py:591-604)
    shapes: bool = False  # point clouds of distinct shapes (synth_shapes) instead of uniform balls: a well-conditioned contrastive batch
    seed: int = 0


_TINY_LENS = dict(perceiver_input_chan=128, perceiver_latent_dim=128, perceiver_latent_heads=2,
                  perceiver_num_latents=16, perceiver_depth=2, perceiver_self_per_cross_attn=2)

CASES = {c.name: c for c in [
    Case("tiny_clip", "ViT-tiny-16", "clip", 4),
    Case("tiny_tri_audio", "ViT-tiny-16", "tri", 4, "audio",
         dict(_TINY_LENS, audio_mel_bins=32, audio_target_length=48), dict(unlock_cls=True)),
    Case("tiny_tri_depth", "ViT-tiny-16", "tri", 4, "depth",
         dict(perceiver_num_latents=16), dict(unlock_cls=True, unlock_trans_first_n_layers=1)),
    Case("tiny_tri_pc", "ViT-tiny-16", "tri", 3, "pc",
         dict(_TINY_LENS, perceiver_input_chan=96, perceiver_self_per_cross_attn=1, pc_npoints=256,
              pc_num_group=16, pc_group_size=8, pc_trans_dim=96, pc_encoder_dims=64), dict(unlock_cls=True)),
    # the same with the tokenizer's BatchNorm layers in training mode (batch statistics + running-statistics update)
    Case("tiny_tri_pc_bntrain", "ViT-tiny-16", "tri", 3, "pc",
         dict(_TINY_LENS, perceiver_input_chan=96, perceiver_self_per_cross_attn=1, pc_npoints=256,
              pc_num_group=16, pc_group_size=8, pc_trans_dim=96, pc_encoder_dims=64), dict(unlock_cls=True), bn_train=True),
    # SURVEY 8(f).2 variants on the same kernels: EEG Conv1d patch embed; tactile = image path with the last `unlocked_groups`
    # ViT groups trainable; the Lens replaced by a plain Transformer (perceiver_as_transformer); 5-D image input
    Case("tiny_tri_eeg", "ViT-tiny-16", "tri", 4, "eeg",
         dict(_TINY_LENS, perceiver_depth=1, perceiver_self_per_cross_attn=1, eeg_chans=16, eeg_time_len=34, eeg_window_size=4, eeg_stride=2),
         dict(unlock_cls=True)),
    Case("tiny_tri_tactile", "ViT-tiny-16", "tri", 4, "tactile", {}, dict(unlocked_groups=2)),
    Case("tiny_tri_audio_as_transformer", "ViT-tiny-16", "tri", 4, "audio",
         dict(_TINY_LENS, perceiver_as_transformer=True, perceiver_num_latents=8, audio_mel_bins=32, audio_target_length=48), dict(unlock_cls=True)),
    Case("tiny_tri_depth_frames", "ViT-tiny-16", "tri", 3, "depth", dict(perceiver_num_latents=16), dict(unlock_cls=True), image_frames=2),
    # BASELINE.json configs[0]: ViT-B/32 image-text ClipLoss, batch 8
    Case("vitb32_clip_bs8", "ViT-B-32", "clip", 8, full_grads=False),
    # reduced-batch versions of configs[2..4] (full-size weights, reference runs them in seconds)
    Case("vitl14_audio128_bs2", "ViT-L-14", "tri", 2, "audio", dict(perceiver_num_latents=128),
         dict(unlock_cls=True), full_grads=False),
    Case("vitl14_depth_bs2", "ViT-L-14", "tri", 2, "depth", {}, dict(unlock_cls=True, unlock_trans_first_n_layers=4),
         full_grads=False),
    Case("vitl14_pc_bs2", "ViT-L-14", "tri", 2, "pc", {}, dict(unlock_cls=True), full_grads=False),
    # Training-mode BatchNorm at full size.  Eight clouds of distinct shapes: batch statistics over two look-alike uniform balls
    # (the first version of this case) left the visual features almost identical (loss = ln 4), so every gradient was a
    # difference of near-equal terms and even the CPU emulation of the bf16 contract only reached cosine 0.91 against fp32.
    Case("vitl14_pc_bs8_bntrain", "ViT-L-14", "tri", 8, "pc", {}, dict(unlock_cls=True), full_grads=False, bn_train=True, shapes=True),
]}


BN_KEYS = ("first_conv.1.running_mean", "first_conv.1.running_var", "second_conv.1.running_mean", "second_conv.1.running_var")


def set_bn_train(model) -> None:
    """Training-mode BatchNorm in the point tokenizer only (everything else stays in eval mode: no dropout noise)."""
    for m in model.visual.visual_adapter.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.train()


def bn_running(state_dict) -> torch.Tensor:
    """The tokenizer's four running-statistics buffers, concatenated (checked after a training-mode forward)."""
    pre = "visual.visual_adapter.encoder."
    return torch.cat([state_dict[pre + k].detach().float().cpu().flatten() for k in BN_KEYS])


def model_cfg(case: Case) -> dict:
    import json

    with open(os.path.join(MODEL_CONFIG_DIR, case.model + ".json")) as f:
        return json.load(f)


def build_inputs(case: Case, args=None) -> Dict[str, torch.Tensor]:
    """`args`: any object with the modality fields (audio_target_length, audio_mel_bins, pc_npoints);
    falls back to the vitlensL defaults (mm_vit_lens/model_cfg.py:9-78)."""
    s = _synth()
    cfg = model_cfg(case)
    img = cfg["vision_cfg"]["image_size"]
    ctx, vocab = cfg["text_cfg"]["context_length"], cfg["text_cfg"]["vocab_size"]
    B = case.batch
    out = {
        "image": s.synth_normal("image", (B, case.image_frames, 3, img, img) if case.image_frames else (B, 3, img, img), seed=case.seed + 1),
        "text": s.synth_text(B, ctx, vocab, seed=case.seed + 1),
    }
    ov = case.overrides

    def g(k, d):
        return ov.get(k, getattr(args, k, d) if args is not None else d)

    if case.modality == "audio":
        out["visual"] = s.synth_normal("audio", (B, g("audio_target_length", 512), g("audio_mel_bins", 128)), seed=case.seed + 1)
    elif case.modality == "depth":
        out["visual"] = s.synth_normal("depth", (B, 1, img, img), seed=case.seed + 1)
    elif case.modality == "tactile":
        out["visual"] = s.synth_normal("tactile", (B, 3, img, img), seed=case.seed + 1)
    elif case.modality == "eeg":
        out["visual"] = s.synth_normal("eeg", (B, g("eeg_chans", 128), g("eeg_time_len", 512)), seed=case.seed + 1)
    elif case.modality == "pc":
        pts, start = (s.synth_shapes if case.shapes else s.synth_points)(B, g("pc_npoints", 8192), seed=case.seed + 1)
        out["visual"] = pts
        out["fps_start"] = start
    return out


def golden_path(case: Case) -> str:
    return os.path.join(GOLDEN_DIR, case.name + ".pt")


def load_golden(name: str) -> Dict[str, torch.Tensor]:
    return torch.load(golden_path(CASES[name]), map_location="cpu", weights_only=True)
