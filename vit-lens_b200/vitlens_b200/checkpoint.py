"""The reference's checkpoint wire formats, so weights and optimizer state move between the two code bases unchanged.

Training checkpoints (training/point_cloud/pc_tri_main.py:580-611, training/main.py): a dict with "epoch", "name", "state_dict"
(model.state_dict(), keys prefixed "module." when the model was wrapped in DDP), "optimizer" (torch.optim.AdamW.state_dict()
layout -- vitlens_b200.optim.AdamW reads and writes the same), "best_acc" and optionally "scaler".  Release checkpoints
(mm_vit_lens/vitlens.py:153-159): {"model_var", "modality_loaded", "state_dict"} with keys `vitlens.<modality>.*`
(ViTLens.export_checkpoint / init_processors_and_model handle those)."""
from __future__ import annotations

import os
from typing import Optional

import torch


def strip_module_prefix(state_dict):
    """factory.py:125-126: checkpoints saved from a DistributedDataParallel wrapper carry a "module." prefix."""
    if state_dict and next(iter(state_dict)).startswith("module."):
        return {k[len("module."):]: v for k, v in state_dict.items()}
    return state_dict


def save_checkpoint(path: str, model, optimizer=None, *, epoch: int = 0, name: str = "", best_acc: float = 0.0, scaler=None, ddp_prefix: bool = False):
    """Write a training checkpoint in the reference's format (atomically: tmp file + rename, as pc_tri_main.py:604-610 does for
    the latest checkpoint)."""
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    if ddp_prefix:
        sd = {"module." + k: v for k, v in sd.items()}
    ckpt = {"epoch": epoch, "name": name, "state_dict": sd, "best_acc": best_acc}
    if optimizer is not None:
        ckpt["optimizer"] = optimizer.state_dict()
    if scaler is not None:
        ckpt["scaler"] = scaler.state_dict()
    tmp = path + ".tmp"
    torch.save(ckpt, tmp)
    os.replace(tmp, path)
    return ckpt


def load_checkpoint(path: str, model, optimizer=None, *, strict: bool = True, map_location="cpu") -> dict:
    """Resume from a training checkpoint written by either code base (training/main.py:360-385: epoch, model, optimizer).
    Returns {"epoch", "best_acc", "name", "incompatible_keys"}.  A bare state_dict file (no "state_dict" key) loads the model only."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    if isinstance(ckpt, dict) and "state_dict" in ckpt:
        sd = strip_module_prefix(ckpt["state_dict"])
        inc = model.load_state_dict(sd, strict=strict)
        if optimizer is not None and "optimizer" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer"])
        return {"epoch": ckpt.get("epoch", 0), "best_acc": ckpt.get("best_acc", 0.0), "name": ckpt.get("name", ""), "incompatible_keys": inc}
    inc = model.load_state_dict(strip_module_prefix(ckpt), strict=strict)
    return {"epoch": 0, "best_acc": 0.0, "name": "", "incompatible_keys": inc}
