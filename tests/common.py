"""Shared helpers for the parity tests: build this package's model for an oracle case, load the
deterministic synthetic weights, run forward/loss/backward, run the oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "vit-lens_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import cases as C  # noqa: E402
from oracle import vitlens_oracle as O  # noqa: E402


def case_args(case):
    from mm_vit_lens.model_cfg import training_args

    return training_args(case.modality, **case.overrides) if case.kind == "tri" else None


def build_model(case, device="cpu"):
    import open_clip
    from vitlens_b200 import synth

    args = case_args(case)
    if case.kind == "clip":
        model = open_clip.create_model(case.model, device="cpu")
    else:
        model = open_clip.tri_create_model(case.model, None, device="cpu", args=args)
    sd = synth.synth_state_dict(model.state_dict(), seed=case.seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    if case.kind == "tri":
        model.lock_image_tower()
        model.lock_text_tower()
        model.lock_visual_tower(**case.lock)
    if case.bn_train:
        C.set_bn_train(model)
    model.to(device)
    return model, sd, args


def run_model(case, model, inp, loss_mod=None):
    import open_clip

    dev = next(model.parameters()).device
    inp = {k: v.to(dev) for k, v in inp.items()}
    if case.kind == "clip":
        fi, ft, ls = model(inp["image"], inp["text"])
        loss = (loss_mod or open_clip.ClipLoss())(fi, ft, ls)
        feats = {"image_features": fi, "text_features": ft}
    else:
        if "fps_start" in inp:  # point clouds: the FPS start indices are an explicit input (misc.py:60 draws them at random)
            fi = model.encode_image(inp["image"], normalize=True)
            ft = model.encode_text(inp["text"], normalize=True)
            fv = model.encode_visual(inp["visual"], normalize=True, fps_start=inp["fps_start"])
            ls = model.logit_scale.exp()
        else:
            fi, ft, fv, ls = model(inp["image"], inp["text"], inp["visual"])
        loss = (loss_mod or open_clip.TriClipLoss())(fi, ft, fv, ls)
        feats = {"image_features": fi, "text_features": ft, "visual_features": fv}
    return feats, ls, loss


def run_oracle(case, sd, args, inp, grad_keys=(), new_stats=None):
    cfg = C.model_cfg(case)
    vh = cfg["vision_cfg"]["width"] // 64
    th = cfg["text_cfg"]["heads"]
    sd = {k: (v.clone().requires_grad_(True) if k in grad_keys else v) for k, v in sd.items()}
    if case.kind == "clip":
        fi, ft, ls = O.clip_forward(sd, inp["image"], inp["text"], vh, th)
        loss = O.clip_loss(fi, ft, ls)
        feats = {"image_features": fi, "text_features": ft}
    else:
        kw = {}
        if case.modality == "audio":
            kw = dict(fstride=args.audio_fstride, tstride=args.audio_tstride)
        if case.modality == "eeg":
            kw = dict(stride=args.eeg_stride)
        if case.modality == "pc":
            kw = dict(fps_start=inp["fps_start"], num_group=args.pc_num_group, group_size=args.pc_group_size)
            if case.bn_train:
                kw.update(bn_train=True, new_stats=new_stats if new_stats is not None else {})
        fi, ft, fv, ls = O.triclip_forward(sd, inp["image"], inp["text"], inp["visual"], case.modality, vh, th,
                                           perceiver_as_identity=bool(args.perceiver_as_identity), perceiver_as_transformer=bool(args.perceiver_as_transformer),
                                           latent_heads=args.perceiver_latent_heads, cross_heads=args.perceiver_cross_heads, **kw)
        loss = O.tri_clip_loss(fi, ft, fv, ls)
        feats = {"image_features": fi, "text_features": ft, "visual_features": fv}
    grads = {}
    if grad_keys:
        loss.backward()
        grads = {k: sd[k].grad for k in grad_keys}
    return feats, ls, loss, grads


def relerr(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def cosine(a, b):
    return float(torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0))
