"""Checkpoint wire formats (reference pc_tri_main.py:580-611, factory.py:119-160, vitlens.py:153-159)."""
import torch

from tests.common import C, build_model


def test_training_checkpoint_round_trip_and_ddp_prefix(tmp_path):
    from vitlens_b200 import checkpoint as K

    case = C.CASES["tiny_tri_audio"]
    model, sd, args = build_model(case)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3)  # the reference's optimizer: its state_dict is the wire format
    for p in params:
        p.grad = torch.ones_like(p)
    opt.step()
    path = str(tmp_path / "epoch_3.pt")
    K.save_checkpoint(path, model, opt, epoch=3, name="run", best_acc=12.5, ddp_prefix=True)
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"epoch", "name", "state_dict", "optimizer", "best_acc"} and next(iter(raw["state_dict"])).startswith("module.")
    model2, _, _ = build_model(case)
    with torch.no_grad():
        for p in model2.parameters():
            p.add_(1.0)
    opt2 = torch.optim.AdamW([p for p in model2.parameters() if p.requires_grad], lr=5e-4)
    info = K.load_checkpoint(path, model2, opt2)
    assert info["epoch"] == 3 and info["best_acc"] == 12.5 and not info["incompatible_keys"].missing_keys
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    assert opt2.state_dict()["param_groups"][0]["lr"] == 1e-3
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert all(torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"]) for i in s1)


def test_factory_load_checkpoint_copies_visual_to_image_and_resizes_pos_embed(tmp_path):
    """factory.load_checkpoint on an open_clip CLIP checkpoint: `visual.*` is duplicated to `image.*` for tri-models
    (factory.py:141-152) and a positional embedding of another grid is resampled for the Lens tower (model.py:1079-1146)."""
    import open_clip
    from vitlens_b200 import synth

    clip = open_clip.create_model("ViT-tiny-16", device="cpu")
    sd = synth.synth_state_dict(clip.state_dict(), seed=5)
    path = str(tmp_path / "clip.pt")
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, path)
    case = C.CASES["tiny_tri_depth"]  # Lens tower with 16 latents: visual.positional_embedding [17, 128] == the CLIP grid (4x4 + 1)
    from tests.common import case_args

    tri = open_clip.tri_create_model("ViT-tiny-16", path, device="cpu", args=case_args(case))
    assert torch.equal(tri.image.conv1.weight, sd["visual.conv1.weight"]) and torch.equal(tri.visual.proj, sd["visual.proj"])
    assert torch.equal(tri.visual.positional_embedding, sd["visual.positional_embedding"])
    # another latent count: the grid part is bicubically resampled 4x4 -> 3x3, the class-token row is kept
    args9 = case_args(case)
    args9.perceiver_num_latents = 9
    tri9 = open_clip.tri_create_model("ViT-tiny-16", path, device="cpu", args=args9)
    assert tuple(tri9.visual.positional_embedding.shape) == (10, 128)
    assert torch.equal(tri9.visual.positional_embedding[0], sd["visual.positional_embedding"][0])
    want = torch.nn.functional.interpolate(sd["visual.positional_embedding"][1:].reshape(1, 4, 4, -1).permute(0, 3, 1, 2), size=(3, 3),
                                           mode="bicubic", antialias=True, align_corners=False).permute(0, 2, 3, 1).reshape(9, -1)
    assert torch.allclose(tri9.visual.positional_embedding[1:], want)
    assert torch.equal(tri9.image.positional_embedding, sd["visual.positional_embedding"])  # the image tower keeps the original
    # a path that does not exist must raise, not silently build a random frozen ViT
    import pytest

    with pytest.raises(RuntimeError):
        open_clip.tri_create_model("ViT-tiny-16", str(tmp_path / "missing.pt"), device="cpu", args=case_args(case))
