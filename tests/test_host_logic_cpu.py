"""Host logic (open_clip modules + vitlens_b200.engine autograd stages) on CPU with the kernel gateway
swapped for tests/emu_ops.py, against the reference's golden outputs.  Tolerances are the bf16 ones."""
import pytest
import torch

from tests.common import C, build_model, cosine, relerr, run_model

CASES = ["tiny_clip", "tiny_tri_audio", "tiny_tri_depth", "tiny_tri_pc", "tiny_tri_pc_bntrain",
         "tiny_tri_eeg", "tiny_tri_tactile", "tiny_tri_audio_as_transformer", "tiny_tri_depth_frames"]


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_vs_reference(name, emu):
    case = C.CASES[name]
    gold = C.load_golden(name)
    model, sd, args = build_model(case)
    inp = C.build_inputs(case, args)
    if "fps_start" in gold:
        inp["fps_start"] = gold["fps_start"]
    feats, ls, loss = run_model(case, model, inp)
    for k, v in feats.items():
        assert cosine(v, gold[k]) > 0.999, (k, cosine(v, gold[k]))
    assert abs(float(loss.detach()) - float(gold["loss"])) < 2e-2 * abs(float(gold["loss"]))
    if case.bn_train:  # running statistics after one training-mode forward (bf16 activations: 1 % of the largest entry)
        assert relerr(C.bn_running(model.state_dict()), gold["bn_running"]) < 1e-2
    loss.backward()
    got = {k: p.grad for k, p in model.named_parameters() if p.requires_grad}
    assert all(g is not None for g in got.values()), [k for k, g in got.items() if g is None]
    keys = sorted(got)
    norms = torch.tensor([float(got[k].norm()) for k in keys])
    assert norms.numel() == gold["grad_norms"].numel()
    bad = []
    for i, k in enumerate(keys):
        gk = "grad:" + k
        if gk in gold and float(gold["grad_norms"][i]) > 1e-4:  # (mathematically zero gradients hold rounding noise)
            c = cosine(got[k], gold[gk])
            if c < 0.98:
                bad.append((k, c))
        rn = abs(float(norms[i]) - float(gold["grad_norms"][i])) / max(float(gold["grad_norms"][i]), 1e-6)
        if rn > 0.1 and float(gold["grad_norms"][i]) > 1e-4:
            bad.append((k, "norm", float(norms[i]), float(gold["grad_norms"][i])))
    assert not bad, bad


def test_frozen_towers_save_nothing(emu):
    case = C.CASES["tiny_tri_audio"]
    model, sd, args = build_model(case)
    inp = C.build_inputs(case, args)
    with torch.no_grad():
        f = model.encode_image(inp["image"], normalize=True)
    assert not f.requires_grad


def test_point_cloud_tower_forward_vs_reference(emu):
    """tiny_tri_pc: FPS / kNN / grouped PointNet host logic (frozen tokenizer) against the reference's features."""
    case = C.CASES["tiny_tri_pc"]
    gold = C.load_golden("tiny_tri_pc")
    model, sd, args = build_model(case)
    inp = C.build_inputs(case, args)
    with torch.no_grad():
        fv = model.encode_visual(inp["visual"], normalize=True, fps_start=gold["fps_start"])
    assert cosine(fv, gold["visual_features"]) > 0.999


def test_weight_cache_entries_die_with_their_parameter(emu):
    """engine.WEIGHTS hands out bf16 operand copies keyed by the parameter OBJECT: an entry must not outlive its parameter
    (a later model can reuse the freed parameter's id, storage address, version and shape) and must follow in-place updates."""
    import gc

    from vitlens_b200 import engine

    engine.WEIGHTS.clear()
    p = torch.nn.Parameter(torch.randn(8, 16))
    q = torch.nn.Parameter(torch.randn(4, 16))
    a = engine.w16(p)
    assert engine.w16(p) is a and torch.equal(a.float(), p.detach().bfloat16().float())
    cat = engine._cat16("qkv", p, q)
    assert engine._cat16("qkv", p, q) is cat and cat.shape == (12, 16)
    with torch.no_grad():
        p.mul_(2.0)  # version bump -> refreshed copy
    b = engine.w16(p)
    assert b is not a and torch.equal(b.float(), p.detach().bfloat16().float())
    n_before = len(engine.WEIGHTS._c)
    assert n_before >= 2
    del p, a, b, cat
    gc.collect()
    assert all(all(r() is not None for r in hit[2]) for hit in engine.WEIGHTS._c.values())  # no entry points at a dead parameter
    assert len(engine.WEIGHTS._c) < n_before


def test_vitlens_encode_vs_reference_class(emu):
    """ViTLens.encode host logic (text closure, audio clip mean, checkpoint wire-format keys) with emulated kernels against
    the reference's own ViTLens run (tests/golden/vitlens_encode.pt)."""
    from tests.api_common import check_vitlens_encode

    check_vitlens_encode("cpu")


def test_zero_shot_eval_vs_reference_functions(emu, monkeypatch):
    """Zero-shot evaluation host logic (classifier from templates, top-k accuracy, mAP, retrieval recall; training/zero_shot.py,
    open_clip/metrics/*) with emulated kernels against the reference's own functions (tests/golden/zero_shot.pt)."""
    import open_clip.metrics.accuracy as A
    import open_clip.metrics.map as M
    import open_clip.metrics.recall as R
    import open_clip.zero_shot_classifier as ZC
    import training.zero_shot as Z

    for mod in (A, M, R, ZC, Z):
        monkeypatch.setattr(mod, "_ops", emu)
    from tests.zeroshot_common import check_zero_shot

    check_zero_shot("cpu")


def test_grad_checkpointing_reproduces_gradients(emu):
    """set_grad_checkpointing (transformer.py:366-368): every block keeps only its input and is re-run inside backward; loss and
    gradients are those of the plain run (the kernels are deterministic, so exactly)."""
    case = C.CASES["tiny_tri_depth"]

    def run(ckpt):
        model, sd, args = build_model(case)
        model.set_grad_checkpointing(ckpt)
        assert model.visual.transformer.grad_checkpointing is ckpt
        inp = C.build_inputs(case, args)
        feats, ls, loss = run_model(case, model, inp)
        loss.backward()
        return float(loss.detach()), {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad}

    l0, g0 = run(False)
    l1, g1 = run(True)
    assert l0 == l1 and g0.keys() == g1.keys()
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k


def test_mask_losses_vs_reference_run(emu):
    """ClipLossSimMask / ClipLossLabelMask / TriClipLossLabelMask (loss.py:485-903) at world size 1 against the REAL reference's
    results (tests/golden/mask_loss.pt): the masks are built on the host side of the kernels, `logits * mask` is applied inside
    the loss epilogues."""
    import open_clip
    from tests import maskloss_common as MC

    gold = MC.load_golden()
    X, Y, V, LX, LY, LV = (MC.flat(t) for t in MC.inputs(gold))
    for kind in MC.KINDS:
        got = MC.run_ours(open_clip, gold, kind, dict(world_size=1), X, Y, V, LX, LY, LV)
        MC.compare(got, gold, f"{kind}_w1")


def test_create_loss_selects_the_reference_classes():
    """factory.create_loss (factory.py:420-470): contra_loss_type / use_dual_loss pick the same classes as the reference."""
    from types import SimpleNamespace

    import open_clip

    base = dict(local_loss=False, gather_with_grad=False, rank=0, world_size=1, horovod=False, distill=False, model="ViT-L-14", n_tower=3,
                sim_thres=0.9)
    pick = lambda **kw: type(open_clip.create_loss(SimpleNamespace(**{**base, **kw}))).__name__  # noqa: E731
    assert pick(contra_loss_type="general", use_dual_loss=False) == "TriClipLoss"
    assert pick(contra_loss_type="general", use_dual_loss=True) == "ClipLossGeneral"
    assert pick(contra_loss_type="label_mask", use_dual_loss=False) == "TriClipLossLabelMask"
    assert pick(contra_loss_type="label_mask", use_dual_loss=True) == "ClipLossLabelMask"
    assert pick(contra_loss_type="sim_mask", use_dual_loss=True) == "ClipLossSimMask"
    assert pick(n_tower=2) == "ClipLoss"
