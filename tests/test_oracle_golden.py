"""The CPU oracle reproduces the REFERENCE's outputs (fixtures written by oracle/make_golden.py from the
real reference) on every case: features, loss and gradients."""
import zlib

import pytest
import torch

from tests.common import C, build_model, relerr, run_oracle

FAST = ["tiny_clip", "tiny_tri_audio", "tiny_tri_depth", "tiny_tri_pc", "tiny_tri_pc_bntrain", "vitb32_clip_bs8",
        "tiny_tri_eeg", "tiny_tri_tactile", "tiny_tri_audio_as_transformer", "tiny_tri_depth_frames"]
SLOW = ["vitl14_audio128_bs2", "vitl14_depth_bs2", "vitl14_pc_bs2", "vitl14_pc_bs8_bntrain"]


def _check(name, with_grads):
    case = C.CASES[name]
    gold = C.load_golden(name)
    model, sd, args = build_model(case)
    # recipe / schema drift guards: same weights and inputs as the reference saw
    chk_w = sum(float(v.double().abs().sum()) for v in sd.values() if v.is_floating_point())
    assert abs(chk_w - float(gold["chk_weights"])) <= 1e-6 * abs(chk_w)
    inp = C.build_inputs(case, args)
    if "fps_start" in gold:
        inp["fps_start"] = gold["fps_start"]
    assert abs(sum(float(v.double().abs().sum()) for v in inp.values()) - float(gold["chk_inputs"])) <= 1e-5 * float(gold["chk_inputs"])
    keys = sorted(k for k, p in model.named_parameters() if p.requires_grad) if with_grads else []
    if with_grads and case.kind == "tri" and case.modality != "pc":
        assert zlib.crc32("\n".join(keys).encode()) == int(gold["grad_keys_crc"])
    new_stats = {}
    feats, ls, loss, grads = run_oracle(case, sd, args, inp, set(keys), new_stats)
    if case.bn_train:  # training-mode BatchNorm: the running statistics after one forward match the reference's buffers
        pre = "visual.visual_adapter.encoder."
        ours = torch.cat([t.float().flatten() for q in ("first_conv.1.", "second_conv.1.") for t in new_stats[pre + q]])
        assert relerr(ours, gold["bn_running"]) < 1e-4
    for k, v in feats.items():
        assert relerr(v.detach(), gold[k]) < 2e-4, k
    assert abs(float(loss) - float(gold["loss"])) < 1e-4 * abs(float(gold["loss"]))
    if with_grads:
        norms = torch.tensor([float(grads[k].norm()) for k in keys])
        if norms.numel() == gold["grad_norms"].numel():
            assert relerr(norms, gold["grad_norms"]) < 2e-3
        for k in keys:
            if "grad:" + k in gold and float(gold["grad:" + k].abs().max()) > 1e-6:  # (exactly-zero gradients hold rounding noise)
                assert relerr(grads[k], gold["grad:" + k]) < 2e-3, k


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_fixture(name):
    _check(name, with_grads=True)


@pytest.mark.parametrize("name", SLOW)
def test_oracle_matches_reference_fixture_vitl(name):
    _check(name, with_grads=False)


def test_oracle_sharded_loss_matches_reference_run_under_gloo():
    """oracle.clip_loss_sharded (the restatement of gather_features + ClipLoss / TriClipLoss at world_size > 1) against the
    per-rank results of the REAL reference run as two gloo processes (oracle/make_golden_dist.py), all four
    (local_loss, gather_with_grad) combinations, both loss classes: loss, d/d(local features), d/d(logit_scale)."""
    from tests import dist_common as DC

    gold = DC.load_dist_golden()
    W, bl, e = int(gold["world"]), int(gold["bl"]), int(gold["e"])
    for tri, ll, gwg in DC.COMBOS:
        name = DC.combo_name(tri, ll, gwg)
        ours = DC.oracle_per_rank(tri, ll, gwg, W, bl, e, float(gold["scale_log"]), tuple(gold["seeds"]))
        for r in range(W):
            for k, v in ours[r].items():
                if v is None:
                    continue
                ref = gold[f"{name}/rank{r}/{k}"]
                err = float((v.detach() - ref).abs().max())
                assert err <= 2e-5 * float(ref.abs().max()) + 1e-7, (name, r, k, err)


def test_oracle_mask_losses_match_reference_run():
    """oracle.sim_mask / label_mask / clip_loss_sharded(mask=) against the REAL reference's ClipLossSimMask, ClipLossLabelMask and
    TriClipLossLabelMask (loss.py:485-903; oracle/make_golden_maskloss.py): world size 1 and per rank under two gloo processes."""
    from tests import maskloss_common as MC

    gold = MC.load_golden()
    W = int(gold["world"])
    for kind in MC.KINDS:
        ours = MC.oracle_per_rank(gold, kind, False, False, 1)[0]
        MC.compare({k: (v.detach() if v is not None else None) for k, v in ours.items()}, gold, f"{kind}_w1", 2e-5, 2e-5)
        for ll, gwg in MC.FLAGS:
            per = MC.oracle_per_rank(gold, kind, ll, gwg, W)
            for r in range(W):
                MC.compare({k: (v.detach() if v is not None else None) for k, v in per[r].items()}, gold,
                           f"{kind}_local{int(ll)}_gwg{int(gwg)}/rank{r}", 2e-5, 2e-5)
