"""Zero-shot evaluation parity shared by the CPU (emulated kernels) and GPU tiers: this repo's classifier builder, top-k
accuracy, mAP and retrieval recall against what the REFERENCE's own functions produced (tests/golden/zero_shot.pt)."""
import torch

from tests.common import C, cosine


def check_zero_shot(device):
    import open_clip
    from oracle.make_golden_zeroshot import CLASSNAMES, OUT, TEMPLATES, eval_inputs, toy_tokenizer
    from open_clip.metrics import MAP, Recall
    from training import zero_shot as Z
    from vitlens_b200 import synth

    gold = torch.load(OUT, map_location="cpu", weights_only=True)
    rows = {}
    # classifier from text templates (zero_shot_classifier.py:27-88)
    model = open_clip.create_model("ViT-tiny-16", device="cpu")
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0), strict=True)
    model.eval().to(device)
    clf = open_clip.build_zero_shot_classifier(model, toy_tokenizer, CLASSNAMES, TEMPLATES, num_classes_per_batch=4, device=device)
    assert tuple(clf.shape) == tuple(gold["classifier"].shape)
    rows["classifier_cos"] = round(cosine(clf.cpu(), gold["classifier"]), 6)
    assert rows["classifier_cos"] > 0.999
    assert float((clf.float().norm(dim=0) - 1).abs().max()) < 1e-4
    inp = {k: v.to(device) for k, v in eval_inputs().items()}
    # top-k accuracy: exact
    (a1, a5), correct = Z.acc(inp["logits"], inp["target"], topk=(1, 5))
    assert torch.equal(correct.cpu(), gold["correct"]) and float(a1) == float(gold["acc1"]) and float(a5) == float(gold["acc5"])
    assert Z.accuracy(inp["logits"], inp["target"], topk=(1, 5)) == [float(gold["correct"][:1].sum()), float(gold["correct"][:5].sum())]
    # mAP with ties and a class without positives (metrics/map.py)
    m = MAP()
    m.initialize(device)
    for lo in range(0, 300, 128):  # streamed in batches like a loader
        m.compute(torch.arange(lo, min(lo + 128, 300), device=device), inp["map_logits"][lo:lo + 128], inp["map_targets"][lo:lo + 128])
    st = m.merge_results()
    rows["map_err"] = abs(st["map"] - float(gold["map"]))
    assert float((st["ap_per_class"].cpu().double() - gold["ap_per_class"]).abs().max()) < 1e-6 and rows["map_err"] < 1e-6
    assert st["map_cnt"] == 300 and int(st["positives_per_class"][6]) == 0
    # retrieval recall (metrics/recall.py)
    r = Recall()
    r.initialize(text_ids=inp["txt_ids"], text_logits=inp["txt"])
    r.compute(inp["img_ids"][:25], inp["img"][:25])
    r.compute(inp["img_ids"][25:], inp["img"][25:])
    log = r.merge_results(output_predict=True)
    for k in ("txt_r1", "txt_r5", "txt_r10", "txt_r_mean", "img_r1", "img_r5", "img_r10", "img_r_mean", "r_mean", "img_count", "txt_count"):
        assert abs(float(log[k]) - float(gold["ret_" + k])) < 1e-4, (k, log[k], float(gold["ret_" + k]))
    got_txt = torch.tensor([log["predict_txt"][i] for i in range(40)])
    assert torch.equal(got_txt, gold["ret_predict_txt"])
    return rows
