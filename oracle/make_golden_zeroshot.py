"""Pin the zero-shot evaluation arithmetic against the real reference  --  TEST INFRASTRUCTURE (build container only).

    python oracle/make_golden_zeroshot.py

Runs, on CPU, the reference's own build_zero_shot_classifier (open_clip/zero_shot_classifier.py:27-88) on its tiny CLIP with the
deterministic synthetic weights, its `acc` (training/zero_shot.py:45-60), its Recall.retrieval_eval (open_clip/metrics/recall.py:
36-78) and the mAP recipe of open_clip/metrics/map.py:36-50 (sigmoid + sklearn.metrics.average_precision_score) on seeded
score / target tensors that contain exact ties, and commits the results as tests/golden/zero_shot.pt.  (The metric classes'
`initialize` methods call .cuda(); the harness sets the same attributes on CPU tensors instead -- their arithmetic methods run
unmodified.)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cases as C  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(C.GOLDEN_DIR, "zero_shot.pt")
CLASSNAMES = ["airplane", "bathtub", "bed", "bench", "bookshelf", "bottle", "bowl", "car", "chair", "cone", "cup", "curtain", "desk"]
TEMPLATES = ["a point cloud model of {}.", "there is a {} in the scene.", "a photo of a {}.", "itap of a {}.", "a 3d rendering of the {}."]


def toy_tokenizer(texts, context_length: int = 16, vocab: int = 512):
    """Deterministic stand-in for the BPE tokenizer (its merge table is not vendored): SOT, one id per word, EOT (= vocab - 1,
    the arg-max that encode_text pools on), zero padding."""
    import zlib

    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [vocab - 2] + [50 + zlib.crc32(w.encode()) % 400 for w in t.lower().replace(".", " .").split()][: context_length - 2] + [vocab - 1]
        out[i, : len(ids)] = torch.tensor(ids)
    return out


def eval_inputs():
    """Seeded scores / targets; the mAP scores are quantised so that samples tie exactly (sklearn's thresholds define the result)."""
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(64, 13, generator=g)                            # (row-wise ties have no defined order in torch.topk)
    target = torch.randint(0, 13, (64,), generator=g)
    map_logits = (torch.randn(300, 7, generator=g) * 2).round() / 8      # ties across samples of a class
    map_targets = (torch.rand(300, 7, generator=g) < 0.15).float()
    map_targets[:, 6] = 0                                                # a class without positives
    # bf16-representable features: the device path feeds the tensor cores bf16 operands, the reference multiplies in fp32
    img = torch.nn.functional.normalize(torch.randn(40, 32, generator=g), dim=-1).bfloat16().float()
    txt = torch.randn(55, 32, generator=g).bfloat16().float()            # captions enter un-normalised (zero_shot.py:745-752)
    img_ids = torch.arange(40)
    txt_ids = torch.randint(0, 40, (55,), generator=g)
    return dict(logits=logits, target=target, map_logits=map_logits, map_targets=map_targets, img=img, txt=txt, img_ids=img_ids, txt_ids=txt_ids)


def main():
    assert ref_import.available(), "needs /root/reference"
    open_clip, _, _ = ref_import.import_reference()
    from open_clip.factory import add_model_config
    from open_clip.metrics.recall import Recall
    from sklearn.metrics import average_precision_score

    sys.path.insert(0, ref_import.REF_SRC)
    import importlib.util

    # training/zero_shot.py imports dataset-side modules; its `acc` is a pure function: load just that source object
    src = open(os.path.join(ref_import.REF_SRC, "training", "zero_shot.py")).read()
    start = src.index("def acc(output, target, topk=(1,)):")
    end = src.index("def cond_acc(")
    ns = {"torch": torch}
    exec(compile(src[start:end], "training/zero_shot.py:acc", "exec"), ns)  # the reference's own text, unmodified
    ref_acc = ns["acc"]

    add_model_config(C.MODEL_CONFIG_DIR)
    model = open_clip.create_model("ViT-tiny-16", precision="fp32", device="cpu")
    synth = C._synth()
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0), strict=True)
    model.eval()
    fx = {}
    fx["classifier"] = open_clip.build_zero_shot_classifier(model, toy_tokenizer, CLASSNAMES, TEMPLATES, num_classes_per_batch=4, device="cpu").clone()
    inp = eval_inputs()
    (a1, a5), correct = ref_acc(inp["logits"], inp["target"], topk=(1, 5))
    fx["acc1"], fx["acc5"], fx["correct"] = a1.clone(), a5.clone(), correct.clone()
    preds = torch.sigmoid(inp["map_logits"]).numpy()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ap = average_precision_score(inp["map_targets"].numpy(), preds, average=None)
    fx["ap_per_class"] = torch.tensor(np.asarray(ap, dtype=np.float64))
    fx["map"] = torch.tensor(float(np.mean(ap)))
    m = Recall()
    m.text_ids, m.text_logits = inp["txt_ids"], inp["txt"]
    m.image_ids, m.image_logits = inp["img_ids"], inp["img"]
    sim = m.image_logits @ m.text_logits.t()
    log = m.retrieval_eval(sim, sim.t(), output_predict=True)
    for k, v in log.items():
        if not isinstance(v, dict):
            fx["ret_" + k] = torch.tensor(float(v))
    fx["ret_predict_txt"] = torch.tensor([log["predict_txt"][i] for i in range(40)])
    torch.save(fx, OUT)
    print({k: (tuple(v.shape) if v.dim() else float(v)) for k, v in fx.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
