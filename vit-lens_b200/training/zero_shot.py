"""Zero-shot evaluation on the device (reference training/zero_shot.py).  The reference's evaluators differ only in where the
class names / templates / batches come from (datasets and JSON files that are out of scope here); their arithmetic is shared:

  accuracy / acc                 :36-60    top-k correctness of a logits block
  run / zero_shot_eval           :84-152   100 * image_features @ classifier, top-1 / top-5
  test_zeroshot_3d_core          :155-257  class text features, modality features @ text^T, top-1 / top-5 + per-class accuracy
  test_audio_single_map/cls/ret  :572-789  MAP / Accuracy / Recall metrics over a loader (audio clips averaged, :610-618)

Here: features come from the towers' sm_100a kernels, similarities from the tcgen05 GEMM (bf16-rounded operands, fp32
accumulate / output), rankings from vl_topk_rows, mAP from vl_average_precision; counters stay on the device until the end.
Loaders are any iterable of dict batches with the reference's keys."""
from __future__ import annotations

import collections
from typing import Callable, Dict, Iterable, Optional, Sequence

import torch

from open_clip.metrics import MAP, Accuracy, Recall
from open_clip.zero_shot_classifier import build_zero_shot_classifier, class_text_features
from vitlens_b200 import ops as _ops


def acc(output, target, topk=(1,)):
    """zero_shot.py:45-60: ([acc@k in percent as 1-element tensors], correct [maxk, B] bool)."""
    with torch.no_grad():
        maxk = max(topk)
        pred = _ops.topk_rows(output.float().contiguous(), min(maxk, output.size(1))).long().t()  # [maxk, B]
        correct = pred.eq(target.reshape(1, -1).expand_as(pred))
        res = [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)) for k in topk]
        return res, correct


def accuracy(output, target, topk=(1,)):
    """zero_shot.py:36-42: number of correct samples per k (floats)."""
    _, correct = acc(output, target, topk)
    return [float(correct[:k].reshape(-1).float().sum()) for k in topk]


def _encode_clips(encode: Callable, x: torch.Tensor) -> torch.Tensor:
    """Audio arrives as [B, n_clip, T, F]: encode every clip, average the clip features (zero_shot.py:610-618)."""
    if x.ndim == 4:
        b, n = x.shape[:2]
        f = encode(x.reshape((b * n,) + tuple(x.shape[2:])))
        return f.reshape(b, n, -1).mean(dim=1)
    return encode(x)


def _unit(f):
    return f / f.norm(dim=-1, keepdim=True)


def run(model, classifier, dataloader, device="cuda"):
    """zero_shot.py:84-110: (top1, top5) fractions of an image-classification loader of (images, target) pairs."""
    top1 = torch.zeros((), device=device)
    top5 = torch.zeros((), device=device)
    n = 0
    with torch.no_grad():
        for images, target in dataloader:
            images, target = images.to(device), target.to(device)
            out = model(image=images)
            feats = out["image_features"] if isinstance(out, dict) else out[0]
            logits = 100.0 * _ops.similarity(feats.float(), gallery_t=classifier)
            _, correct = acc(logits, target, topk=(1, min(5, logits.size(1))))
            top1 += correct[:1].sum()
            top5 += correct[:5].sum()
            n += images.size(0)
    return float(top1) / n, float(top5) / n


def zero_shot_classification(model, tokenizer, loader: Iterable[Dict], labels: Sequence[str], templates, *, input_key: str, label_key: str = "label",
                             name_key: Optional[str] = "class_name", encode: Optional[Callable] = None, device="cuda"):
    """The shared body of test_zeroshot_3d_core / test_rgbd_cls_single / test_tactle_cls_single / test_eeg_cls_single
    (zero_shot.py:155-257 and siblings): top-1 / top-5 accuracy in percent plus per-class accuracies."""
    model.eval()
    encode = encode or (model.encode_visual if hasattr(model, "encode_visual") else model.encode_image)
    text_features = class_text_features(model, tokenizer, labels, templates, device)  # [C, E]
    n_cls = len(labels)
    seen = torch.zeros(n_cls, device=device, dtype=torch.long)
    hit1 = torch.zeros(n_cls, device=device, dtype=torch.long)
    hit5 = torch.zeros(n_cls, device=device, dtype=torch.long)
    with torch.no_grad():
        for batch in loader:
            x, target = batch[input_key].to(device), batch[label_key]
            target = (torch.as_tensor(target) if not torch.is_tensor(target) else target).long().to(device)
            feats = _unit(_encode_clips(encode, x).float())
            logits = _ops.similarity(feats, gallery=text_features)
            _, correct = acc(logits, target, topk=(1, min(5, n_cls)))
            seen += torch.bincount(target, minlength=n_cls)
            hit1 += torch.bincount(target[correct[0]], minlength=n_cls)
            hit5 += torch.bincount(target[correct[:5].any(0)], minlength=n_cls)
    total = int(seen.sum())
    per1 = collections.OrderedDict((labels[i], float(hit1[i]) / max(int(seen[i]), 1)) for i in range(n_cls) if int(seen[i]) > 0)
    per5 = collections.OrderedDict((labels[i], float(hit5[i]) / max(int(seen[i]), 1)) for i in range(n_cls) if int(seen[i]) > 0)
    return {"acc1": 100.0 * float(hit1.sum()) / total, "acc5": 100.0 * float(hit5.sum()) / total, "n": total,
            "top1_accuracy_per_class": per1, "top5_accuracy_per_class": per5}


def test_zeroshot_3d_core(test_loader, model, tokenizer, labels, templates, device="cuda"):
    """zero_shot.py:155-257 (labels / templates passed in instead of read from PC_META_DATA_DIR)."""
    r = zero_shot_classification(model, tokenizer, test_loader, labels, templates, input_key="pc", device=device)
    return dict(modelnet40={"acc1": r["acc1"], "acc5": r["acc5"]}, detail=r)


def _audio_text_features(model, tokenizer, labels, templates, device):
    return class_text_features(model, tokenizer, labels, templates, device)


def test_audio_single_map(testloader, model, tokenizer, labels, templates, dataset_name="Eval Audio mAP", device="cuda"):
    """zero_shot.py:572-638: multi-label audio tagging scored by mean average precision."""
    model.eval()
    metric = MAP()
    metric.initialize(device)
    text_features = _audio_text_features(model, tokenizer, labels, templates, device)
    with torch.no_grad():
        for batch in testloader:
            ids = torch.as_tensor(batch["id"]).to(device)
            feats = _unit(_encode_clips(model.encode_visual, batch["audio"].to(device)).float())
            metric.compute(ids, _ops.similarity(feats, gallery=text_features), batch["target"].to(device))
    stats = metric.merge_results()
    stats["acc1"] = stats["map"]
    return stats


def test_audio_single_cls(testloader, model, tokenizer, labels, templates, dataset_name="Eval Audio Cls", device="cuda"):
    """zero_shot.py:641-706."""
    model.eval()
    metric = Accuracy()
    metric.initialize(device)
    text_features = _audio_text_features(model, tokenizer, labels, templates, device)
    with torch.no_grad():
        for batch in testloader:
            ids = torch.as_tensor(batch["id"]).to(device)
            targets = torch.as_tensor(batch["label"]).long().to(device)
            feats = _unit(_encode_clips(model.encode_visual, batch["audio"].to(device)).float())
            metric.compute(ids, _ops.similarity(feats, gallery=text_features), targets)
    stats = metric.merge_results()
    stats["acc1"] = stats["accuracy"]
    return stats


def test_audio_single_ret(testloader, model, tokenizer, text_ids, texts, dataset_name="Eval Audio Ret", device="cuda"):
    """zero_shot.py:709-788: audio <-> caption retrieval, recall@1/5/10 both ways.  (As in the reference the caption features
    enter the similarity un-normalised, :745-752.)"""
    model.eval()
    metric = Recall()
    with torch.no_grad():
        text_ids = torch.as_tensor(text_ids).to(device)
        chunks = []
        for i in range(0, len(texts), 50):
            chunks.append(model.encode_text(tokenizer(list(texts[i:i + 50])).to(device)))
        metric.initialize(text_ids=text_ids, text_logits=torch.cat(chunks, dim=0).float())
        for batch in testloader:
            ids = torch.as_tensor(batch["uniq_id"]).to(device)
            metric.compute(ids, _unit(_encode_clips(model.encode_visual, batch["audio"].to(device)).float()))
        stats = metric.merge_results()
    for key in list(stats.keys()):
        if key.startswith("img"):
            stats[key.replace("img", "audio")] = stats.pop(key)
    stats["acc1"] = (stats["txt_r1"] + stats["txt_r5"] + stats["txt_r10"] + stats["audio_r1"] + stats["audio_r5"] + stats["audio_r10"]) / 600.0
    return stats
