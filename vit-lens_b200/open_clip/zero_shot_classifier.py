"""Zero-shot classifier weights from text templates (reference open_clip/zero_shot_classifier.py:27-88): for every class, encode
all templates, L2-normalise each, average, L2-normalise the mean.  The text tower runs on the sm_100a kernels; the per-class
reduction is one vl_template_mean launch per batch of classes (no [classes, templates, E] intermediate on the host)."""
from __future__ import annotations

from itertools import islice
from typing import Callable, Optional, Sequence, Union

import torch

from vitlens_b200 import ops as _ops


def batched(iterable, n):
    it = iter(iterable)
    while True:
        batch = list(islice(it, n))
        if not batch:
            break
        yield batch


def class_text_features(model, tokenizer, classnames: Sequence[str], templates: Sequence[Union[Callable, str]], device="cuda",
                        num_classes_per_batch: Optional[int] = 10) -> torch.Tensor:
    """[num_classes, E] fp32 rows (the layout test_zeroshot_3d_core / the audio tests stack, training/zero_shot.py:175-190)."""
    assert isinstance(templates, Sequence) and len(templates) > 0
    assert isinstance(classnames, Sequence) and len(classnames) > 0
    use_format = isinstance(templates[0], str)
    out = []
    with torch.no_grad():
        for batch in batched(classnames, num_classes_per_batch or len(classnames)):
            texts = [t.format(c) if use_format else t(c) for c in batch for t in templates]
            tokens = tokenizer(texts).to(device)
            if tokens.dim() < 2:
                tokens = tokens[None, ...]
            emb = model.encode_text(tokens).float()  # [len(batch) * T, E]
            out.append(_ops.template_mean(emb, len(templates)))
    return torch.cat(out, dim=0)


def build_zero_shot_classifier(model, tokenizer, classnames: Sequence[str], templates: Sequence[Union[Callable, str]],
                               num_classes_per_batch: Optional[int] = 10, device: Union[str, torch.device] = "cuda", use_tqdm: bool = False):
    """-> [E, num_classes] fp32 (`logits = 100. * image_features @ classifier`, training/zero_shot.py:101)."""
    return class_text_features(model, tokenizer, classnames, templates, device, num_classes_per_batch).t().contiguous()


def build_zero_shot_classifier_legacy(model, tokenizer, classnames, templates, device="cuda", use_tqdm=False):
    """One class per forward (zero_shot_classifier.py:91-130): same result."""
    return build_zero_shot_classifier(model, tokenizer, classnames, templates, num_classes_per_batch=1, device=device)
