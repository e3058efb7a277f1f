"""Minimal tensor-only image preprocessing (reference open_clip/transform.py): resize + centre crop +
normalise on torch tensors [C,H,W] in [0,1].  File decoding / PIL augmentation is host-side I/O and out of scope."""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .constants import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD


@dataclass
class AugmentationCfg:
    scale: Tuple[float, float] = (0.9, 1.0)


class _TensorTransform:
    def __init__(self, size, mean, std):
        self.size = size if isinstance(size, (tuple, list)) else (size, size)
        self.mean = torch.tensor(mean or OPENAI_DATASET_MEAN).view(-1, 1, 1)
        self.std = torch.tensor(std or OPENAI_DATASET_STD).view(-1, 1, 1)

    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(img):
            raise TypeError("pass a float tensor [C,H,W] in [0,1] (image decoding is outside this package)")
        c, h, w = img.shape
        s = max(self.size[0] / h, self.size[1] / w)
        nh, nw = max(self.size[0], round(h * s)), max(self.size[1], round(w * s))
        img = F.interpolate(img[None].float(), size=(nh, nw), mode="bicubic", align_corners=False)[0]
        t, l = (nh - self.size[0]) // 2, (nw - self.size[1]) // 2
        img = img[:, t:t + self.size[0], l:l + self.size[1]]
        return (img - self.mean) / self.std


def image_transform(image_size, is_train: bool, mean: Optional[Tuple[float, ...]] = None, std: Optional[Tuple[float, ...]] = None,
                    resize_longest_max: bool = False, fill_color: int = 0, aug_cfg=None):
    return _TensorTransform(image_size, mean, std)
