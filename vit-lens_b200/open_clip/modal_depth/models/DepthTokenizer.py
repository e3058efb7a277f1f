"""Depth patch embed (reference modal_depth/models/DepthTokenizer.py:7-60, input_patchnorm=False)."""
import torch
import torch.nn as nn

from vitlens_b200 import engine as E

from ...transformer import TokenMat
from ...util.Sample import Sample


class DepthTokenizer(nn.Module):
    def __init__(self, grid_size, patch_size, width, input_patchnorm):
        super().__init__()
        if input_patchnorm:
            raise NotImplementedError("input_patchnorm=True is broken in the reference (DepthTokenizer.py:17) and unused")
        self.grid_size = grid_size
        self.patch_size = patch_size
        self.width = width
        self.input_patchnorm = input_patchnorm
        self.patchnorm_pre_ln = nn.Identity()
        self.conv1 = nn.Conv2d(in_channels=1, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.pos_emb = nn.Parameter(scale * torch.randn(self.grid_size[0] * self.grid_size[1], width))

    def forward(self, x):
        B, C, H, W = x.shape
        kh, kw = self.patch_size
        gh, gw = H // kh, W // kw
        x = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        geom = dict(B=B, C=C, OH=gh, OW=gw, kh=kh, kw=kw, stride_h=kh, stride_w=kw, sb=x.stride(0), sc=x.stride(1), sh=x.stride(2), sw=x.stride(3))
        tok = E.PatchEmbedFn.apply(x, self.conv1.weight, geom)
        return Sample({"x": TokenMat(tok, B, gh * gw), "pos": self.pos_emb})
