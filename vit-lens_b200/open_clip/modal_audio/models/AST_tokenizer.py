"""AST patch embed (reference modal_audio/models/AST_tokenizer.py:7-57): overlapping 2-D conv over the
(mel x time) spectrogram, expressed as patch-gather + tcgen05 GEMM; the unsqueeze/transpose of the
reference (:46-47) is folded into the gather strides."""
import torch
import torch.nn as nn

from vitlens_b200 import engine as E

from ...transformer import TokenMat
from ...util.Sample import Sample


class AST_tokenizer(nn.Module):
    def __init__(self, fstride, tstride, input_fdim, input_tdim, patch_size=(16, 16), width=768):
        super().__init__()
        if isinstance(patch_size, int):
            patch_size = (patch_size, patch_size)
        self.fstride, self.tstride = fstride, tstride
        self.input_fdim, self.input_tdim = input_fdim, input_tdim
        self.width = width
        self.patch_size = tuple(patch_size)
        self.conv1 = nn.Conv2d(in_channels=1, out_channels=width, kernel_size=patch_size, stride=(fstride, tstride), bias=False)
        self.fdim, self.tdim = self.get_tokenized_dim()
        self.num_patches = self.fdim * self.tdim
        scale = width ** -0.5
        self.pos_emb = nn.Parameter(scale * torch.randn(self.num_patches, width))

    def get_tokenized_dim(self):
        kh, kw = self.patch_size
        return (self.input_fdim - kh) // self.fstride + 1, (self.input_tdim - kw) // self.tstride + 1

    def forward(self, x):
        # x: [B, time, mel]; the conv runs over the image [B, 1, mel, time]
        B, T, F = x.shape
        kh, kw = self.patch_size
        oh, ow = (F - kh) // self.fstride + 1, (T - kw) // self.tstride + 1
        x = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        geom = dict(B=B, C=1, OH=oh, OW=ow, kh=kh, kw=kw, stride_h=self.fstride, stride_w=self.tstride,
                    sb=x.stride(0), sc=0, sh=x.stride(2), sw=x.stride(1))
        tok = E.PatchEmbedFn.apply(x, self.conv1.weight, geom)
        return Sample({"x": TokenMat(tok, B, oh * ow), "pos": self.pos_emb})
