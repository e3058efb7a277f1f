"""EEG 1-D patch embed (reference modal_eeg/models/EEG_tokenizer.py:7-42): Conv1d(in_chans, width, k=window, stride)
over time == a [C x window] patch gather + GEMM."""
import torch
import torch.nn as nn

from vitlens_b200 import engine as E

from ...transformer import TokenMat
from ...util.Sample import Sample


class PatchEmbed1D(nn.Module):
    def __init__(self, time_len=512, in_chans=128, window_size=4, stride=2, width=768):
        super().__init__()
        self.time_len, self.in_chans, self.window_size, self.stride, self.width = time_len, in_chans, window_size, stride, width
        self.num_patches = (time_len - window_size) // stride + 1
        self.proj = nn.Conv1d(in_chans, width, kernel_size=window_size, stride=stride)
        scale = width ** -0.5
        self.pos_emb = nn.Parameter(scale * torch.randn(self.num_patches, width))

    def forward(self, x, **kwargs):
        # x: [B, C, T] -> tokens [B, L, width]; Conv1d == Conv2d over [C=chans, H=1, W=T] with a [1 x window] kernel
        B, C, T = x.shape
        L = (T - self.window_size) // self.stride + 1
        x = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        geom = dict(B=B, C=C, OH=1, OW=L, kh=1, kw=self.window_size, stride_h=1, stride_w=self.stride,
                    sb=x.stride(0), sc=x.stride(1), sh=0, sw=x.stride(2))
        tok = E.PatchEmbedBiasFn.apply(x, self.proj.weight, self.proj.bias, geom)
        return Sample({"x": TokenMat(tok, B, L), "pos": self.pos_emb})
