"""ViT-L/14 weight-gradient GEMMs (MN-major operands, fp32 out): plain vs with the fused bias gradient (row sums of A)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import ops  # noqa: E402

T = 65792
for name, M, N in (("fc", 4096, 1024), ("proj", 1024, 4096), ("qkv", 3072, 1024), ("out", 1024, 1024)):
    dy = torch.randn(T, M, device="cuda").bfloat16()
    x = torch.randn(T, N, device="cuda").bfloat16()
    res = []
    for rs in (False, True):
        def call():
            return ops.gemm(dy, x, a_t=True, b_t=True, out_dtype=torch.float32, accumulate=True, want_rowsum=rs)
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        res.append(f"{'with bias grad' if rs else 'plain':15s} {ms:.3f} ms {2.0 * M * N * T / ms / 1e9:6.0f} TFLOP/s (split {ops.pick_split_k(M, N, T)})")
    print(f"[wgrad {name:4s}] " + " | ".join(res), flush=True)
