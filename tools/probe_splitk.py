"""Split-K sweep for the ViT-L/14 weight-gradient GEMMs (MN-major operands, fp32 atomics), CTA-pair kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L  # noqa: E402

dev = "cuda"
T = 65792
for name, M, N in (("fc", 4096, 1024), ("proj", 1024, 4096), ("qkv", 3072, 1024), ("out", 1024, 1024)):
    A = torch.randn(T, M, device=dev).bfloat16()
    B = torch.randn(T, N, device=dev).bfloat16()
    D = torch.zeros(M, N, device=dev)
    res = []
    for s in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16):
        def call():
            L.gemm(A, B, D, M=M, N=N, K=T, lda=M, ldb=N, ldd=N, a_mn=True, b_mn=True, accumulate=True, split_k=s)
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
        res.append((s, ms, 2.0 * M * N * T / ms / 1e9))
    print(name, " ".join(f"s{s}:{ms:.3f}ms/{tf:.0f}" for s, ms, tf in res), flush=True)
