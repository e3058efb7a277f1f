"""GEMM main-loop / epilogue attribution: time three ViT-L shapes under the bring-up knobs (debug key 9) for both the
single-CTA and the CTA-pair kernel.  Results with knobs set are NOT correct GEMMs - timing attribution only."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L  # noqa: E402

dev = "cuda"
T = 65792
SHAPES = [("8k^3", 8192, 8192, 8192, L.EPI_LINEAR), ("qkv", T, 3072, 1024, L.EPI_LINEAR), ("fc+gelu", T, 4096, 1024, L.EPI_GELU),
          ("proj+res", T, 1024, 4096, L.EPI_RESIDUAL), ("out+res", T, 1024, 1024, L.EPI_RESIDUAL),
          ("dgrad gelu'", T, 4096, 1024, L.EPI_GELU_BWD)]


def timeit(M, N, K, epi, iters=8):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    aux = torch.randn(M, N, device=dev).bfloat16() if epi in (L.EPI_RESIDUAL, L.EPI_GELU_BWD) else None
    aux_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi == L.EPI_GELU else None
    D = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)

    def call():
        L.gemm(A, B, D, M=M, N=N, K=K, lda=K, ldb=K, ldd=N, epilogue=epi, bias=(None if epi == L.EPI_GELU_BWD else bias), aux_in=aux, aux_out=aux_out, ldaux=N)

    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9


if __name__ == "__main__":
    for kern, kname in ((2, "pair"), (1, "single")):
        L.debug_set(8, kern)
        for ew in (8, 16, 0):
            L.debug_set(10, 1 if ew == 0 else 0)
            L.debug_set(11, ew)
            tname = f"tma-epi/{ew}w" if ew else "direct/16w "
            for dbg, dname in ((0, "normal"), (2, "epilogue without global traffic"), (1, "no epilogue")):
                L.debug_set(9, dbg)
                for name, M, N, K, epi in SHAPES:
                    ms, tf = timeit(M, N, K, epi)
                    print(f"[{kname:6s} {tname}] {dname:32s} {name:11s} {ms:.3f} ms {tf:6.0f} TFLOP/s", flush=True)
    for k in (8, 9, 10, 11):
        L.debug_set(k, 0)
