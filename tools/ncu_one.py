"""Launch one kernel shape a few times (for `ncu --set full -k regex:... -s 3 -c 2`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L  # noqa: E402

dev = "cuda"
what = sys.argv[1]
T = 65792
torch.manual_seed(0)
if what in ("gelu", "linear", "residual", "gelu_bwd"):
    N, K = (4096, 1024) if what in ("gelu", "gelu_bwd") else ((3072, 1024) if what == "linear" else (1024, 4096))
    A = torch.randn(T, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    aux = torch.randn(T, N, device=dev).bfloat16()
    D = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
    epi = {"gelu": L.EPI_GELU, "linear": L.EPI_LINEAR, "residual": L.EPI_RESIDUAL, "gelu_bwd": L.EPI_GELU_BWD}[what]
    for _ in range(6):
        L.gemm(A, B, D, M=T, N=N, K=K, lda=K, ldb=K, ldd=N, epilogue=epi, bias=None if what == "gelu_bwd" else bias,
               aux_in=aux if what in ("residual", "gelu_bwd") else None, aux_out=aux if what == "gelu" else None, ldaux=N)
elif what == "wgrad":
    M, N = 4096, 1024
    A = torch.randn(T, M, device=dev).bfloat16()
    B = torch.randn(T, N, device=dev).bfloat16()
    D = torch.zeros(M, N, device=dev)
    for _ in range(6):
        L.gemm(A, B, D, M=M, N=N, K=T, lda=M, ldb=N, ldd=N, a_mn=True, b_mn=True, accumulate=True, split_k=4)
elif what in ("attn_fwd", "attn_bwd"):
    Bn, H, Nn = 256, 16, 257
    Dm = H * 64
    qkv = torch.randn(Bn * Nn, 3 * Dm, device=dev).bfloat16()
    q, k, v = qkv[:, :Dm], qkv[:, Dm:2 * Dm], qkv[:, 2 * Dm:]
    o = torch.zeros(Bn * Nn, Dm, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(Bn, H, Nn, device=dev)
    do = torch.randn(Bn * Nn, Dm, device=dev).bfloat16()
    dqkv = torch.zeros_like(qkv)
    for _ in range(6):
        L.attention_fwd(q, k, v, o, lse, B=Bn, H=H, nq=Nn, nk=Nn, ldq=3 * Dm, ldk=3 * Dm, ldv=3 * Dm, ldo=Dm, scale=0.125)
        if what == "attn_bwd":
            L.attention_bwd(q, k, v, o, do, lse, dqkv[:, :Dm], dqkv[:, Dm:2 * Dm], dqkv[:, 2 * Dm:], B=Bn, H=H, nq=Nn, nk=Nn, ldq=3 * Dm,
                            ldk=3 * Dm, ldv=3 * Dm, ldo=Dm, lddo=Dm, lddq=3 * Dm, lddk=3 * Dm, lddv=3 * Dm, scale=0.125)
torch.cuda.synchronize()
print("ok", what)
