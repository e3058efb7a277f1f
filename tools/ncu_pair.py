import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L
L.debug_set(8, 2)
M = N = K = 8192
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    L.gemm(A, B, D, M=M, N=N, K=K, lda=K, ldb=K, ldd=N)
torch.cuda.synchronize(); print("ok")
