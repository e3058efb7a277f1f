"""A few launches of vl_attention_fwd at the ViT-L/14 bench shape (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L
B, H, N = 256, 16, 257
D = H * 64
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
o = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B, H, N, device="cuda")
for _ in range(4):
    L.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, lse, B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D, ldv=3 * D, ldo=D, scale=0.125)
torch.cuda.synchronize()
