"""Does the [T, 3D] packed layout (128-byte head slices at a 6 KB stride) limit attention?  Same FLOPs / bytes, three layouts."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L
B, H, N = 256, 16, 257
D = H * 64
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"[layout] {name}: {ms:.3f} ms  {4 * B * N * D * 2 / ms / 1e6:.0f} GB/s", flush=True)
for knob in (0, 1):
    L.debug_set(13, knob)
    tag = "one-tile kernel" if knob else "fwd2"
    qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
    o = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device="cuda")
    timeit(f"{tag} packed [T,3D]", lambda: L.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, lse, B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D, ldv=3 * D, ldo=D, scale=0.125))
    q, k, v = (torch.randn(B * N, D, device="cuda").bfloat16() for _ in range(3))
    timeit(f"{tag} separate [T,D] x3", lambda: L.attention_fwd(q, k, v, o, lse, B=B, H=H, nq=N, nk=N, ldq=D, ldk=D, ldv=D, ldo=D, scale=0.125))
    # head-major: every (b, h) tile contiguous -> expressed as B*H "samples" of one head
    q, k, v = (torch.randn(B * H * N, 64, device="cuda").bfloat16() for _ in range(3))
    o2 = torch.zeros(B * H * N, 64, device="cuda", dtype=torch.bfloat16)
    lse2 = torch.zeros(B * H, 1, N, device="cuda")
    timeit(f"{tag} head-major contiguous tiles", lambda: L.attention_fwd(q, k, v, o2, lse2, B=B * H, H=1, nq=N, nk=N, ldq=64, ldk=64, ldv=64, ldo=64, scale=0.125))
L.debug_set(13, 0)
