"""A few launches of the LayerNorm kernels at the ViT-L/14 bench shape (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import ops
T, D = 256 * 257, 1024
x = torch.randn(T, D, device="cuda").bfloat16()
dy = torch.randn(T, D, device="cuda").bfloat16()
dres = torch.randn(T, D, device="cuda").bfloat16()
w = torch.randn(D, device="cuda")
b = torch.randn(D, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, byts in (("ln_fwd", lambda: ops.layernorm_fwd(x, w, b), 2 * T * D * 2),
                       ("ln_bwd(+dres, wgrad)", lambda: ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_wgrad=True), 4 * T * D * 2),
                       ("ln_bwd(frozen)", lambda: ops.layernorm_bwd(dy, x, w, mean, rstd, want_wgrad=False), 3 * T * D * 2)):
    if name == "ln_fwd":
        y, mean, rstd = fn()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"[ln] {name}: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s", flush=True)
