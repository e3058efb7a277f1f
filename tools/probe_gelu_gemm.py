"""Timing of the two activation-epilogue GEMMs of a ViT-L/14 block at batch 256 (fc + GELU forward, c_proj dgrad x GELU')."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import ops
T, D = 256 * 257, 1024
x = torch.randn(T, D, device="cuda").bfloat16()
wfc = (torch.randn(4 * D, D, device="cuda") * 0.02).bfloat16()
bfc = torch.randn(4 * D, device="cuda") * 0.1
dy = torch.randn(T, D, device="cuda").bfloat16()
wproj = (torch.randn(D, 4 * D, device="cuda") * 0.02).bfloat16()
u = torch.randn(T, 4 * D, device="cuda").bfloat16()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
fl = 2.0 * T * D * 4 * D
for name, fn in (("fc+GELU fwd (two outputs)", lambda: ops.gemm(x, wfc, bias=bfc, epilogue=ops.EPI_GELU, want_aux_out=True)),
                 ("c_proj dgrad x GELU'", lambda: ops.gemm(dy, wproj, b_t=True, epilogue=ops.EPI_GELU_BWD, aux_in=u)),
                 ("fc linear (no act)", lambda: ops.gemm(x, wfc, bias=bfc))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"[gelu-gemm] {name}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
