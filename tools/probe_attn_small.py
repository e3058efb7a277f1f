"""Single-tile sequences (the 128-latent Lens + cls = 129 tokens of the audio / depth / point recipes): one-tile-per-CTA forward
kernel vs the persistent kernel with one idle softmax group (debug knob 17), parity and timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
import probe_attn as P  # noqa: E402  (runs nothing on import)
from vitlens_b200 import lib as L  # noqa: E402

L.debug_set(17, 1)
ok = True
for args in ((1, 1, 128, 128), (1, 1, 64, 64), (2, 2, 17, 17), (4, 12, 50, 50), (3, 2, 129, 129), (2, 2, 100, 300), (2, 1, 128, 600)):
    ok &= P.run(*args, packed=args[2] == args[3])
L.debug_set(17, 0)
print("parity with knob 17:", "OK" if ok else "BAD", flush=True)
B, H = 512, 16
D = H * 64
for N in (129, 128, 197):
    qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    o = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device="cuda")
    for knob in (0, 1):
        L.debug_set(17, knob)
        f = lambda: L.attention_fwd(q, k, v, o, lse, B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D, ldv=3 * D, ldo=D, scale=0.125)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[time] attn fwd B={B} H={H} N={N} knob17={knob}: {ms:.3f} ms  {4 * B * N * D * 2 / ms / 1e6:.0f} GB/s", flush=True)
    L.debug_set(17, 0)
