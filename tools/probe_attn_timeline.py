"""clock64 timeline of attn_bwd2_kernel's two compute groups (first CTAs) -- where does a CTA's time go?"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L
h = L.load()
B, H, N = 64, 16, 257
D = H * 64
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
o = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B, H, N, device="cuda")
do = torch.randn(B * N, D, device="cuda").bfloat16()
dqkv = torch.zeros_like(qkv)
L.attention_fwd(q, k, v, o, lse, B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D, ldv=3 * D, ldo=D, scale=0.125)
def bwd():
    L.attention_bwd(q, k, v, o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D,
                    ldv=3 * D, ldo=D, lddo=D, lddq=3 * D, lddk=3 * D, lddv=3 * D, scale=0.125)
for _ in range(3):
    bwd()
buf = torch.zeros(8 * 64, device="cuda", dtype=torch.int64)
h.vl_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
bwd()
torch.cuda.synchronize()
h.vl_debug_buffer(ctypes.c_void_p(0))
t = buf.cpu().reshape(8, 64)
# per group (32 slots each): start, prologue done, then per key block: [S ready, P written, dP ready, dS written] (when the
# group has a pair), dKV complete, block done; then dQ complete, tile outputs written, end.
seq = ["start", "prologue done"]
for j in range(2):
    seq += [f"blk{j} S ready", f"blk{j} P written", f"blk{j} dP ready", f"blk{j} dS written", f"blk{j} dKV complete", f"blk{j} tail dots",
            f"blk{j} dKV rows stored", f"blk{j} done (mat-vec)"]
seq += ["dQ complete", "tail-key dots", "dQ rows stored", "tail mat-vecs done", "end"]
for c in range(2):
    for g in range(2):
        row = t[c][g * 32:(g + 1) * 32]
        base = int(t[c][0])
        print(f"--- CTA {c} group {g}")
        prev = int(row[0])
        for n, nm in enumerate(seq):
            val = int(row[n])
            if val == 0:
                break
            print(f"   {nm:28s} +{val - prev:7d}   (t={val - base})")
            prev = val
