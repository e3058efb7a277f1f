"""clock64 timeline of attn_bwd_kernel's compute warps (first 8 CTAs) -- where does a CTA's time go?"""
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
names = ["start", "prologue done"]
for pr in range(4):
    names += [f"p{pr} S ready", f"p{pr} prev retired", f"p{pr} P written", f"p{pr} dP ready", f"p{pr} dS written"]
# insertion points of dK/dV stamps: after pairs of each key block (2 pairs per block)
seq = ["start", "prologue done"]
for j in range(2):
    for i in range(2):
        pr = j * 2 + i
        seq += [f"p{pr} S ready", f"p{pr} prev retired", f"p{pr} P written", f"p{pr} dP ready", f"p{pr} dS written"]
    seq += [f"blk{j} dKV complete", f"blk{j} dKV written"]
seq += ["dQ complete"]
for c in range(3):
    row = t[c]
    base = int(row[0])
    print(f"--- CTA {c}: total to compute-done {int(row[63]) - base} cycles")
    prev = base
    for n, nm in enumerate(seq):
        val = int(row[n])
        if val == 0:
            break
        print(f"   {nm:22s} +{val - prev:7d}   (t={val - base})")
        prev = val
