"""clock64 timeline of attn_fwd2_kernel (first CTAs, steady-state items): producer / MMA issuer 0 / softmax warp 4 / tail warp 12."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L
h = L.load()
B, H, N = 256, 16, 257
D = H * 64
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
o = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B, H, N, device="cuda")
def fwd():
    L.attention_fwd(q, k, v, o, lse, B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D, ldv=3 * D, ldo=D, scale=0.125)
for _ in range(3):
    fwd()
buf = torch.zeros(4 * 4 * 64, device="cuda", dtype=torch.int64)
h.vl_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
fwd()
torch.cuda.synchronize()
h.vl_debug_buffer(ctypes.c_void_p(0))
t = buf.cpu().reshape(4, 4, 64)
names = {0: ["stage free"], 1: ["KV landed", "S issued", "waiting P", "P ready/issue PV"],
         2: ["block start", "S ready", "scores in regs", "exps done", "P handed over", "(rows stored)"], 3: ["KV landed", "tail done"]}
roles = ["producer", "mma0", "softmax w4", "tail w12"]
for c in range(2):
    base = min(int(x) for x in t[c].flatten() if int(x) > 0)
    for r in range(4):
        row = [int(x) for x in t[c][r] if int(x) > 0]
        print(f"--- CTA {c} {roles[r]} ({len(row)} stamps)")
        prev = row[0] if row else 0
        out = []
        for i, val in enumerate(row[:48]):
            out.append(f"{val - base}(+{val - prev})")
            prev = val
        print("   " + " ".join(out))
