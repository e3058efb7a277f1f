"""clock64 timeline of attn_bwd3_kernel's two compute groups (CTAs of a later wave) -- where does a CTA's time go?
Stamps (thread 0 of each group): start, prologue done, then per key block and per query tile [S^T/dP^T ready, unit handed over],
[dKV complete, dKV rows stored], then dQ complete, end."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L
h = L.load()
B, H, N = 64, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 257
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
buf = torch.zeros(17 * 64, device="cuda", dtype=torch.int64)
h.vl_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
bwd()
torch.cuda.synchronize()
h.vl_debug_buffer(ctypes.c_void_p(0))
t = buf.cpu().reshape(17, 64)
nqt = min(2, ((N - (N % 128 if 0 < N % 128 <= 4 and N > 128 else 0)) + 127) // 128)
tail = N % 128 if 0 < N % 128 <= 4 and N > 128 else 0
nkb = (N - tail + 127) // 128
UPT = 1  # units per group and tile (kUPT in attention_bwd3.cu)
seq = ["start", "prologue done"]
for j in range(nkb):
    for i in range(nqt):
        seq += [f"blk{j} tile{i} unit{k} handed over" for k in range(UPT)]
    seq += [f"blk{j} dKV complete", f"blk{j} dKV rows stored"]
seq += ["dQ complete", "end"]
for c in range(2):
    for g in range(2):
        row = t[c][g * 32:(g + 1) * 32]
        base = int(t[c][0])
        print(f"--- CTA {c} group {g}")
        prev = int(row[0])
        names = seq if g == 0 or nqt > 1 else seq[:-2] + ["end"]
        for n, nm in enumerate(names):
            val = int(row[n])
            if val == 0:
                break
            print(f"   {nm:32s} +{val - prev:7d}   (t={val - base})")
            prev = val

# issuer thread: start, first four S units issued, then per unit [hand-over received, MMAs + refill issued]
for c in range(1):
    row = t[8 + c]
    base = int(t[c][0])
    print(f"--- CTA {c} issuer (t relative to group 0's start)")
    names = ["start", "Q/dO landed, 4 S units issued"]
    nun = 2 * UPT * nqt + (1 if tail else 0)
    for j in range(nkb):
        for u in range(nun):
            nm = "tail" if u == 2 * UPT * nqt else f"t{u // (2 * UPT)}h{(u >> 1) % UPT}g{u & 1}"
            names += [f"blk{j} {nm} received", f"blk{j} {nm} issued"]
    prev = int(row[0])
    for n, nm in enumerate(names):
        if n >= 63 or int(row[n]) == 0:
            break
        val = int(row[n])
        print(f"   {nm:32s} +{val - prev:7d}   (t={val - base})")
        prev = val

for c in range(4):
    row = t[16][c * 8:(c + 1) * 8]
    base = int(row[0])
    print(f"--- CTA {c} prologue: entry=0 (group start stamp at {int(t[c][0]) - base}), producer starts {int(row[1]) - base}, loads issued {int(row[2]) - base}, "
          f"Q0 landed {int(row[3]) - base}, dO0/O0 landed {int(row[4]) - base}, K0/V0 landed {int(row[5]) - base}")
