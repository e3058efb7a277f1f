"""GPU bring-up probe for vl_attention_fwd / vl_attention_bwd against a plain fp32 torch reference.
    python tools/probe_attn.py fwd|bwd|time"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def ref_attn(q, k, v, scale, causal):
    # q [B,nq,H,64] etc. fp32
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if causal:
        n = q.shape[1]
        s = s + torch.full((n, n), float("-inf"), device=q.device).triu_(1)
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v)
    return o, torch.logsumexp(s, dim=-1)


def make(B, H, nq, nk, packed):
    D = H * 64
    if packed:
        qkv = (torch.randn(B * nq, 3 * D, device=dev) * 1.0).bfloat16()
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        ld = 3 * D
        return qkv, q, k, v, ld, ld, ld
    qb = torch.randn(B * nq, D, device=dev).bfloat16()
    kvb = torch.randn(B * nk, 2 * D, device=dev).bfloat16()
    return (qb, kvb), qb, kvb[:, :D], kvb[:, D:], D, 2 * D, 2 * D


def check(name, got, ref, tol=3e-2):
    err = (got.float() - ref).abs().max().item()
    sc = ref.abs().max().item()
    ok = err <= tol * sc + 1e-3 and math.isfinite(err)
    print(f"   [{'OK ' if ok else 'BAD'}] {name}: err={err:.3e} scale={sc:.3e}", flush=True)
    if not ok:
        bad = (got.float() - ref).abs() > tol * sc + 1e-3
        idx = bad.nonzero()
        print(f"        bad frac={bad.float().mean().item():.4f} first={idx[:6].tolist()} nan={torch.isnan(got.float()).sum().item()}", flush=True)
    return ok


def run(B, H, nq, nk, causal=False, packed=True, bwd=False, qmul=1.0):
    D = H * 64
    scale = 64 ** -0.5
    hold, q, k, v, ldq, ldk, ldv = make(B, H, nq, nk, packed)
    if qmul != 1.0:  # large score spread: exercises the lazily moved softmax offset (O rescale in TMEM)
        q.mul_(qmul)
    o = torch.zeros(B * nq, D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, nq, device=dev)
    L.attention_fwd(q, k, v, o, lse, B=B, H=H, nq=nq, nk=nk, ldq=ldq, ldk=ldk, ldv=ldv, ldo=D, scale=scale, causal=causal)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, nq, H, 64).requires_grad_(True)
    kf = k.float().reshape(B, nk, H, 64).requires_grad_(True)
    vf = v.float().reshape(B, nk, H, 64).requires_grad_(True)
    oref, lref = ref_attn(qf, kf, vf, scale, causal)
    print(f"attn B={B} H={H} nq={nq} nk={nk} causal={int(causal)} packed={int(packed)}", flush=True)
    ok = check("O", o.reshape(B, nq, H, 64), oref.detach())
    ok &= check("LSE", lse, lref.detach(), tol=1e-2)
    if bwd:
        do = torch.randn(B * nq, D, device=dev).bfloat16()
        if packed:
            dqkv = torch.zeros(B * nq, 3 * D, device=dev, dtype=torch.bfloat16)
            dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
            lq = lk = lv = 3 * D
        else:
            dq = torch.zeros(B * nq, D, device=dev, dtype=torch.bfloat16)
            dkv = torch.zeros(B * nk, 2 * D, device=dev, dtype=torch.bfloat16)
            dk, dv = dkv[:, :D], dkv[:, D:]
            lq, lk, lv = D, 2 * D, 2 * D
        L.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B=B, H=H, nq=nq, nk=nk, ldq=ldq, ldk=ldk, ldv=ldv, ldo=D, lddo=D,
                        lddq=lq, lddk=lk, lddv=lv, scale=scale, causal=causal)
        torch.cuda.synchronize()
        gq, gk, gv = torch.autograd.grad(oref, (qf, kf, vf), do.float().reshape(B, nq, H, 64))
        ok &= check("dQ", dq.reshape(B, nq, H, 64), gq)
        ok &= check("dK", dk.reshape(B, nk, H, 64), gk)
        ok &= check("dV", dv.reshape(B, nk, H, 64), gv)
    return ok


def timing():
    B, H, N = 256, 16, 257
    D = H * 64
    scale = 64 ** -0.5
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    o = torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device=dev)
    do = torch.randn(B * N, D, device=dev).bfloat16()
    dqkv = torch.zeros_like(qkv)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def f():
        L.attention_fwd(q, k, v, o, lse, B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D, ldv=3 * D, ldo=D, scale=scale)

    def g():
        L.attention_bwd(q, k, v, o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], B=B, H=H, nq=N, nk=N, ldq=3 * D, ldk=3 * D,
                        ldv=3 * D, ldo=D, lddo=D, lddq=3 * D, lddk=3 * D, lddv=3 * D, scale=scale)

    def f_old():
        L.debug_set(13, 1)
        f()
        L.debug_set(13, 0)

    def g_old():
        L.debug_set(12, 2)
        g()
        L.debug_set(12, 0)

    def g_knob(c):
        def fn():
            L.debug_set(15, c)
            g()
            L.debug_set(15, 0)
        return fn

    extra = [(f"bwd(L2 prefetch distance {c})", g_knob(c), 8 * B * N * D * 2, 10.0 * B * H * N * N * 64) for c in (-1, 74, 296)]
    for name, fn, byts, flops in extra + [("fwd", f, 4 * B * N * D * 2, 4.0 * B * H * N * N * 64), ("fwd(one-tile kernel)", f_old, 4 * B * N * D * 2, 4.0 * B * H * N * N * 64),
                                  ("bwd", g, 8 * B * N * D * 2, 10.0 * B * H * N * N * 64),
                                  ("bwd(gen-2 kernel)", g_old, 8 * B * N * D * 2, 10.0 * B * H * N * N * 64)]:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[time] attn {name} B={B} H={H} N={N}: {ms:.3f} ms  {byts / ms / 1e6:.0f} GB/s algorithmic  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
    # torch SDPA comparison
    qh = torch.randn(B, H, N, 64, device=dev).bfloat16().requires_grad_(True)
    kh = torch.randn_like(qh).requires_grad_(True)
    vh = torch.randn_like(qh).requires_grad_(True)
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        out = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
    e1.record()
    torch.cuda.synchronize()
    print(f"[time] torch SDPA fwd: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    gr = torch.randn_like(out)
    e0.record()
    for _ in range(5):
        out = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
        out.backward(gr)
    e1.record()
    torch.cuda.synchronize()
    print(f"[time] torch SDPA fwd+bwd: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    if mode == "time":
        timing()
    else:
        b = mode == "bwd"
        run(1, 1, 128, 128, bwd=b)
        run(1, 1, 64, 64, bwd=b)
        run(2, 2, 257, 257, bwd=b)
        run(3, 2, 77, 77, causal=True, bwd=b)
        run(2, 16, 256, 256, bwd=b)
        run(2, 1, 256, 600, packed=False, bwd=b)
        run(2, 2, 17, 17, bwd=b)
        run(2, 2, 300, 300, causal=True, bwd=b)
        run(4, 12, 50, 50, bwd=b)
        run(40, 16, 257, 257, bwd=b)
        run(2, 2, 513, 513, bwd=b)
        run(1, 1, 640, 132, packed=False, bwd=b)
        run(2, 1, 229, 229, bwd=b)
        run(2, 2, 260, 260, bwd=b)
        run(3, 4, 257, 257, bwd=b, qmul=8.0)
        run(2, 2, 384, 700, packed=False, bwd=b, qmul=6.0)
    print("done", flush=True)
