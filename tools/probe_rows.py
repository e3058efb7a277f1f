"""GPU bring-up probe for the row kernels (LN, colsum, patchify, assemble, embed, l2norm, geglu, adamw)."""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-lens_b200"))
from vitlens_b200 import lib as L  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
BF = torch.bfloat16


def check(name, got, ref, tol=2e-2, atol=1e-3):
    err = (got.float() - ref.float()).abs().max().item()
    sc = ref.float().abs().max().item()
    ok = err <= tol * sc + atol and math.isfinite(err)
    print(f"[{'OK ' if ok else 'BAD'}] {name}: err={err:.3e} scale={sc:.3e}", flush=True)
    return ok


def ln(T, D, gather=False):
    x = (torch.randn(T, D, device=dev) * 2 + 0.5).to(BF)
    w = torch.randn(D, device=dev) * 0.1 + 1
    b = torch.randn(D, device=dev) * 0.1
    idx = None
    Tn = T
    if gather:
        idx = torch.randperm(T, device=dev)[: T // 3].contiguous()
        Tn = idx.numel()
    y = torch.empty(Tn, D, device=dev, dtype=BF)
    mean = torch.empty(Tn, device=dev)
    rstd = torch.empty(Tn, device=dev)
    L.layernorm_fwd(x, w, b, y, mean, rstd, T=Tn, D=D, ldx=D, ldy=D, row_index=idx)
    xs = x.float()[idx] if gather else x.float()
    xr = xs.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), wr, br, 1e-5)
    check(f"ln_fwd T={T} D={D} gather={int(gather)}", y, yr.detach())
    dy = torch.randn(Tn, D, device=dev).to(BF)
    yr.backward(dy.float())
    dx = torch.zeros(T, D, device=dev, dtype=BF)
    dw = torch.zeros(D, device=dev)
    db = torch.zeros(D, device=dev)
    dres = torch.randn(T, D, device=dev).to(BF)
    L.layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db, T=Tn, D=D, lddy=D, ldx=D, lddx=D, dres=dres, lddres=D, row_index=idx)
    ref_dx = xr.grad + (dres.float()[idx] if gather else dres.float())
    got_dx = dx[idx] if gather else dx
    check("  ln_bwd dx(+dres)", got_dx, ref_dx)
    check("  ln_bwd dw", dw, wr.grad, tol=2e-2, atol=0.05)
    check("  ln_bwd db", db, br.grad, tol=2e-2, atol=0.05)
    dx2 = torch.zeros(T, D, device=dev, dtype=BF)
    L.layernorm_bwd(dy, x, w, mean, rstd, dx2, None, None, T=Tn, D=D, lddy=D, ldx=D, lddx=D, row_index=idx)
    check("  ln_bwd dx (frozen)", dx2[idx] if gather else dx2, xr.grad)


def main():
    ln(1000, 1024)
    ln(77, 128)
    ln(513, 768, gather=True)
    ln(300, 512)
    ln(64, 384)
    # colsum
    T, N = 5000, 3072
    dy = torch.randn(T, N, device=dev).to(BF)
    db = torch.zeros(N, device=dev)
    L.colsum(dy, db, T=T, N=N, ld=N)
    check("colsum", db, dy.float().sum(0), tol=1e-3, atol=0.05)
    dy2 = torch.randn(33, 72, device=dev).to(BF)
    db2 = torch.zeros(72, device=dev)
    L.colsum(dy2, db2, T=33, N=72, ld=72)
    check("colsum small", db2, dy2.float().sum(0), tol=1e-3, atol=0.01)
    # patchify: image conv
    B, C, H, W, k = 3, 3, 224, 224, 14
    img = torch.randn(B, C, H, W, device=dev)
    K = C * k * k
    Kpad = (K + 7) // 8 * 8
    out = torch.empty(B * 256, Kpad, device=dev, dtype=BF)
    L.patchify(img, out, B=B, C=C, OH=16, OW=16, kh=k, kw=k, stride_h=k, stride_w=k, sb=C * H * W, sc=H * W, sh=W, sw=1, Kpad=Kpad)
    ref = F.unfold(img, k, stride=k).transpose(1, 2).reshape(B * 256, K)
    check("patchify image", out[:, :K], ref, tol=1e-2)
    print("   pad zero:", float(out[:, K:].abs().max()) if Kpad > K else 0.0)
    # audio: x [B, T, F] -> image [B,1,F,T]
    Tn, Fm = 512, 128
    x = torch.randn(B, Tn, Fm, device=dev)
    OH, OW = (Fm - 14) // 10 + 1, (Tn - 14) // 10 + 1
    out = torch.empty(B * OH * OW, 200, device=dev, dtype=BF)
    L.patchify(x, out, B=B, C=1, OH=OH, OW=OW, kh=14, kw=14, stride_h=10, stride_w=10, sb=Tn * Fm, sc=0, sh=1, sw=Fm, Kpad=200)
    ref = F.unfold(x.unsqueeze(1).transpose(2, 3), 14, stride=10).transpose(1, 2).reshape(B * OH * OW, 196)
    check("patchify audio", out[:, :196], ref, tol=1e-2)
    # assemble
    Bb, Lt, D = 5, 16, 128
    tok = torch.randn(Bb, Lt, D, device=dev).to(BF)
    cls = torch.randn(D, device=dev)
    pos = torch.randn(Lt + 1, D, device=dev)
    o = torch.empty(Bb, Lt + 1, D, device=dev, dtype=BF)
    L.assemble_tokens(tok, cls, pos, o, B=Bb, L=Lt, D=D, has_cls=True)
    ref = torch.cat([cls.view(1, 1, D).expand(Bb, 1, D), tok.float()], 1) + pos
    check("assemble cls", o, ref)
    o2 = torch.empty(Bb, Lt, D, device=dev, dtype=BF)
    L.assemble_tokens(tok, None, pos[:Lt].contiguous(), o2, B=Bb, L=Lt, D=D, has_cls=False)
    check("assemble nocls", o2, tok.float() + pos[:Lt])
    dx = torch.randn(Bb, Lt + 1, D, device=dev).to(BF)
    dtok = torch.zeros(Bb, Lt, D, device=dev, dtype=BF)
    dpos = torch.zeros(Lt + 1, D, device=dev)
    dcls = torch.zeros(D, device=dev)
    L.assemble_tokens_bwd(dx, dtok, dpos, dcls, B=Bb, L=Lt, D=D, has_cls=True)
    check("assemble_bwd dtok", dtok, dx[:, 1:].float())
    check("assemble_bwd dpos", dpos, dx.float().sum(0), atol=0.02)
    check("assemble_bwd dcls", dcls, dx.float()[:, 0].sum(0), atol=0.02)
    # embed
    V, ctx, Dm = 512, 16, 128
    ids = torch.randint(0, V, (4, ctx), device=dev)
    table = torch.randn(V, Dm, device=dev)
    ppos = torch.randn(ctx, Dm, device=dev)
    eo = torch.empty(4 * ctx, Dm, device=dev, dtype=BF)
    L.embed_tokens(ids, table, ppos, eo, rows=4 * ctx, ctx=ctx, D=Dm)
    check("embed", eo, (table[ids] + ppos).reshape(4 * ctx, Dm))
    dxe = torch.randn(4 * ctx, Dm, device=dev).to(BF)
    dt = torch.zeros(V, Dm, device=dev)
    dp = torch.zeros(ctx, Dm, device=dev)
    L.embed_tokens_bwd(ids, dxe, dt, dp, rows=4 * ctx, ctx=ctx, D=Dm)
    rt = torch.zeros(V, Dm, device=dev).index_add_(0, ids.reshape(-1), dxe.float())
    check("embed_bwd table", dt, rt, atol=0.02)
    check("embed_bwd pos", dp, dxe.float().reshape(4, ctx, Dm).sum(0), atol=0.02)
    # l2norm
    xx = torch.randn(37, 768, device=dev)
    yy = torch.empty_like(xx)
    inv = torch.empty(37, device=dev)
    L.l2norm_fwd(xx, yy, inv, B=37, E=768)
    xr = xx.clone().requires_grad_(True)
    yr = F.normalize(xr, dim=-1)
    check("l2norm", yy, yr.detach(), tol=1e-5, atol=1e-6)
    g = torch.randn_like(xx)
    yr.backward(g)
    dxx = torch.empty_like(xx)
    L.l2norm_bwd(g, yy, inv, dxx, B=37, E=768)
    check("l2norm_bwd", dxx, xr.grad, tol=1e-4, atol=1e-6)
    # geglu
    M, Fd = 300, 512
    h = torch.randn(M, 2 * Fd, device=dev).to(BF)
    go = torch.empty(M, Fd, device=dev, dtype=BF)
    L.geglu_fwd(h, go, M=M, F=Fd)
    hr = h.float().requires_grad_(True)
    a, gt = hr.chunk(2, -1)
    r = a * F.gelu(gt)
    check("geglu", go, r.detach())
    dout = torch.randn(M, Fd, device=dev).to(BF)
    r.backward(dout.float())
    dh = torch.empty(M, 2 * Fd, device=dev, dtype=BF)
    L.geglu_bwd(h, dout, dh, M=M, F=Fd)
    check("geglu_bwd", dh, hr.grad)
    # cast / add
    c = torch.randn(1003, device=dev)
    co = torch.empty(1003, device=dev, dtype=BF)
    L.cast_f32_bf16(c, co)
    check("cast", co, c, tol=1e-2)
    a1, a2 = torch.randn(4096, device=dev).to(BF), torch.randn(4096, device=dev).to(BF)
    ao = torch.empty_like(a1)
    L.add_bf16(a1, a2, ao)
    check("add", ao, a1.float() + a2.float())
    # adamw
    p = torch.randn(10000, device=dev)
    gr = torch.randn(10000, device=dev)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in (1, 2, 3):
        pr.grad = gr.clone()
        opt.step()
        L.adamw_step(p, gr, m, v, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.2, step=step)
    check("adamw", p, pr.detach(), tol=1e-5, atol=1e-6)
    # LN timing at ViT-L size
    T, D = 65792, 1024
    x = torch.randn(T, D, device=dev).to(BF)
    w = torch.ones(D, device=dev)
    b = torch.zeros(D, device=dev)
    y = torch.empty_like(x)
    mean = torch.empty(T, device=dev)
    rstd = torch.empty(T, device=dev)
    dxx = torch.empty_like(x)
    dw = torch.zeros(D, device=dev)
    dbb = torch.zeros(D, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn, byts in (
        ("ln_fwd", lambda: L.layernorm_fwd(x, w, b, y, mean, rstd, T=T, D=D, ldx=D, ldy=D), 2 * T * D * 2),
        ("ln_bwd", lambda: L.layernorm_bwd(y, x, w, mean, rstd, dxx, dw, dbb, T=T, D=D, lddy=D, ldx=D, lddx=D, dres=y, lddres=D), 4 * T * D * 2),
        ("colsum", lambda: L.colsum(x, dw, T=T, N=D, ld=D), T * D * 2),
    ):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"[time] {name}: {ms * 1e3:.1f} us  {byts / ms / 1e6:.0f} GB/s", flush=True)
    print("done", flush=True)


if __name__ == "__main__":
    main()
