"""Kernel parity on the B200, through the C ABI (vitlens_b200.lib -> libvitlens_b200.so), against plain fp32 torch
references of the same ops.  Tolerances: bf16 outputs 2e-2 of the tensor's max (8 mantissa bits + fp32 accumulate)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from vitlens_b200 import ops as o

    torch.manual_seed(0)
    return o


def close(got, ref, tol=2e-2, atol=1e-3):
    err = (got.float() - ref.float()).abs().max().item()
    sc = ref.float().abs().max().item()
    assert math.isfinite(err) and err <= tol * sc + atol, (err, sc)


def _gelu(x, quick):
    return x * torch.sigmoid(1.702 * x) if quick else F.gelu(x)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 384, 1024), (304, 72, 200), (257 * 8, 3072, 1024), (72, 128, 128)])
@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, True)])
def test_gemm_layouts(ops, M, N, K, a_t, b_t):
    A = torch.randn(M, K, device="cuda").to(BF)
    B = torch.randn(N, K, device="cuda").to(BF)
    out = ops.gemm(A.t().contiguous() if a_t else A, B.t().contiguous() if b_t else B, a_t=a_t, b_t=b_t, out_dtype=torch.float32)
    close(out, A.float() @ B.float().t(), tol=1e-3)


@pytest.mark.parametrize("quick", [False, True])
def test_gemm_epilogues(ops, quick):
    M, N, K = 520, 1024, 256
    A = torch.randn(M, K, device="cuda").to(BF)
    B = torch.randn(N, K, device="cuda").to(BF) * 0.1
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(BF)
    acc = A.float() @ B.float().t()
    close(ops.gemm(A, B, bias=bias), acc + bias)
    h, u = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, want_aux_out=True, act_quick=quick)
    close(u, acc + bias)
    close(h, _gelu(acc + bias, quick))
    close(ops.gemm(A, B, bias=bias, epilogue=ops.EPI_RESIDUAL, aux_in=res), acc + bias + res.float())
    xr = res.float().requires_grad_(True)
    g = torch.autograd.grad(_gelu(xr, quick).sum(), xr)[0]
    close(ops.gemm(A, B, epilogue=ops.EPI_GELU_BWD, aux_in=res, act_quick=quick), acc * g)


@pytest.mark.parametrize("M,N,K", [(34, 192, 192), (300, 576, 200), (130, 768, 3072), (500, 264, 64)])
def test_gemm_epilogues_small_m(ops, M, N, K):
    """M < 512 takes the single-CTA kernel with the TMA-staged epilogue (row and column tails clipped by the tensor maps)."""
    A = torch.randn(M, K, device="cuda").to(BF)
    B = torch.randn(N, K, device="cuda").to(BF) * 0.1
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(BF)
    acc = A.float() @ B.float().t()
    close(ops.gemm(A, B, bias=bias), acc + bias)
    h, u = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, want_aux_out=True)
    close(u, acc + bias)
    close(h, F.gelu(acc + bias))
    close(ops.gemm(A, B, bias=bias, epilogue=ops.EPI_RESIDUAL, aux_in=res), acc + bias + res.float())
    xr = res.float().requires_grad_(True)
    g = torch.autograd.grad(F.gelu(xr).sum(), xr)[0]
    close(ops.gemm(A, B, epilogue=ops.EPI_GELU_BWD, aux_in=res), acc * g)


def test_gemm_splitk_accumulate(ops):
    M, N, K = 384, 512, 8192
    A = torch.randn(K, M, device="cuda").to(BF)
    B = torch.randn(K, N, device="cuda").to(BF)
    out = ops.gemm(A, B, a_t=True, b_t=True, out_dtype=torch.float32, accumulate=True)
    close(out, A.float().t() @ B.float(), tol=1e-3)


@pytest.mark.parametrize("M,N,K", [(1024, 512, 8192), (4096, 1024, 4112), (520, 136, 1000)])
def test_gemm_wgrad_with_fused_bias_grad(ops, M, N, K):
    """dW = dy^T x with the bias gradient (column sums of dy = row sums of the A operand) produced by the same GEMM."""
    dy = torch.randn(K, M, device="cuda").to(BF)
    x = torch.randn(K, N, device="cuda").to(BF)
    assert ops.rowsum_fusable(M, N)
    dw, db = ops.gemm(dy, x, a_t=True, b_t=True, out_dtype=torch.float32, accumulate=True, want_rowsum=True)
    close(dw, dy.float().t() @ x.float(), tol=1e-3)
    close(db, dy.float().sum(0), tol=1e-3, atol=0.05)


def _ref_attn(q, k, v, causal):
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * 0.125
    if causal:
        n = q.shape[1]
        s = s + torch.full((n, n), float("-inf"), device=q.device).triu_(1)
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,H,nq,nk,causal,packed", [
    (2, 2, 257, 257, False, True), (3, 2, 77, 77, True, True), (2, 16, 256, 256, False, True), (2, 1, 256, 600, False, False),
    (2, 2, 17, 17, False, True), (4, 12, 50, 50, False, True), (2, 1, 128, 512, False, False), (1, 1, 1, 1, False, True),
    (2, 2, 513, 513, False, True), (2, 1, 229, 229, False, True), (1, 2, 130, 259, False, False), (2, 1, 300, 300, True, True),
    (1, 1, 640, 132, False, False), (2, 2, 16, 48, False, False), (2, 1, 8, 16, False, False), (1, 2, 40, 208, False, False),
    (2, 1, 80, 80, True, True),
    # the third-generation backward's special paths: tail queries / tail keys of N = 128 k + t (t <= 4) alone and together, several
    # query ranges with accumulation into dK / dV, the ViT-Lens latent lengths 129 = 128 + cls and 260 = 2 x 128 + 4
    (3, 2, 129, 129, False, True), (2, 2, 128, 129, False, False), (2, 2, 129, 128, False, False), (2, 3, 260, 260, False, True),
    (1, 2, 516, 258, False, False), (2, 1, 64, 131, False, False)])
def test_attention_fwd_bwd(ops, B, H, nq, nk, causal, packed):
    D = H * 64
    if packed:
        qkv = torch.randn(B * nq, 3 * D, device="cuda").to(BF)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
    else:
        q = torch.randn(B * nq, D, device="cuda").to(BF)
        kv = torch.randn(B * nk, 2 * D, device="cuda").to(BF)
        k, v = kv[:, :D], kv[:, D:]
        dq = torch.zeros_like(q)
        dkv = torch.zeros_like(kv)
        dk, dv = dkv[:, :D], dkv[:, D:]
    o, lse = ops.attention_fwd(q, k, v, B=B, H=H, nq=nq, nk=nk, causal=causal)
    qf = q.float().reshape(B, nq, H, 64).requires_grad_(True)
    kf = k.float().reshape(B, nk, H, 64).requires_grad_(True)
    vf = v.float().reshape(B, nk, H, 64).requires_grad_(True)
    oref, lref = _ref_attn(qf, kf, vf, causal)
    close(o.reshape(B, nq, H, 64), oref.detach(), tol=3e-2)
    close(lse, lref.detach(), tol=1e-3, atol=1e-3)
    do = torch.randn(B * nq, D, device="cuda").to(BF)
    ops.attention_bwd(q, k, v, o, do, lse, dq, dk, dv, B=B, H=H, nq=nq, nk=nk, causal=causal)
    gq, gk, gv = torch.autograd.grad(oref, (qf, kf, vf), do.float().reshape(B, nq, H, 64))
    close(dq.reshape(B, nq, H, 64), gq, tol=3e-2)
    close(dk.reshape(B, nk, H, 64), gk, tol=3e-2)
    close(dv.reshape(B, nk, H, 64), gv, tol=3e-2)


@pytest.mark.parametrize("B,H,nq,nk,qmul,packed", [
    (3, 4, 257, 257, 8.0, True),     # large score spread: the lazily moved softmax offset rescales O in TMEM
    (2, 2, 384, 700, 6.0, False),    # odd number of query tiles (one group idles in the last range), 11 key blocks, masked last block
    (40, 16, 257, 257, 1.0, True),   # 640 work items over 148 persistent CTAs: K/V ring wrap, Q stage recycling, deferred epilogues
    (2, 3, 257, 64, 1.0, False),     # a single key block per item (nothing to defer the epilogue behind)
    (1, 2, 260, 33, 2.0, False)])    # remainder 4 -> padded third tile (second range with one tile), 33 keys
def test_attention_fwd_persistent_kernel(ops, B, H, nq, nk, qmul, packed):
    """Shapes that route to attn_fwd2_kernel (non-causal, more than one 128-row query tile): O and LSE against fp32 torch."""
    D = H * 64
    if packed:
        qkv = torch.randn(B * nq, 3 * D, device="cuda").to(BF)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q = torch.randn(B * nq, D, device="cuda").to(BF)
        kv = torch.randn(B * nk, 2 * D, device="cuda").to(BF)
        k, v = kv[:, :D], kv[:, D:]
    q.mul_(qmul)
    o, lse = ops.attention_fwd(q, k, v, B=B, H=H, nq=nq, nk=nk, causal=False)
    oref, lref = _ref_attn(q.float().reshape(B, nq, H, 64), k.float().reshape(B, nk, H, 64), v.float().reshape(B, nk, H, 64), False)
    close(o.reshape(B, nq, H, 64), oref, tol=3e-2)
    close(lse, lref, tol=1e-3, atol=2e-3)


@pytest.mark.parametrize("T,D,gather", [(1000, 1024, False), (77, 128, False), (513, 768, True), (300, 512, False)])
def test_layernorm(ops, T, D, gather):
    x = (torch.randn(T, D, device="cuda") * 2 + 0.5).to(BF)
    w = torch.randn(D, device="cuda") * 0.1 + 1
    b = torch.randn(D, device="cuda") * 0.1
    idx = torch.randperm(T, device="cuda")[: T // 3].contiguous() if gather else None
    y, mean, rstd = ops.layernorm_fwd(x, w, b, row_index=idx)
    xs = x.float()[idx] if gather else x.float()
    xr, wr, br = xs.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), wr, br, 1e-5)
    close(y, yr.detach())
    dy = torch.randn_like(yr).to(BF)
    yr.backward(dy.float())
    dres = torch.randn(T, D, device="cuda").to(BF)
    dx, dw, db, rsum = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, row_index=idx, want_dres_sum=True)
    close(rsum, (dres.float()[idx] if gather else dres.float()).sum(0), tol=1e-3, atol=0.05)
    close(dx[idx] if gather else dx, xr.grad + (dres.float()[idx] if gather else dres.float()))
    close(dw, wr.grad, tol=1e-3, atol=0.05)
    close(db, br.grad, tol=1e-3, atol=0.05)
    if gather:
        mask = torch.ones(T, dtype=torch.bool, device="cuda")
        mask[idx] = False
        assert float(dx[mask].abs().max()) == 0.0


def test_attention_fwd_persistent_kernel_is_deterministic(ops):
    """attn_fwd2_kernel has no atomics: repeated launches on the same inputs must agree bit for bit.  A race in its
    mbarrier / TMEM hand-offs (S and P share TMEM columns, O tiles are staged in retired Q tiles) would show up here."""
    B, H, N = 64, 16, 257
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device="cuda").to(BF)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    o0, l0 = ops.attention_fwd(q, k, v, B=B, H=H, nq=N, nk=N)
    for _ in range(25):
        o, lse = ops.attention_fwd(q, k, v, B=B, H=H, nq=N, nk=N)
        assert torch.equal(o, o0) and torch.equal(lse, l0)


@pytest.mark.parametrize("T", [2048, 4099, 6 * 257 + 2048])
def test_layernorm_d1024_fast_paths(ops, T):
    """The D = 1024 kernels of every ViT-L LayerNorm (T >= 2048, no row gather): forward with its weight slice in registers,
    backward with eight columns per thread (with / without the residual-gradient add, with / without weight gradients)."""
    D = 1024
    x = (torch.randn(T, D, device="cuda") * 2 + 0.5).to(BF)
    w = torch.randn(D, device="cuda") * 0.1 + 1
    b = torch.randn(D, device="cuda") * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, w, b)
    xr, wr, br = x.float().clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), wr, br, 1e-5)
    close(y, yr.detach())
    close(mean, x.float().mean(1), tol=1e-4, atol=1e-4)
    dy = torch.randn_like(yr).to(BF)
    yr.backward(dy.float())
    dres = torch.randn(T, D, device="cuda").to(BF)
    dx, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres)
    close(dx, xr.grad + dres.float())
    close(dw, wr.grad, tol=1e-3, atol=0.05)
    close(db, br.grad, tol=1e-3, atol=0.05)
    dx2, dw2, db2 = ops.layernorm_bwd(dy, x, w, mean, rstd, want_wgrad=False)
    assert dw2 is None and db2 is None
    close(dx2, xr.grad)


def test_row_kernels(ops):
    dy = torch.randn(5000, 3072, device="cuda").to(BF)
    close(ops.colsum(dy), dy.float().sum(0), tol=1e-3, atol=0.05)
    img = torch.randn(3, 3, 224, 224, device="cuda")
    cols = ops.patchify(img, B=3, C=3, OH=16, OW=16, kh=14, kw=14, stride_h=14, stride_w=14, sb=3 * 224 * 224, sc=224 * 224, sh=224, sw=1, Kpad=592)
    close(cols[:, :588], F.unfold(img, 14, stride=14).transpose(1, 2).reshape(-1, 588), tol=1e-2)
    assert float(cols[:, 588:].abs().max()) == 0.0
    x = torch.randn(3, 512, 128, device="cuda")
    cols = ops.patchify(x, B=3, C=1, OH=12, OW=50, kh=14, kw=14, stride_h=10, stride_w=10, sb=512 * 128, sc=0, sh=1, sw=128, Kpad=200)
    close(cols[:, :196], F.unfold(x.unsqueeze(1).transpose(2, 3), 14, stride=10).transpose(1, 2).reshape(-1, 196), tol=1e-2)
    tok = torch.randn(5 * 16, 128, device="cuda").to(BF)
    cls, pos = torch.randn(128, device="cuda"), torch.randn(17, 128, device="cuda")
    out = ops.assemble_tokens(tok, cls, pos, B=5, L_=16, D=128)
    ref = torch.cat([cls.view(1, 1, 128).expand(5, 1, 128), tok.float().reshape(5, 16, 128)], 1) + pos
    close(out.reshape(5, 17, 128), ref)
    dx = torch.randn(5 * 17, 128, device="cuda").to(BF)
    dtok, dpos, dcls = ops.assemble_tokens_bwd(dx, B=5, L_=16, D=128, has_cls=True)
    d3 = dx.float().reshape(5, 17, 128)
    close(dtok.reshape(5, 16, 128), d3[:, 1:])
    close(dpos, d3.sum(0), atol=0.02)
    close(dcls, d3[:, 0].sum(0), atol=0.02)
    xx = torch.randn(37, 768, device="cuda")
    yy, inv = ops.l2norm_fwd(xx)
    xr = xx.clone().requires_grad_(True)
    yr = F.normalize(xr, dim=-1)
    close(yy, yr.detach(), tol=1e-5, atol=1e-6)
    g = torch.randn_like(xx)
    yr.backward(g)
    close(ops.l2norm_bwd(g, yy, inv), xr.grad, tol=1e-4, atol=1e-6)
    h = torch.randn(300, 1024, device="cuda").to(BF)
    hr = h.float().requires_grad_(True)
    a, gt = hr.chunk(2, -1)
    r = a * F.gelu(gt)
    close(ops.geglu_fwd(h), r.detach())
    dout = torch.randn(300, 512, device="cuda").to(BF)
    r.backward(dout.float())
    close(ops.geglu_bwd(h, dout), hr.grad)


@pytest.mark.parametrize("Bl,Ball,off", [(8, 8, 0), (4, 4, 0), (512, 512, 0), (64, 256, 128), (100, 300, 100)])
def test_contrastive_epilogues(ops, Bl, Ball, off):
    E = 768
    x = F.normalize(torch.randn(Bl, E, device="cuda"), dim=-1).to(BF)
    y = F.normalize(torch.randn(Ball, E, device="cuda"), dim=-1).to(BF)
    s = 14.3
    s_dev = torch.tensor([s], device="cuda")
    lse, tot = ops.rowlse(x, y, alpha=s_dev, label_off=off)
    z = s * (x.float() @ y.float().t())
    ref = torch.logsumexp(z, -1)
    close(lse, ref, tol=1e-4, atol=1e-3)
    diag = z[torch.arange(Bl), torch.arange(Bl) + off]
    assert abs(float(tot) - float((ref - diag).sum())) < 1e-3 * Bl
    col = torch.randn(Ball, device="cuda") + 3
    for cl in (None, col):
        g, ds = ops.clipgrad(x, y, alpha=s_dev, row_lse=lse, col_lse=cl, label_off=off, gscale=0.02, gscale_dev=torch.tensor([0.5], device="cuda"))
        gr = torch.exp(z - lse[:, None])
        k = 1.0
        if cl is not None:
            gr = gr + torch.exp(z - cl[None, :])
            k = 2.0
        oh = torch.zeros_like(z)
        oh[torch.arange(Bl), torch.arange(Bl) + off] = 1
        gr = 0.01 * (gr - k * oh)
        close(g, gr, tol=1e-2, atol=1e-5)
        assert abs(float(ds) - float((gr * (z / s)).sum())) < 2e-2 * float((gr * (z / s)).abs().sum()) + 1e-4


@pytest.mark.parametrize("Bl,Ball,off,E,masked", [(8, 8, 0, 768, False), (512, 512, 0, 768, False), (64, 256, 128, 768, True), (100, 300, 100, 512, True),
                                                  (256, 2048, 512, 1024, False), (130, 130, 0, 64, False), (256, 1024, 256, 640, False)])
def test_fused_contrastive_backward(ops, Bl, Ball, off, E, masked):
    """vl_clip_backward (one launch per direction: logits and their gradient stay in TMEM) against the two-GEMM path (VL_EPI_CLIPGRAD
    writes g to HBM, a plain GEMM reads it back) and against fp32 torch: dX = s * g @ Y, d(loss)/d(scale) sum, with and without the
    column term, the row-only d(scale) of the local-loss variants, and the byte mask of the mask losses."""
    torch.manual_seed(3)
    x = F.normalize(torch.randn(Bl, E, device="cuda"), dim=-1).to(BF)
    y = F.normalize(torch.randn(Ball, E, device="cuda"), dim=-1).to(BF)
    s = 14.3
    s_dev = torch.tensor([s], device="cuda")
    mask = (torch.rand(Bl, Ball, device="cuda") > 0.2).to(torch.uint8) if masked else None
    if masked:
        mask[torch.arange(Bl), torch.arange(Bl) + off] = 1
    lse, _ = ops.rowlse(x, y, alpha=s_dev, label_off=off, **({} if mask is None else dict(mask=mask)))
    col = torch.randn(Ball, device="cuda") + 3
    assert ops.clip_backward_fusable(E, y)
    for cl in (None, col):
        for row_only in (False, True):
            kw = dict(alpha=s_dev, row_lse=lse, col_lse=cl, label_off=off, gscale=0.02, gscale_dev=torch.tensor([0.5], device="cuda"),
                      ds_row_only=row_only, **({} if mask is None else dict(mask=mask)))
            dx, ds = ops.clip_backward(x, y, **kw)
            g, ds_ref = ops.clipgrad(x, y, **kw)
            dx_ref = ops.gemm(g.contiguous() if g.stride(0) % 8 else g, y, b_t=True, out_dtype=torch.float32, alpha_dev=s_dev)
            close(dx, dx_ref, tol=2e-3, atol=1e-6)   # same bf16-rounded g, fp32 accumulation in a different grouping
            assert abs(float(ds) - float(ds_ref)) <= 1e-4 * abs(float(ds_ref)) + 1e-6, (float(ds), float(ds_ref))
            # fp32 torch
            acc = x.float() @ y.float().t()
            z = s * acc
            keep = torch.ones_like(z) if mask is None else (mask != 0).float()
            z = z * keep
            gr = torch.exp(z - lse[:, None])
            k = 1.0
            if cl is not None:
                gr = gr + torch.exp(z - cl[None, :])
                k = 2.0
            oh = torch.zeros_like(z)
            oh[torch.arange(Bl), torch.arange(Bl) + off] = 1
            gr = 0.01 * (gr - k * oh) * keep
            close(dx, s * (gr @ y.float()), tol=1e-2, atol=1e-6)
    torch.cuda.synchronize()
    d1, _ = ops.clip_backward(x, y, **kw)
    d2, _ = ops.clip_backward(x, y, **kw)
    assert torch.equal(d1, d2)  # deterministic


def test_adamw(ops):
    from vitlens_b200 import lib as L

    p = torch.randn(10000, device="cuda")
    gr = torch.randn(10000, device="cuda")
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2, 3):
        pr.grad = gr.clone()
        opt.step()
        L.adamw_step(p, gr, m, v, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.2, step=step)
    close(p, pr.detach(), tol=1e-5, atol=1e-6)


def test_adamw_multi_tensor(ops):
    """The one-launch optimizer equals torch.optim.AdamW over decay / no-decay groups (open_clip main.py:328-345
    grouping) and refreshes the cached bf16 operand copies in the same pass."""
    from vitlens_b200 import engine, optim

    torch.manual_seed(1)
    shapes = {"a.weight": (300, 130), "a.bias": (300,), "ln.weight": (130,), "big.weight": (40000, 3), "logit_scale": ()}
    ours = {k: torch.nn.Parameter(torch.randn(s, device="cuda")) for k, s in shapes.items()}
    ref = {k: torch.nn.Parameter(v.detach().clone()) for k, v in ours.items()}
    no_decay = [k for k, v in ref.items() if v.ndim < 2 or "ln" in k or "bias" in k or "logit_scale" in k]
    ropt = torch.optim.AdamW([dict(params=[ref[k] for k in no_decay], weight_decay=0.0),
                              dict(params=[ref[k] for k in ref if k not in no_decay], weight_decay=0.2)], lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    opt = optim.AdamW(list(ours.items()), lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    cached = engine.w16(ours["a.weight"])
    for step in range(3):
        for k in ours:
            g = torch.randn(shapes[k], device="cuda")
            ours[k].grad = g.clone()
            ref[k].grad = g.clone()
        opt.step()
        ropt.step()
        assert engine.w16(ours["a.weight"]) is cached  # refreshed in place, not re-cast
    for k in ours:
        close(ours[k].detach(), ref[k].detach(), tol=1e-5, atol=1e-6)
    close(cached.float(), ours["a.weight"].detach().bfloat16().float(), tol=0, atol=0)


@pytest.mark.parametrize("max_norm,gscale", [(0.5, 1.0), (1e4, 1.0), (2.0, 0.25)])
def test_adamw_gradient_norm_clipping(ops, max_norm, gscale):
    """--grad-clip-norm (train.py:212-240): clip_grad_norm_ over ALL parameters + AdamW == vl_multi_sqnorm + vl_adamw_multi_clip
    (the clip coefficient is applied inside the fused update; the gradients themselves are left alone)."""
    from vitlens_b200 import optim

    torch.manual_seed(2)
    shapes = {"a.weight": (257, 130), "a.bias": (257,), "big.weight": (50000, 3), "logit_scale": ()}
    ours = {k: torch.nn.Parameter(torch.randn(s, device="cuda")) for k, s in shapes.items()}
    ref = {k: torch.nn.Parameter(v.detach().clone()) for k, v in ours.items()}
    no_decay = [k for k, v in ref.items() if v.ndim < 2 or "bias" in k or "logit_scale" in k]
    ropt = torch.optim.AdamW([dict(params=[ref[k] for k in no_decay], weight_decay=0.0),
                              dict(params=[ref[k] for k in ref if k not in no_decay], weight_decay=0.2)], lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    opt = optim.AdamW(list(ours.items()), lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    for step in range(3):
        for k in ours:
            g = torch.randn(shapes[k], device="cuda") * (0.01 if step == 1 else 1.0)
            ours[k].grad = g.clone()
            ref[k].grad = g.clone() * gscale
        want_norm = torch.nn.utils.clip_grad_norm_(list(ref.values()), max_norm, norm_type=2.0)
        ropt.step()
        opt.step(grad_scale=gscale, clip_norm=max_norm)
        assert abs(float(opt.grad_norm) - float(want_norm)) < 1e-4 * float(want_norm)
    for k in ours:
        close(ours[k].detach(), ref[k].detach(), tol=1e-5, atol=2e-6)


@pytest.mark.parametrize("B,N,G,k", [(3, 256, 16, 8), (2, 8192, 512, 32), (2, 1000, 64, 32)])
def test_point_cloud_sampling_and_grouping(ops, B, N, G, k):
    """FPS indices are bit-exact against the oracle's restatement of misc.fps; kNN neighbourhoods equal the exact
    k-nearest sets (compared as sorted distance lists: ties / fp re-association may swap equidistant points)."""
    from oracle import vitlens_oracle as O
    from vitlens_b200 import synth

    pts, start = synth.synth_points(B, N, seed=3)
    idx, centers = ops.fps(pts.cuda(), start.cuda(), G)
    ref_idx = O.fps_indices(pts, G, start)
    assert torch.equal(idx.cpu(), ref_idx)
    ref_c = torch.gather(pts, 1, ref_idx.unsqueeze(-1).expand(-1, -1, 3)).reshape(B * G, 3)
    assert torch.equal(centers.cpu(), ref_c)
    nb, nidx = ops.knn_group(pts.cuda(), centers, G, k, want_idx=True)
    d = ((pts.unsqueeze(1) - ref_c.reshape(B, G, 1, 3)) ** 2).sum(-1)           # [B,G,N] exact form
    ref_sorted = torch.topk(d, k, dim=-1, largest=False, sorted=True)[0]
    got_d = torch.gather(d, 2, nidx.cpu().reshape(B, G, k)).sort(dim=-1)[0]
    assert float((got_d - ref_sorted).abs().max()) < 1e-5
    sel = torch.gather(pts.unsqueeze(1).expand(-1, G, -1, -1), 2, nidx.cpu().reshape(B, G, k, 1).expand(-1, -1, -1, 3))
    assert torch.equal(nb.cpu().reshape(B, G, k, 3), sel - ref_c.reshape(B, G, 1, 3))


def test_point_cloud_row_kernels(ops):
    x = torch.randn(1000, 3, device="cuda")
    w = torch.randn(128, 3, device="cuda")
    sc, sh = torch.rand(128, device="cuda") + 0.5, torch.randn(128, device="cuda")
    close(ops.linear3(x, w, sc, sh, 1), torch.relu((x @ w.t()) * sc + sh))
    close(ops.linear3(x, w, sc, sh, 2), F.gelu((x @ w.t()) * sc + sh))
    f = torch.randn(64 * 32, 256, device="cuda").to(BF)
    mx, arg = ops.group_max(f, 32, want_arg=True)
    rv, ra = f.float().reshape(64, 32, 256).max(dim=1)
    assert torch.equal(mx.float(), rv)
    assert torch.equal(torch.gather(f.float().reshape(64, 32, 256), 1, arg.long().unsqueeze(1)).squeeze(1), rv)
    a = torch.randn(64 * 32, 256, device="cuda").to(BF)
    b = (torch.randn(512, 256, device="cuda") * 0.1).to(BF)
    gp = torch.randn(64, 512, device="cuda").to(BF)
    out = ops.gemm_grouped_residual_relu(a, b, gp, 32)
    close(out, torch.relu(a.float() @ b.float().t() + gp.float().repeat_interleave(32, 0)))


def test_point_cloud_backward_kernels(ops):
    d = torch.randn(64, 256, device="cuda").to(BF)
    arg = torch.randint(0, 32, (64, 256), device="cuda", dtype=torch.int32)
    dx = ops.group_max_bwd(d, arg, 32)
    ref = torch.zeros(64, 32, 256, device="cuda").scatter_(1, arg.long().unsqueeze(1), d.float().unsqueeze(1))
    assert torch.equal(dx.float().reshape(64, 32, 256), ref)
    x = torch.randn(64 * 32, 512, device="cuda").to(BF)
    close(ops.group_sum(x, 32), x.float().reshape(64, 32, 512).sum(1))
    a, b = torch.randn(3000, 512, device="cuda").to(BF), torch.randn(3000, 512, device="cuda").to(BF)
    s1, s2 = ops.colsum2(a, b)
    close(s1, a.float().sum(0), tol=1e-3, atol=0.05)
    close(s2, (a.float() * b.float()).sum(0), tol=1e-3, atol=0.05)
    dy, xx = torch.randn(5000, 128, device="cuda").to(BF), torch.randn(5000, 3, device="cuda")
    close(ops.wgrad3(dy, xx), dy.float().t() @ xx, tol=1e-3, atol=0.05)
    w = torch.randn(128, 3, device="cuda")
    out, pre = ops.linear3(xx, w, torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"), 2, want_pre=True)
    close(pre, xx @ w.t())
    close(out, F.gelu(xx @ w.t()))
    A = torch.randn(520, 256, device="cuda").to(BF)
    Bm = (torch.randn(1024, 256, device="cuda") * 0.1).to(BF)
    res = torch.randn(520, 1024, device="cuda").to(BF)
    acc = A.float() @ Bm.float().t()
    close(ops.gemm(A, Bm, epilogue=ops.EPI_GELU_BWD, aux_in=res, act_quick=2), acc * (res.float() > 0))


def test_point_cloud_batchnorm_kernels(ops):
    """vl_moments3 / vl_col_affine_bf16 (training-mode BatchNorm of the point tokenizer) and the ReLU-free grouped GEMM."""
    x = torch.randn(70001, 3, device="cuda") * torch.tensor([1.0, 0.3, 2.0], device="cuda") + torch.tensor([0.1, -0.2, 0.05], device="cuda")
    m = ops.moments3(x)
    xd = x.double()
    ref = torch.cat([xd.sum(0), (xd.t() @ xd).reshape(9)]).float()
    close(m, ref, tol=1e-4, atol=1e-2)
    R, C = 4099, 512
    a, b = torch.randn(R, C, device="cuda").to(BF), torch.randn(R, C, device="cuda").to(BF)
    p0, p1, p2 = (torch.randn(C, device="cuda") for _ in range(3))
    close(ops.col_affine(a, p0, p2, relu=True), torch.relu(a.float() * p0 + p2))
    close(ops.col_affine(a, p0, p2, b=b, p1=p1), a.float() * p0 + b.float() * p1 + p2)
    k = 8
    A = torch.randn(64 * k, 256, device="cuda").to(BF)
    Bm = (torch.randn(512, 256, device="cuda") * 0.1).to(BF)
    gp = torch.randn(64, 512, device="cuda").to(BF)
    want = A.float() @ Bm.float().t() + gp.float().repeat_interleave(k, dim=0)
    close(ops.gemm_grouped_residual_relu(A, Bm, gp, k, relu=False), want)
    close(ops.gemm_grouped_residual_relu(A, Bm, gp, k), torch.relu(want))


def test_row_reductions_are_bit_deterministic(ops):
    """Every cross-CTA reduction on the path (LayerNorm dgamma / dbeta, bias-gradient column sums incl. the ones riding on a
    weight-gradient GEMM, assemble's dpos / dcls, the loss sum and d(logit_scale), BatchNorm sums and moments) is a two-stage
    fixed-order sum: repeated launches on the same inputs must agree bit for bit, and outputs start from uninitialised memory."""
    torch.manual_seed(3)
    T, D = 6 * 257 + 2048, 1024
    x = torch.randn(T, D, device="cuda").to(BF)
    dy = torch.randn(T, D, device="cuda").to(BF)
    w = torch.randn(D, device="cuda")
    y, mean, rstd = ops.layernorm_fwd(x, w, torch.zeros_like(w))
    pts = torch.randn(5000, 3, device="cuda")
    runs = []
    for _ in range(3):
        torch.empty(1 << 22, device="cuda").fill_(float("nan"))  # poison the allocator's free blocks
        r = {}
        _, r["ln_dw"], r["ln_db"] = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dy)
        _, r["lng_dw"], r["lng_db"], r["lng_rs"] = ops.layernorm_bwd(dy[:300, :512].contiguous(), x[:300, :512].contiguous(), w[:512],
                                                                    mean[:300], rstd[:300], dres=dy[:300, :512].contiguous(), want_dres_sum=True)
        r["colsum"] = ops.colsum(dy)
        _, r["rowsum"] = ops.gemm(dy, x, a_t=True, b_t=True, out_dtype=torch.float32, accumulate=True, want_rowsum=True)
        _, r["dpos"], r["dcls"] = ops.assemble_tokens_bwd(dy[: 6 * 257].contiguous(), B=6, L_=256, D=D, has_cls=True)
        p16 = torch.nn.functional.normalize(x[:512, :768].float(), dim=-1).to(BF).contiguous()
        q16 = torch.nn.functional.normalize(dy[:512, :768].float(), dim=-1).to(BF).contiguous()
        alpha = torch.tensor([14.3], device="cuda")
        lse, r["loss_sum"] = ops.rowlse(p16, q16, alpha=alpha)
        _, r["dscale"] = ops.clipgrad(p16, q16, alpha=alpha, row_lse=lse, col_lse=lse, label_off=0, gscale=1.0 / 1024)
        r["cs2a"], r["cs2b"] = ops.colsum2(dy, x)
        r["mom"] = ops.moments3(pts)
        r["wg3"] = ops.wgrad3(dy[:5000, :128].contiguous(), pts)
        torch.cuda.synchronize()
        assert all(torch.isfinite(v).all() for v in r.values()), [k for k, v in r.items() if not torch.isfinite(v).all()]
        runs.append(r)
    for k in runs[0]:
        for other in runs[1:]:
            assert torch.equal(runs[0][k], other[k]), k
    close(runs[0]["colsum"], dy.float().sum(0), tol=1e-3, atol=0.05)
    close(runs[0]["ln_db"], dy.float().sum(0), tol=1e-3, atol=0.05)
    close(runs[0]["dcls"], dy[: 6 * 257].float().reshape(6, 257, D)[:, 0].sum(0), tol=1e-3, atol=0.02)
    close(runs[0]["mom"][:3], pts.sum(0), tol=1e-3, atol=0.05)


def test_adamw_param_groups_state_dict_and_missing_grads(ops):
    """torch.optim.AdamW drop-in behaviour the reference's loop relies on: per-group learning rates rewritten by the scheduler
    (assign_learning_rate writes param_group["lr"]), parameters whose grad is None are skipped, state_dict() / load_state_dict()
    round-trip in torch's layout (a torch.optim.AdamW state dict loads into this optimizer and vice versa), and the plan follows
    a parameter whose storage moved."""
    from vitlens_b200 import optim

    torch.manual_seed(4)
    shapes = {"w": (130, 72), "b": (130,), "unused": (17, 8), "big": (40000, 3)}
    ours = {k: torch.nn.Parameter(torch.randn(s, device="cuda")) for k, s in shapes.items()}
    ref = {k: torch.nn.Parameter(v.detach().clone()) for k, v in ours.items()}

    def groups(d):
        return [dict(params=[d["b"]], weight_decay=0.0, lr=3e-3), dict(params=[d["w"], d["unused"], d["big"]], weight_decay=0.2)]

    ropt = torch.optim.AdamW(groups(ref), lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    opt = optim.AdamW(groups(ours), lr=1e-3, betas=(0.9, 0.98), eps=1e-6)

    def run(o_ours, o_ref, steps, lr_mul=1.0):
        for step in range(steps):
            for k in ours:
                if k == "unused":
                    continue  # never receives a gradient
                g = torch.randn(shapes[k], device="cuda")
                ours[k].grad, ref[k].grad = g.clone(), g.clone()
            for o in (o_ours, o_ref):  # the scheduler's assign_learning_rate
                for grp in o.param_groups:
                    grp["lr"] = grp["lr"] * lr_mul
            o_ours.step()
            o_ref.step()

    run(opt, ropt, 3, lr_mul=0.9)
    for k in ours:
        close(ours[k].detach(), ref[k].detach(), tol=1e-5, atol=1e-6)
    assert torch.equal(ours["unused"].detach(), ref["unused"].detach())
    # checkpoint / resume through each other's state dicts
    sd_ours, sd_ref = opt.state_dict(), ropt.state_dict()
    assert sd_ours["param_groups"][0]["params"] == sd_ref["param_groups"][0]["params"]
    opt2 = optim.AdamW(groups(ours), lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    opt2.load_state_dict(sd_ref)  # torch -> ours
    ropt2 = torch.optim.AdamW(groups(ref), lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    ropt2.load_state_dict(sd_ours)  # ours -> torch
    with torch.no_grad():  # move one parameter's storage: the fused plan must notice
        ours["w"].data = ours["w"].data.clone()
    run(opt2, ropt2, 2)
    for k in ours:
        close(ours[k].detach(), ref[k].detach(), tol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,H,n", [(3, 16, 257), (2, 4, 600), (2, 2, 77)])
def test_attention_backward_is_bit_deterministic(ops, B, H, n):
    """The attention backward (incl. the CUDA-core tail rows of N = 128k + t, whose mat-vec partial sums meet in a fixed warp
    order) returns bit-identical dQ / dK / dV on repeated launches."""
    torch.manual_seed(5)
    D = H * 64
    qkv = torch.randn(B * n, 3 * D, device="cuda").to(BF)
    do = torch.randn(B * n, D, device="cuda").to(BF)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    o, lse = ops.attention_fwd(q, k, v, B=B, H=H, nq=n, nk=n)
    outs = []
    for _ in range(4):
        d = torch.empty_like(qkv)
        ops.attention_bwd(q, k, v, o, do, lse, d[:, :D], d[:, D:2 * D], d[:, 2 * D:], B=B, H=H, nq=n, nk=n)
        torch.cuda.synchronize()
        outs.append(d)
    for other in outs[1:]:
        assert torch.equal(outs[0], other)


@pytest.mark.parametrize("B,H,nq,nk", [(2, 4, 257, 257), (2, 2, 256, 600), (2, 2, 129, 129)])
def test_attention_backward_generations_agree(ops, B, H, nq, nk):
    """attn_bwd3 (keys on the TMEM lanes; default for non-causal attention) against attn_bwd2 (queries on the lanes; still the causal
    text-tower kernel, `vl_debug_set(12, 2)` forces it): same math, different summation orders -> agreement to bf16 rounding."""
    from vitlens_b200 import lib as L

    torch.manual_seed(11)
    D = H * 64
    q = torch.randn(B * nq, D, device="cuda").to(BF)
    kv = torch.randn(B * nk, 2 * D, device="cuda").to(BF)
    k, v = kv[:, :D], kv[:, D:]
    do = torch.randn(B * nq, D, device="cuda").to(BF)
    o, lse = ops.attention_fwd(q, k, v, B=B, H=H, nq=nq, nk=nk)
    res = []
    for knob in (0, 2):
        L.debug_set(12, knob)
        try:
            dq, dkv = torch.zeros_like(q), torch.zeros_like(kv)
            ops.attention_bwd(q, k, v, o, do, lse, dq, dkv[:, :D], dkv[:, D:], B=B, H=H, nq=nq, nk=nk)
            torch.cuda.synchronize()
        finally:
            L.debug_set(12, 0)
        res.append((dq, dkv))
    close(res[0][0], res[1][0].float(), tol=1e-2)
    close(res[0][1], res[1][1].float(), tol=1e-2)


@pytest.mark.parametrize("W,rows", [(1, 300), (2, 256), (4, 512)])
def test_peer_sharded_gemm_matches_gathered_operand(ops, W, rows):
    """vl_gemm_bf16 with b_peers (the feature all-gather fused into the loss GEMMs): B's row blocks live in separate buffers --
    here W local buffers standing in for W ranks' arenas; over NVLink only the addresses differ -- with a ticket flag per block.
    Row-LSE, gradient epilogue and the dX GEMM must equal the same calls on the concatenated matrix bit for bit."""
    torch.manual_seed(6)
    E, Bl = 768, rows
    blocks = [torch.nn.functional.normalize(torch.randn(rows, E, device="cuda"), dim=-1).to(BF).contiguous() for _ in range(W)]
    cat = torch.cat(blocks).contiguous()
    x = torch.nn.functional.normalize(torch.randn(Bl, E, device="cuda"), dim=-1).to(BF).contiguous()
    flags = torch.full((8,), 7, device="cuda", dtype=torch.int32)
    peer = ops.PeerRows(addrs=[b.data_ptr() for b in blocks], rows=rows, E=E, flags=flags.data_ptr(), ticket=7, local=blocks[0])
    alpha = torch.tensor([14.3], device="cuda")
    off = rows * (W - 1)
    lse_a, sum_a = ops.rowlse(x, cat, alpha=alpha, label_off=off)
    lse_b, sum_b = ops.rowlse(x, peer, alpha=alpha, label_off=off)
    assert torch.equal(lse_a, lse_b) and torch.equal(sum_a, sum_b)
    col = torch.randn(W * rows, device="cuda") * 0.1 + float(lse_a.mean())
    ga, dsa = ops.clipgrad(x, cat, alpha=alpha, row_lse=lse_a, col_lse=col, label_off=off, gscale=0.5 / Bl)
    gb, dsb = ops.clipgrad(x, peer, alpha=alpha, row_lse=lse_a, col_lse=col, label_off=off, gscale=0.5 / Bl)
    assert torch.equal(ga, gb) and torch.equal(dsa, dsb)
    if rows % 64 == 0:
        dxa = ops.gemm(ga, cat, b_t=True, out_dtype=torch.float32, alpha_dev=alpha)
        dxb = ops.gemm(ga, peer, b_t=True, out_dtype=torch.float32, alpha_dev=alpha)
        close(dxb, dxa, tol=1e-5, atol=1e-7)  # (pair kernel vs single-CTA kernel: same products, different accumulation grouping)
        close(dxb, alpha * (ga.float() @ cat.float()), tol=2e-3)


@pytest.mark.parametrize("M,K,Fd", [(1024, 1024, 4096), (640, 256, 512), (65 * 128 + 40, 1024, 4096)])
def test_fused_geglu_epilogue_equals_unfused(ops, M, K, Fd):
    """VL_EPI_GEGLU (Lens FeedForward, perceiver.py:85-102, one launch) against Linear + the stand-alone GEGLU kernel and against
    fp32 torch: the saved pre-activations h come back in the original (value | gate) column order."""
    torch.manual_seed(7)
    a = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn(2 * Fd, K, device="cuda") * K ** -0.5)
    b = torch.randn(2 * Fd, device="cuda") * 0.1
    w16 = w.to(BF)
    h_ref = ops.gemm(a, w16, bias=b)
    gg_ref = ops.geglu_fwd(h_ref)
    gg, h = ops.gemm_geglu(a, ops.cast_bf16(ops.geglu_permute_rows(w)), ops.geglu_permute_rows(b))
    assert torch.equal(h, h_ref)                       # same accumulators, same bias, same rounding
    close(gg, gg_ref, tol=1e-2)                        # the fused product is taken from the fp32 pre-activations, the unfused from bf16
    full = a.float() @ w16.float().t() + b
    close(gg, full[:, :Fd] * F.gelu(full[:, Fd:]), tol=2e-2)
