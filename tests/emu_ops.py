"""Torch-CPU emulation of vitlens_b200.ops (same function signatures)  --  TEST INFRASTRUCTURE ONLY.

Lets the CPU-only test tier exercise the host logic of vitlens_b200.engine / open_clip.* (which
kernel is called with which operands, what is saved for backward, how gradients are assembled)
against the oracle.  It mimics the kernels' numerics contract: bf16 storage of activations, fp32
accumulation / statistics.  It is never imported by the product package.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BF16, F32 = torch.bfloat16, torch.float32
EPI_LINEAR, EPI_GELU, EPI_RESIDUAL, EPI_GELU_BWD = 0, 1, 2, 3
launches = 0


def _n():
    global launches
    launches += 1


def _act(x, quick):
    if quick == 2:
        return torch.relu(x)
    return x * torch.sigmoid(1.702 * x) if quick else F.gelu(x)


def _act_grad(x, quick):
    if quick == 2:
        return (x > 0).float()
    if quick:
        s = torch.sigmoid(1.702 * x)
        return s * (1 + 1.702 * x * (1 - s))
    return 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


def rowsum_fusable(M, N):
    return M >= 512 and N > 128


def gemm(a, b, *, a_t=False, b_t=False, bias=None, epilogue=EPI_LINEAR, aux_in=None, want_aux_out=False, out=None,
         out_dtype=BF16, alpha=1.0, accumulate=False, split_k=None, act_quick=False, alpha_dev=None, want_rowsum=False):
    _n()
    if alpha_dev is not None:
        alpha = alpha * float(alpha_dev)
    assert a.dtype == BF16 and b.dtype == BF16
    A = a.float().t() if a_t else a.float()
    B = b.float() if b_t else b.float().t()
    acc = (A @ B) * alpha
    aux_out = None
    if epilogue == EPI_GELU_BWD:
        res = acc * _act_grad(aux_in.float(), act_quick)
    else:
        if bias is not None:
            acc = acc + bias
        if epilogue == EPI_GELU:
            if want_aux_out:
                aux_out = acc.to(BF16)
            res = _act(acc, act_quick)
        elif epilogue == EPI_RESIDUAL:
            res = acc + aux_in.float()
        else:
            res = acc
    if out is None:
        out = torch.zeros(res.shape, dtype=out_dtype)
        accumulate = False
    if accumulate:
        out += res.to(out.dtype)
    else:
        out.copy_(res.to(out.dtype))
    if want_rowsum:
        return out, A.sum(1)
    return (out, aux_out) if want_aux_out else out


def _split(t, B, n, H):
    return t.float().reshape(B, n, H, 64).permute(0, 2, 1, 3)


def attention_fwd(q, k, v, *, B, H, nq, nk, causal=False, scale=None):
    _n()
    scale = 64 ** -0.5 if scale is None else scale
    s = (_split(q, B, nq, H) @ _split(k, B, nk, H).transpose(-1, -2)) * scale
    if causal:
        s = s + torch.full((nq, nk), float("-inf")).triu_(1)
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - lse.unsqueeze(-1)).to(BF16).float()  # P is rounded to bf16 before P@V on the tensor cores
    o = (p @ _split(v, B, nk, H)).permute(0, 2, 1, 3).reshape(B * nq, H * 64).to(BF16)
    return o, lse


def attention_bwd(q, k, v, o, dout, lse, dq, dk, dv, *, B, H, nq, nk, causal=False, scale=None):
    _n()
    scale = 64 ** -0.5 if scale is None else scale
    Q, K, V = _split(q, B, nq, H), _split(k, B, nk, H), _split(v, B, nk, H)
    dO, O = _split(dout, B, nq, H), _split(o, B, nq, H)
    s = (Q @ K.transpose(-1, -2)) * scale
    if causal:
        s = s + torch.full((nq, nk), float("-inf")).triu_(1)
    p = torch.exp(s - lse.unsqueeze(-1))
    D = (dO * O).sum(-1, keepdim=True)
    dP = dO @ V.transpose(-1, -2)
    dS = (p * (dP - D) * scale).to(BF16).float()
    pb = p.to(BF16).float()

    def merge(t, n):
        return t.permute(0, 2, 1, 3).reshape(B * n, H * 64).to(BF16)

    dv.copy_(merge(pb.transpose(-1, -2) @ dO, nk))
    dk.copy_(merge(dS.transpose(-1, -2) @ Q, nk))
    dq.copy_(merge(dS @ K, nq))


def layernorm_fwd(x, w, b, *, row_index=None, eps=1e-5, want_stats=True):
    _n()
    xs = x.float() if row_index is None else x.float()[row_index]
    mean = xs.mean(-1)
    var = ((xs - mean[:, None]) ** 2).mean(-1)
    rstd = torch.rsqrt(var + eps)
    y = ((xs - mean[:, None]) * rstd[:, None] * w + b).to(BF16)
    return y, (mean if want_stats else None), (rstd if want_stats else None)


def layernorm_bwd(dy, x, w, mean, rstd, *, dres=None, row_index=None, want_wgrad=True, want_dres_sum=False):
    _n()
    xs = x.float() if row_index is None else x.float()[row_index]
    g = dy.float()
    xh = (xs - mean[:, None]) * rstd[:, None]
    wg = g * w
    dxs = rstd[:, None] * (wg - wg.mean(-1, keepdim=True) - xh * (wg * xh).mean(-1, keepdim=True))
    if row_index is None:
        if dres is not None:
            dxs = dxs + dres.float()
        dx = dxs.to(BF16)
    else:
        dx = torch.zeros(x.shape, dtype=BF16)
        if dres is not None:
            dxs = dxs + dres.float()[row_index]
        dx[row_index] = dxs.to(BF16)
    out = (dx, (g * xh).sum(0), g.sum(0)) if want_wgrad else (dx, None, None)
    if want_dres_sum:
        rs = dres.float() if row_index is None else dres.float()[row_index]
        out = out + (rs.sum(0),)
    return out


def colsum(dy):
    _n()
    return dy.float().sum(0)


def patchify(inp, *, B, C, OH, OW, kh, kw, stride_h, stride_w, sb, sc, sh, sw, Kpad):
    _n()
    flat = inp.reshape(-1) if inp.is_contiguous() else None
    base = inp.as_strided((B, C, (OH - 1) * stride_h + kh, (OW - 1) * stride_w + kw), (sb, sc, sh, sw))
    cols = F.unfold(base.float(), (kh, kw), stride=(stride_h, stride_w)).transpose(1, 2).reshape(B * OH * OW, C * kh * kw)
    out = torch.zeros((B * OH * OW, Kpad), dtype=BF16)
    out[:, : C * kh * kw] = cols.to(BF16)
    return out


def assemble_tokens(tok, cls, pos, *, B, L_, D):
    _n()
    t = tok.float().reshape(B, L_, D)
    if cls is not None:
        t = torch.cat([cls.float().view(1, 1, D).expand(B, 1, D), t], 1)
    if pos is not None:
        t = t + pos.float()
    return t.reshape(-1, D).to(BF16)


def assemble_tokens_bwd(dx, *, B, L_, D, has_cls, want_tok=True, want_pos=True, want_cls=True):
    _n()
    Lo = L_ + int(has_cls)
    d = dx.float().reshape(B, Lo, D)
    dtok = d[:, int(has_cls):].reshape(B * L_, D).to(BF16) if want_tok else None
    dpos = d.sum(0) if want_pos else None
    dcls = d[:, 0].sum(0) if (want_cls and has_cls) else None
    return dtok, dpos, dcls


def embed_tokens(ids, table, pos):
    _n()
    return (table[ids] + pos).reshape(-1, table.shape[1]).to(BF16)


def embed_tokens_bwd(ids, dx, *, vocab, want_table=True, want_pos=True):
    _n()
    ctx, D = ids.shape[-1], dx.shape[1]
    d = dx.float()
    dt = torch.zeros(vocab, D).index_add_(0, ids.reshape(-1), d) if want_table else None
    dp = d.reshape(-1, ctx, D).sum(0) if want_pos else None
    return dt, dp


def l2norm_fwd(x, eps=1e-12):
    _n()
    inv = 1.0 / x.norm(dim=-1).clamp_min(eps)
    return x * inv[:, None], inv


def l2norm_bwd(dy, y, inv):
    _n()
    return (dy - y * (dy * y).sum(-1, keepdim=True)) * inv[:, None]


def geglu_fwd(h):
    _n()
    a, g = h.float().chunk(2, -1)
    return (a * F.gelu(g)).to(BF16)


def geglu_bwd(h, dout):
    _n()
    a, g = h.float().chunk(2, -1)
    d = dout.float()
    return torch.cat([d * F.gelu(g), d * a * _act_grad(g, False)], -1).to(BF16)


def cast_bf16(x):
    _n()
    return x.contiguous().to(BF16)


def add_bf16(a, b):
    _n()
    return (a.float() + b.float()).to(BF16)


def rowlse(p16, q16, *, alpha, label_off=0, mask=None):
    _n()
    alpha = float(alpha)
    z = alpha * (p16.float() @ q16.float().t())
    if mask is not None:
        z = z * (mask != 0)
    lse = torch.logsumexp(z, -1)
    M = z.shape[0]
    diag = z[torch.arange(M), torch.arange(M) + label_off]
    return lse, (lse - diag).sum().reshape(1)


def clipgrad(p16, q16, *, alpha, row_lse, col_lse, label_off, gscale, gscale_dev=None, ds_row_only=False, mask=None):
    _n()
    alpha = float(alpha)
    if gscale_dev is not None:
        gscale = gscale * float(gscale_dev)
    acc = p16.float() @ q16.float().t()
    z = alpha * acc
    keep = None if mask is None else (mask != 0).float()
    if keep is not None:
        z = z * keep
    M, N = z.shape
    g = torch.exp(z - row_lse[:, None])
    grow = g
    k = 1.0
    if col_lse is not None:
        g = g + torch.exp(z - col_lse[None, :])
        k = 2.0
    onehot = torch.zeros(M, N)
    onehot[torch.arange(M), torch.arange(M) + label_off] = 1.0
    g = gscale * (g - k * onehot)
    grow = gscale * (grow - onehot)
    if keep is not None:
        g, grow = g * keep, grow * keep
    ds = (grow * acc).sum() if ds_row_only else (g * acc).sum()
    return g.to(BF16), ds.reshape(1)


# ----------------------------------------------------------------------------- point-cloud tokenizer
def fps(pts, start, npoint):
    _n()
    B, N, _ = pts.shape
    idx = torch.zeros(B, npoint, dtype=torch.long)
    dist = torch.full((B, N), 1e10)
    far = start.clone()
    bi = torch.arange(B)
    for i in range(npoint):
        idx[:, i] = far
        c = pts[bi, far].view(B, 1, 3)
        dist = torch.minimum(dist, ((pts - c) ** 2).sum(-1))
        far = dist.max(-1)[1]
    centers = torch.gather(pts, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).reshape(B * npoint, 3)
    return idx, centers


def clip_backward_fusable(E, q16):
    return E % 64 == 0


def clip_backward(p16, q16, *, alpha, row_lse, col_lse, label_off, gscale, gscale_dev=None, ds_row_only=False, mask=None, want_ds=True):
    """The fused one-launch backward (vl_clip_backward): g rounded to bf16 as the kernel does, then dX = alpha * g @ Q."""
    g, ds = clipgrad(p16, q16, alpha=alpha, row_lse=row_lse, col_lse=col_lse, label_off=label_off, gscale=gscale, gscale_dev=gscale_dev,
                     ds_row_only=ds_row_only, mask=mask)
    return float(alpha) * (g.float() @ q16.float()), ds


def knn_group(pts, centers, G, k, want_idx=False):
    _n()
    B, N, _ = pts.shape
    c = centers.reshape(B, G, 3)
    d = -2 * c @ pts.transpose(1, 2) + (c ** 2).sum(-1, keepdim=True) + (pts ** 2).sum(-1).unsqueeze(1)
    idx = torch.topk(d, k, dim=-1, largest=False)[1]
    nb = torch.gather(pts.unsqueeze(1).expand(-1, G, -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, 3)) - c.unsqueeze(2)
    nb = nb.reshape(B * G * k, 3)
    return (nb, idx.reshape(-1)) if want_idx else nb


def linear3(x, w, scale, shift, act, want_pre=False):
    _n()
    pre = (x @ w.t()) * scale + shift
    v = pre
    if act == 1:
        v = torch.relu(v)
    elif act == 2:
        v = F.gelu(v)
    return (v.to(BF16), pre.to(BF16)) if want_pre else v.to(BF16)


def group_max_bwd(dout, arg, G):
    _n()
    groups, C = dout.shape
    dx = torch.zeros(groups, G, C)
    dx.scatter_(1, arg.long().unsqueeze(1), dout.float().unsqueeze(1))
    return dx.reshape(groups * G, C).to(BF16)


def group_sum(x, G):
    _n()
    rows, C = x.shape
    return x.float().reshape(rows // G, G, C).sum(1).to(BF16)


def colsum2(a, b):
    _n()
    return a.float().sum(0), (a.float() * b.float()).sum(0)


def moments3(x):
    x = x.float()
    return torch.cat([x.sum(0), (x.t() @ x).reshape(9)])


def col_affine(a, p0, p2, *, b=None, p1=None, relu=False):
    y = a.float() * p0.float() + p2.float()
    if b is not None:
        y = y + b.float() * p1.float()
    return (torch.relu(y) if relu else y).to(torch.bfloat16)


def wgrad3(dy, x):
    _n()
    return dy.float().t() @ x


def group_max(x, G, want_arg=False):
    _n()
    rows, C = x.shape
    v, a = x.float().reshape(rows // G, G, C).max(dim=1)
    return (v.to(BF16), a.to(torch.int32)) if want_arg else v.to(BF16)


def gemm_grouped_residual_relu(a, b, gp, group, *, bias=None, relu=True):
    _n()
    acc = a.float() @ b.float().t()
    if bias is not None:
        acc = acc + bias
    acc = acc + gp.float().repeat_interleave(group, dim=0)
    return (torch.relu(acc) if relu else acc).to(BF16)


# ----------------------------------------------------------------------------- zero-shot evaluation
def template_mean(x, n_templates, transpose_out=False):
    _n()
    G, E = x.shape[0] // n_templates, x.shape[1]
    v = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    m = v.reshape(G, n_templates, E).mean(1)
    m = m / m.norm(dim=-1, keepdim=True)
    return m.t().contiguous() if transpose_out else m


def topk_rows(scores, k, want_values=False):
    _n()
    # descending value, ties -> smaller column: a stable sort of the negated scores
    order = torch.sort(-scores, dim=1, stable=True)[1][:, :k].to(torch.int32)
    if order.shape[1] < k:
        order = torch.cat([order, torch.full((scores.shape[0], k - order.shape[1]), -1, dtype=torch.int32)], 1)
    return (order, torch.gather(scores, 1, order.long().clamp_min(0))) if want_values else order


def average_precision(scores, targets, apply_sigmoid=True):
    _n()
    s = torch.sigmoid(scores) if apply_sigmoid else scores
    N, C = s.shape
    ap = torch.zeros(C)
    npos = (targets > 0.5).sum(0).to(torch.int32)
    for c in range(C):
        pos = (targets[:, c] > 0.5).nonzero().flatten()
        if pos.numel() == 0:
            continue
        ge = s[:, c][None, :] >= s[pos, c][:, None]
        ap[c] = (((ge & (targets[:, c] > 0.5)[None, :]).sum(1).double() / ge.sum(1).double()).sum() / pos.numel()).float()
    return ap, npos


def similarity(feats, gallery_t=None, gallery=None):
    _n()
    a = feats.to(BF16).float()
    g = gallery.to(BF16).float().t() if gallery is not None else gallery_t.to(BF16).float()
    return a @ g


def geglu_fusable(M, F):
    return M >= 512 and F % 128 == 0


def geglu_permute_rows(w):
    F = w.shape[0] // 2
    idx = torch.arange(F).view(F // 128, 128)
    return w.index_select(0, torch.cat([idx, idx + F], dim=1).reshape(-1)).contiguous()


def gemm_geglu(a, w_perm16, bias_perm):
    _n()
    F = w_perm16.shape[0] // 2
    inv = torch.empty(2 * F, dtype=torch.long)
    idx = torch.arange(F).view(F // 128, 128)
    inv[torch.cat([idx, idx + F], dim=1).reshape(-1)] = torch.arange(2 * F)
    acc = a.float() @ w_perm16.float().t() + bias_perm
    h = acc[:, inv]  # back to the original (value | gate) column order
    return (h[:, :F] * F_gelu(h[:, F:])).to(BF16), h.to(BF16)


def F_gelu(x):
    return F.gelu(x)
