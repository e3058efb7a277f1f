"""Tensor-level front end of the C ABI: derives (M, N, K, ld...) from torch views, allocates
outputs, and calls vitlens_b200.lib (the ctypes binding of libvitlens_b200.so).

Every function here launches hand-written sm_100a kernels; torch only provides memory.
2-D arguments are row-major views with unit inner stride (``t.stride(-1) == 1``); the row
stride is the leading dimension.  ``tests/emu_ops.py`` mirrors this module's signatures in
plain torch so the host logic (engine.py) can be exercised on a CPU-only box -- that emulation
lives in tests/ and is never importable from the product package.
"""
from __future__ import annotations

from typing import Optional, Tuple

import os

import torch

from . import lib as L

BF16, F32 = torch.bfloat16, torch.float32
EPI_LINEAR, EPI_GELU, EPI_RESIDUAL, EPI_GELU_BWD = L.EPI_LINEAR, L.EPI_GELU, L.EPI_RESIDUAL, L.EPI_GELU_BWD


def _v2(t: torch.Tensor, dtype=None) -> torch.Tensor:
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), f"need a row-major 2-D view, got {tuple(t.shape)} {t.stride()}"
    assert dtype is None or t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    assert t.is_cuda, "vitlens_b200 kernels need CUDA tensors (there is no CPU path)"
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


class PeerRows:
    """A bf16 matrix [world * rows, E] whose row block p lives in rank p's peer arena (vitlens_b200.comm): the all-gathered
    features of the contrastive loss, never materialised -- the GEMMs load each tile from its owner over NVLink."""

    __slots__ = ("addrs", "rows", "E", "flags", "ticket", "local", "arena")

    def __init__(self, addrs, rows, E, flags, ticket, local, arena=None):
        self.addrs, self.rows, self.E, self.flags, self.ticket, self.local, self.arena = list(addrs), rows, E, flags, ticket, local, arena

    @property
    def shape(self):
        return (len(self.addrs) * self.rows, self.E)

    @property
    def device(self):
        return self.local.device

    def kw(self, wait=True):
        return dict(b_peers=self.addrs, b_peer_rows=self.rows, peer_flags=self.flags if wait else 0, peer_flag_value=self.ticket)


def peer_rows_ok(rows: int) -> bool:
    """Row blocks the peer GEMMs can address in place (a 256-column logits tile must not straddle two ranks)."""
    return rows % 256 == 0


def pick_split_k(M: int, N: int, K: int) -> int:
    """Split the reduction of a weight-gradient GEMM when its (M, N) grid alone cannot fill the GPU.

    Cost model fitted to the B200 sweep in profiles/r01_splitk_sweep.log: time ~ waves(s) / s for the main loops plus a
    per-extra-split charge for the fp32 atomics that merge the partial tiles (it grows with the output size).  The kernel
    choice mirrors vl_gemm_bf16: CTA pairs (256 x 256 tiles, 74 clusters) when M >= 512 and N > 128, else 148 single CTAs."""
    kb = (K + 63) // 64
    if M >= 512 and N > 128:
        tiles, slots = ((M + 255) // 256) * ((N + 255) // 256), 74
    else:
        tiles, slots = ((M + 127) // 128) * ((N + 255) // 256 if N > 128 else 1), 148
    charge = 0.0012 * (M * N / 65536.0)
    best, best_cost = 1, None
    for s in range(1, max(1, min(16, kb // 8)) + 1):
        cost = -(-tiles * s // slots) / s + charge * (s - 1)
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = s, cost
    return best


def rowsum_fusable(M: int, N: int) -> bool:
    """True when a weight-gradient GEMM of this shape runs on the CTA-pair kernel, which can also emit the row sums of A
    (= the bias gradient) from the tiles it already holds (VlGemmArgs.rowsum_out)."""
    return M >= 512 and N > 128


def gemm(a, b, *, a_t=False, b_t=False, bias=None, epilogue=EPI_LINEAR, aux_in=None, want_aux_out=False, out=None,
         out_dtype=BF16, alpha=1.0, accumulate=False, split_k=None, act_quick=False, alpha_dev=None, want_rowsum=False):
    """out[M,N] = epilogue(alpha * A @ B^T); A = a ([M,K]) or a^T when a_t (a is [K,M]); B = b ([N,K]) or b^T when b_t.
    want_rowsum (fp32 LINEAR outputs, rowsum_fusable shapes): also returns sum_k A[m, k] as fp32 [M]."""
    _v2(a, BF16)
    peer = b if isinstance(b, PeerRows) else None
    if peer is None:
        _v2(b, BF16)
    M, K = (a.shape[1], a.shape[0]) if a_t else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_t else (b.shape[0], b.shape[1])
    assert K == Kb, f"K mismatch {K} vs {Kb}"
    fresh = out is None
    aux_out = torch.empty((M, N), device=a.device, dtype=BF16) if want_aux_out else None
    if aux_in is not None:
        _v2(aux_in, BF16)
        assert tuple(aux_in.shape) == (M, N)
    ldaux = _ld(aux_in) if aux_in is not None else (N if want_aux_out else 0)
    if split_k is None:
        split_k = pick_split_k(M, N, K) if (accumulate and (out_dtype if fresh else out.dtype) == F32 and epilogue == EPI_LINEAR) else 1
    if fresh:
        if accumulate and split_k == 1:
            accumulate = False  # a single split writes every element exactly once: plain stores, no zero fill, no atomics
        out = (torch.zeros if accumulate else torch.empty)((M, N), device=a.device, dtype=out_dtype)
    _v2(out)
    assert tuple(out.shape) == (M, N)
    rowsum = None
    if want_rowsum:
        assert rowsum_fusable(M, N) and out.dtype == F32 and epilogue == EPI_LINEAR and not want_aux_out
        rowsum = torch.empty((M,), device=a.device, dtype=F32)
    if peer is not None:
        assert not accumulate and not want_rowsum and split_k == 1
        L.gemm(a, None, out, M=M, N=N, K=K, lda=_ld(a), ldb=peer.E, ldd=_ld(out), a_mn=a_t, b_mn=b_t, epilogue=epilogue, bias=bias,
               aux_in=aux_in, aux_out=aux_out, ldaux=ldaux, alpha=alpha, act_quick=act_quick, alpha_dev=alpha_dev, **peer.kw())
    else:
        L.gemm(a, b, out, M=M, N=N, K=K, lda=_ld(a), ldb=_ld(b), ldd=_ld(out), a_mn=a_t, b_mn=b_t, epilogue=epilogue, bias=bias,
               aux_in=aux_in, aux_out=aux_out, ldaux=ldaux, alpha=alpha, accumulate=accumulate, split_k=split_k, act_quick=act_quick,
               alpha_dev=alpha_dev, rowsum_out=rowsum)
    if want_rowsum:
        return out, rowsum
    return (out, aux_out) if want_aux_out else out


def attention_fwd(q, k, v, *, B, H, nq, nk, causal=False, scale=None):
    _v2(q, BF16), _v2(k, BF16), _v2(v, BF16)
    assert q.shape == (B * nq, H * 64) and k.shape == (B * nk, H * 64) and v.shape == (B * nk, H * 64)
    scale = 64 ** -0.5 if scale is None else scale
    o = torch.empty((B * nq, H * 64), device=q.device, dtype=BF16)
    lse = torch.empty((B, H, nq), device=q.device, dtype=F32)
    L.attention_fwd(q, k, v, o, lse, B=B, H=H, nq=nq, nk=nk, ldq=_ld(q), ldk=_ld(k), ldv=_ld(v), ldo=H * 64, scale=scale, causal=causal)
    return o, lse


def attention_bwd(q, k, v, o, dout, lse, dq, dk, dv, *, B, H, nq, nk, causal=False, scale=None):
    for t in (q, k, v, o, dout, dq, dk, dv):
        _v2(t, BF16)
    scale = 64 ** -0.5 if scale is None else scale
    L.attention_bwd(q, k, v, o, dout, lse, dq, dk, dv, B=B, H=H, nq=nq, nk=nk, ldq=_ld(q), ldk=_ld(k), ldv=_ld(v), ldo=_ld(o),
                    lddo=_ld(dout), lddq=_ld(dq), lddk=_ld(dk), lddv=_ld(dv), scale=scale, causal=causal)


def layernorm_fwd(x, w, b, *, row_index=None, eps=1e-5, want_stats=True):
    _v2(x, BF16)
    T = x.shape[0] if row_index is None else row_index.numel()
    D = x.shape[1]
    y = torch.empty((T, D), device=x.device, dtype=BF16)
    mean = torch.empty((T,), device=x.device, dtype=F32) if want_stats else None
    rstd = torch.empty((T,), device=x.device, dtype=F32) if want_stats else None
    L.layernorm_fwd(x, w, b, y, mean, rstd, T=T, D=D, ldx=_ld(x), ldy=D, row_index=row_index, eps=eps)
    return y, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd, *, dres=None, row_index=None, want_wgrad=True, want_dres_sum=False):
    """Returns (dx, dw, db) or (dx, dw, db, colsum(dres)) with want_dres_sum.  With row_index, dx has x's full row
    count and is zero outside the gathered rows."""
    _v2(dy, BF16), _v2(x, BF16)
    T, D = dy.shape
    if row_index is None:
        dx = torch.empty((x.shape[0], D), device=x.device, dtype=BF16)
    else:
        dx = torch.zeros((x.shape[0], D), device=x.device, dtype=BF16)
    # parameter-gradient outputs are written (not accumulated) by a deterministic two-stage reduction: no zero fill
    dw = torch.empty((D,), device=x.device, dtype=F32) if want_wgrad else None
    db = torch.empty((D,), device=x.device, dtype=F32) if want_wgrad else None
    if dres is not None:
        _v2(dres, BF16)
    rsum = torch.empty((D,), device=x.device, dtype=F32) if want_dres_sum else None
    L.layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db, T=T, D=D, lddy=_ld(dy), ldx=_ld(x), lddx=D, dres=dres,
                    lddres=_ld(dres) if dres is not None else 0, row_index=row_index, dres_sum=rsum)
    return (dx, dw, db, rsum) if want_dres_sum else (dx, dw, db)


def colsum(dy):
    _v2(dy, BF16)
    T, N = dy.shape
    db = torch.empty((N,), device=dy.device, dtype=F32)
    L.colsum(dy, db, T=T, N=N, ld=_ld(dy))
    return db


def patchify(inp, *, B, C, OH, OW, kh, kw, stride_h, stride_w, sb, sc, sh, sw, Kpad):
    out = torch.empty((B * OH * OW, Kpad), device=inp.device, dtype=BF16)
    L.patchify(inp, out, B=B, C=C, OH=OH, OW=OW, kh=kh, kw=kw, stride_h=stride_h, stride_w=stride_w, sb=sb, sc=sc, sh=sh, sw=sw, Kpad=Kpad)
    return out


def assemble_tokens(tok, cls, pos, *, B, L_, D):
    """tok: contiguous [B*L_, D] bf16 -> [B*(L_+has_cls), D]"""
    has_cls = cls is not None
    out = torch.empty((B * (L_ + int(has_cls)), D), device=tok.device, dtype=BF16)
    L.assemble_tokens(tok, cls, pos, out, B=B, L=L_, D=D, has_cls=has_cls)
    return out


def assemble_tokens_bwd(dx, *, B, L_, D, has_cls, want_tok=True, want_pos=True, want_cls=True):
    dev = dx.device
    dtok = torch.empty((B * L_, D), device=dev, dtype=BF16) if want_tok else None
    dpos = torch.empty((L_ + int(has_cls), D), device=dev, dtype=F32) if want_pos else None
    dcls = torch.empty((D,), device=dev, dtype=F32) if (want_cls and has_cls) else None
    L.assemble_tokens_bwd(dx, dtok, dpos, dcls, B=B, L=L_, D=D, has_cls=has_cls)
    return dtok, dpos, dcls


def embed_tokens(ids, table, pos):
    rows, ctx, D = ids.numel(), ids.shape[-1], table.shape[1]
    out = torch.empty((rows, D), device=table.device, dtype=BF16)
    L.embed_tokens(ids.contiguous(), table, pos, out, rows=rows, ctx=ctx, D=D)
    return out


def embed_tokens_bwd(ids, dx, *, vocab, want_table=True, want_pos=True):
    rows, ctx, D = ids.numel(), ids.shape[-1], dx.shape[1]
    dtable = torch.zeros((vocab, D), device=dx.device, dtype=F32) if want_table else None
    dpos = torch.zeros((ctx, D), device=dx.device, dtype=F32) if want_pos else None
    L.embed_tokens_bwd(ids.contiguous(), dx, dtable, dpos, rows=rows, ctx=ctx, D=D)
    return dtable, dpos


def l2norm_fwd(x, eps=1e-12):
    B, E = x.shape
    y = torch.empty_like(x)
    inv = torch.empty((B,), device=x.device, dtype=F32)
    L.l2norm_fwd(x, y, inv, B=B, E=E, eps=eps)
    return y, inv


def l2norm_bwd(dy, y, inv):
    dx = torch.empty_like(y)
    L.l2norm_bwd(dy.contiguous(), y, inv, dx, B=y.shape[0], E=y.shape[1])
    return dx


def geglu_fusable(M: int, F: int) -> bool:
    """True when Linear(d, 2F) + GEGLU can run as ONE GEMM with the VL_EPI_GEGLU epilogue (CTA-pair kernel, whole tiles)."""
    return M >= 512 and F % 128 == 0


def geglu_permute_rows(w):
    """Row permutation of the FeedForward's first Linear for the fused epilogue: every 256-row tile = 128 value rows followed by
    the 128 gate rows of the same output columns.  w: [2F, ...] (weight) or [2F] (bias)."""
    F2 = w.shape[0]
    F = F2 // 2
    idx = torch.arange(F, device=w.device).view(F // 128, 128)
    perm = torch.cat([idx, idx + F], dim=1).reshape(-1)
    return w.index_select(0, perm).contiguous()


def gemm_geglu(a, w_perm16, bias_perm):
    """(value * gelu(gate) [M, F], h [M, 2F] pre-activations in the ORIGINAL column order) of a @ W^T + b with W / b given in the
    geglu_permute_rows layout -- reference perceiver.py:85-102 in one launch."""
    _v2(a, BF16), _v2(w_perm16, BF16)
    M, K = a.shape
    N = w_perm16.shape[0]
    assert geglu_fusable(M, N // 2) and w_perm16.shape[1] == K
    out = torch.empty((M, N // 2), device=a.device, dtype=BF16)
    h = torch.empty((M, N), device=a.device, dtype=BF16)
    L.gemm(a, w_perm16, out, M=M, N=N, K=K, lda=_ld(a), ldb=_ld(w_perm16), ldd=N // 2, epilogue=L.EPI_GEGLU, bias=bias_perm, aux_out=h, ldaux=N)
    return out, h


def geglu_fwd(h):
    M, F2 = h.shape
    out = torch.empty((M, F2 // 2), device=h.device, dtype=BF16)
    L.geglu_fwd(h, out, M=M, F=F2 // 2)
    return out


def geglu_bwd(h, dout):
    dh = torch.empty_like(h)
    L.geglu_bwd(h, dout, dh, M=h.shape[0], F=h.shape[1] // 2)
    return dh


def cast_bf16(x, out=None):
    """fp32 -> bf16 copy (weights for the tensor cores, dfeatures for the head GEMMs); `out`: a contiguous bf16 destination
    (e.g. a slot of the peer arena)."""
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    assert out.dtype == BF16 and out.is_contiguous() and out.numel() == x.numel()
    L.cast_f32_bf16(x, out)
    return out


def add_bf16(a, b):
    out = torch.empty_like(a)
    L.add_bf16(a, b, out)
    return out


def _mask8(mask, M, N):
    if mask is None:
        return None
    assert mask.dtype == torch.uint8 and tuple(mask.shape) == (M, N) and mask.stride(1) == 1 and mask.is_cuda
    return mask


def rowlse(p16, q16, *, alpha, label_off=0, mask=None):
    """Row-wise log-sum-exp of alpha * P @ Q^T without materialising the logits (alpha: 1-element fp32 device tensor); `mask`
    (uint8 [M, N]): logits where it is 0 are replaced by 0 (`logits * mask`, loss.py:540-575).
    Returns (lse [M] fp32, sum_i(lse_i - z[i, i + label_off]) as a 1-element fp32 tensor)."""
    _v2(p16, BF16)
    peer = q16 if isinstance(q16, PeerRows) else None
    if peer is None:
        _v2(q16, BF16)
    M, E = p16.shape
    N = q16.shape[0]
    nparts = L.rowlse_parts(N)
    dev = p16.device
    pm = torch.full((M, nparts), float("-inf"), device=dev, dtype=F32)
    ps = torch.zeros((M, nparts), device=dev, dtype=F32)
    diag = torch.zeros((M,), device=dev, dtype=F32)
    L.gemm(p16, None if peer else q16, None, M=M, N=N, K=E, lda=_ld(p16), ldb=peer.E if peer else _ld(q16), ldd=0, epilogue=L.EPI_ROWLSE,
           alpha=1.0, alpha_dev=alpha, out_vec0=pm, out_vec1=ps, out_vec2=diag, iparam=label_off, mask=_mask8(mask, M, N),
           **(peer.kw() if peer else {}))
    lse = torch.empty((M,), device=dev, dtype=F32)
    loss_sum = torch.empty((1,), device=dev, dtype=F32)
    L.lse_combine(pm, ps, diag, lse, loss_sum, M=M, nparts=nparts)
    return lse, loss_sum


def clipgrad(p16, q16, *, alpha, row_lse, col_lse, label_off, gscale, gscale_dev=None, ds_row_only=False, mask=None):
    """g[M,N] (bf16) = gscale * (exp(z - row_lse_i) + [col_lse] exp(z - col_lse_j) - k * onehot(j == i + label_off)),
    z = alpha * P @ Q^T;  also returns sum(g * P@Q^T) (d loss / d alpha) as a 1-element fp32 tensor -- with ds_row_only the sum
    runs over the row term gscale * (exp(z - row_lse_i) - onehot) alone."""
    _v2(p16, BF16)
    peer = q16 if isinstance(q16, PeerRows) else None
    if peer is None:
        _v2(q16, BF16)
    M, E = p16.shape
    N = q16.shape[0]
    N8 = (N + 7) // 8 * 8
    g = torch.zeros((M, N8), device=p16.device, dtype=BF16)
    ds = torch.empty((1,), device=p16.device, dtype=F32)
    L.gemm(p16, None if peer else q16, g, M=M, N=N, K=E, lda=_ld(p16), ldb=peer.E if peer else _ld(q16), ldd=N8, epilogue=L.EPI_CLIPGRAD,
           alpha=1.0, alpha_dev=alpha, row_vec=row_lse, col_vec=col_lse, iparam=label_off, fparam=gscale, fparam_dev=gscale_dev, scalar_out=ds,
           loss_flags=1 if ds_row_only else 0, mask=_mask8(mask, M, N), **(peer.kw(wait=False) if peer else {}))
    return g[:, :N], ds


def clip_backward_fusable(E: int, q16) -> bool:
    """True when vl_clip_backward can take this direction of the loss in one launch: embed dim a multiple of 64 and, with the
    features sharded over peer arenas, row blocks that a 128-row block never straddles.  VL_LOSS_BWD=unfused forces the two-GEMM
    path (g through HBM) for A/B checks."""
    if os.environ.get("VL_LOSS_BWD", "") == "unfused" or E % 64 != 0:
        return False
    if isinstance(q16, PeerRows):
        return len(q16.addrs) == 1 or q16.rows % 128 == 0
    return True


def clip_backward(p16, q16, *, alpha, row_lse, col_lse, label_off, gscale, gscale_dev=None, ds_row_only=False, mask=None, want_ds=True):
    """One direction of the contrastive backward in ONE launch: dX [M, E] fp32 = alpha * g @ Q with
    g = gscale * (exp(z - row_lse_i) + [col_lse] exp(z - col_lse_j) - k * onehot(j == i + label_off)), z = alpha * P @ Q^T (masked
    entries carry no gradient), and sum(g * P@Q^T) as a 1-element tensor -- what clipgrad() + gemm() compute through an HBM copy of
    g.  q16: bf16 [N, E] or PeerRows (every rank's rows read in place over NVLink)."""
    _v2(p16, BF16)
    peer = q16 if isinstance(q16, PeerRows) else None
    if peer is None:
        _v2(q16, BF16)
    M, E = p16.shape
    N = q16.shape[0]
    dx = torch.empty((M, E), device=p16.device, dtype=F32)
    ds = torch.empty((1,), device=p16.device, dtype=F32) if want_ds else None
    L.clip_backward(p16, None if peer else q16, dx, M=M, N=N, E=E, ldx=_ld(p16), ldy=peer.E if peer else _ld(q16), row_lse=row_lse, col_lse=col_lse,
                    label_off=label_off, alpha_dev=alpha, gscale=gscale, gscale_dev=gscale_dev, ds_out=ds, ds_row_only=ds_row_only,
                    mask=_mask8(mask, M, N), y_peers=peer.addrs if peer else None, y_peer_rows=peer.rows if peer else 0)
    return dx, ds


# ----------------------------------------------------------------------------- point-cloud tokenizer
def fps(pts, start, npoint):
    """pts [B,N,3] fp32, start [B] int64 -> (idx [B,npoint] int64, centers [B*npoint, 3] fp32)"""
    B, N, _ = pts.shape
    pts = pts.contiguous()
    idx = torch.empty((B, npoint), device=pts.device, dtype=torch.int64)
    centers = torch.empty((B * npoint, 3), device=pts.device, dtype=F32)
    L.fps(pts, start.contiguous(), idx, centers, B=B, N=N, npoint=npoint)
    return idx, centers


def knn_group(pts, centers, G, k, want_idx=False):
    """-> neighbourhoods minus centre [B*G*k, 3] fp32 (and the point indices [B*G*k] int64)"""
    B, N, _ = pts.shape
    nb = torch.empty((B * G * k, 3), device=pts.device, dtype=F32)
    idx = torch.empty((B * G * k,), device=pts.device, dtype=torch.int64) if want_idx else None
    L.knn_group(pts.contiguous(), centers, nb, idx, B=B, N=N, G=G, k=k)
    return (nb, idx) if want_idx else nb


def linear3(x, w, scale, shift, act, want_pre=False):
    """act(x[R,3] @ w[C,3]^T * scale + shift) -> bf16 [R,C] (and the pre-activation); act: 0 none, 1 relu, 2 gelu"""
    R, C = x.shape[0], w.shape[0]
    out = torch.empty((R, C), device=x.device, dtype=BF16)
    pre = torch.empty((R, C), device=x.device, dtype=BF16) if want_pre else None
    L.linear3(x.contiguous(), w.contiguous(), scale.contiguous(), shift.contiguous(), out, R=R, C=C, act=act, pre_out=pre)
    return (out, pre) if want_pre else out


def group_max_bwd(dout, arg, G):
    _v2(dout, BF16)
    groups, C = dout.shape
    dx = torch.empty((groups * G, C), device=dout.device, dtype=BF16)
    L.group_max_bwd(dout.contiguous(), arg, dx, groups=groups, G=G, C=C)
    return dx


def group_sum(x, G):
    _v2(x, BF16)
    rows, C = x.shape
    out = torch.empty((rows // G, C), device=x.device, dtype=BF16)
    L.group_sum(x, out, groups=rows // G, G=G, C=C)
    return out


def colsum2(a, b):
    """(sum_t a[t,:], sum_t a[t,:]*b[t,:]) in fp32"""
    _v2(a, BF16), _v2(b, BF16)
    T, N = a.shape
    s1 = torch.empty((N,), device=a.device, dtype=F32)
    s2 = torch.empty((N,), device=a.device, dtype=F32)
    L.colsum2(a.contiguous(), b.contiguous(), s1, s2, T=T, N=N)
    return s1, s2


def moments3(x):
    """[sum x (3) | sum x x^T (9)] over the rows of x[R,3] fp32 -> fp32 [12]"""
    assert x.is_cuda and x.dtype == F32 and x.dim() == 2 and x.shape[1] == 3
    out = torch.empty((12,), device=x.device, dtype=F32)
    L.moments3(x.contiguous(), out, R=x.shape[0])
    return out


def col_affine(a, p0, p2, *, b=None, p1=None, relu=False):
    """act(p0[c] * a + p1[c] * b + p2[c]) per column, bf16 [R,C] (b / p1 optional)"""
    _v2(a, BF16)
    if b is not None:
        _v2(b, BF16)
        assert b.shape == a.shape and p1 is not None
    R, C = a.shape
    out = torch.empty((R, C), device=a.device, dtype=BF16)
    f = lambda t: None if t is None else t.to(F32).contiguous()  # noqa: E731
    L.col_affine(a.contiguous(), None if b is None else b.contiguous(), f(p0), f(p1), f(p2), out, R=R, C=C, act=1 if relu else 0)
    return out


def wgrad3(dy, x):
    """dy[R,C]^T @ x[R,3] -> fp32 [C,3]"""
    _v2(dy, BF16)
    R, C = dy.shape
    dw = torch.empty((C, 3), device=dy.device, dtype=F32)
    L.wgrad3(dy.contiguous(), x.contiguous(), dw, R=R, C=C)
    return dw


def group_max(x, G, want_arg=False):
    _v2(x, BF16)
    rows, C = x.shape
    out = torch.empty((rows // G, C), device=x.device, dtype=BF16)
    arg = torch.empty((rows // G, C), device=x.device, dtype=torch.int32) if want_arg else None
    L.group_max(x, out, arg, groups=rows // G, G=G, C=C)
    return (out, arg) if want_arg else out


def gemm_grouped_residual_relu(a, b, gp, group, *, bias=None, relu=True):
    """relu(a @ b^T + gp[row // group]) -- second_conv.0 on cat(global, local) with folded BatchNorm (dvae.py:206-208);
    relu=False: the pre-BatchNorm conv output (training mode normalises it with batch statistics afterwards)."""
    _v2(a, BF16), _v2(b, BF16), _v2(gp, BF16)
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=BF16)
    L.gemm(a, b, out, M=M, N=N, K=K, lda=_ld(a), ldb=_ld(b), ldd=N, epilogue=L.EPI_RESIDUAL, bias=bias, aux_in=gp, ldaux=_ld(gp),
           aux_row_div=group, relu=relu)
    return out


# ----------------------------------------------------------------------------- zero-shot evaluation
def template_mean(x, n_templates, transpose_out=False):
    """normalize(mean over templates of normalize(x)) per class: x fp32 [G*T, E] -> [G, E] (or [E, G] with transpose_out)."""
    assert x.is_cuda and x.dtype == F32 and x.dim() == 2 and x.shape[0] % n_templates == 0
    G, E = x.shape[0] // n_templates, x.shape[1]
    out = torch.empty((E, G) if transpose_out else (G, E), device=x.device, dtype=F32)
    L.template_mean(x.contiguous(), out, G=G, T=n_templates, E=E, ldo=out.shape[1], transpose_out=transpose_out)
    return out


def topk_rows(scores, k, want_values=False):
    """Row-wise top-k of fp32 scores [R, C] -> int32 indices [R, k] (ties -> smaller column) and optionally the values."""
    assert scores.is_cuda and scores.dtype == F32 and scores.dim() == 2 and scores.stride(1) == 1
    R, Cn = scores.shape
    idx = torch.empty((R, k), device=scores.device, dtype=torch.int32)
    val = torch.empty((R, k), device=scores.device, dtype=F32) if want_values else None
    L.topk_rows(scores, idx, val, ld=scores.stride(0), rows=R, cols=Cn, k=k)
    return (idx, val) if want_values else idx


def average_precision(scores, targets, apply_sigmoid=True):
    """Per-class average precision (sklearn definition): scores, targets fp32 [N, C] -> (ap [C] fp32, positives [C] int32)."""
    assert scores.is_cuda and scores.dtype == F32 and targets.dtype == F32 and scores.shape == targets.shape
    N, Cn = scores.shape
    ap = torch.empty((Cn,), device=scores.device, dtype=F32)
    npos = torch.empty((Cn,), device=scores.device, dtype=torch.int32)
    L.average_precision(scores.contiguous(), targets.contiguous(), ap, npos, lds=Cn, ldt=Cn, N=N, C=Cn, apply_sigmoid=apply_sigmoid)
    return ap, npos


def similarity(feats, gallery_t=None, gallery=None):
    """fp32 [B, C] similarities of fp32 features [B, E] with a gallery given as [E, C] (classifier layout) or [C, E] -- the
    tcgen05 GEMM with bf16-rounded operands and fp32 accumulation / output.  (The gallery is zero-padded to a multiple of 8
    entries for the GEMM; the result is the [B, C] view of the padded output.)"""
    a = cast_bf16(feats)
    g = cast_bf16(gallery if gallery is not None else gallery_t.t().contiguous())  # [C, E]
    Cn = g.shape[0]
    C8 = (Cn + 7) // 8 * 8
    if C8 != Cn:
        g = torch.cat([g, g.new_zeros((C8 - Cn, g.shape[1]))]).contiguous()
    return gemm(a, g, out_dtype=F32)[:, :Cn]
