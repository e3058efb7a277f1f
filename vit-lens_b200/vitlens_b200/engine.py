"""Autograd glue between the reference-compatible nn.Modules (open_clip/*) and the sm_100a kernels.

Each ``torch.autograd.Function`` below covers one fused stage of the hot path, runs only
vitlens_b200 kernels (through ``ops``) in forward and backward, owns exactly the activations
it needs, and skips weight-gradient work for frozen parameters (the ViT-Lens recipes freeze
the ViT and train the Lens: transformer.py:553-627).  Activations and the residual stream are
bf16, LayerNorm/softmax statistics and parameter gradients fp32.

`_ops` is the only gateway to arithmetic; tests may swap it for a torch emulation to exercise
this file on a CPU-only box (tests/emu_ops.py).  There is no fallback in the product.
"""
from __future__ import annotations

from typing import Optional

import weakref

import torch

from . import ops as _ops
from .ops import PeerRows as _PeerRows

BF16, F32 = torch.bfloat16, torch.float32


# ----------------------------------------------------------------------------- bf16 weight cache
class _WeightCache:
    """bf16 copies of fp32 master weights, refreshed when the parameter changes (its autograd
    version counter moves on every in-place optimizer update).

    An entry belongs to one live parameter OBJECT: it holds a weak reference to it and dies with it.  (Keying on id() alone
    is not enough: a freed parameter's id, storage address, version and shape can all be reused by the next model built in
    the same process, which would hand that model the previous model's weights.)"""

    def __init__(self):
        self._c = {}

    def _alive(self, hit, ps) -> bool:
        return hit is not None and len(hit[2]) == len(ps) and all(r() is q for r, q in zip(hit[2], ps))

    def _put(self, key, ver, t, ps):
        drop = lambda _ref, key=key, c=self._c: c.pop(key, None)  # noqa: E731  (entry dies with its parameter)
        self._c[key] = (ver, t, tuple(weakref.ref(q, drop) for q in ps))

    def get(self, p: torch.Tensor, tag: str = "", make=None) -> torch.Tensor:
        key = (id(p), tag)
        ver = (p._version, p.data_ptr(), tuple(p.shape))
        hit = self._c.get(key)
        if self._alive(hit, (p,)) and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            t = make(p) if make is not None else _ops.cast_bf16(p.detach().reshape(p.shape[0], -1) if p.dim() > 1 else p.detach())
        self._put(key, ver, t, (p,))
        return t

    def get_cat(self, tag: str, ps) -> torch.Tensor:
        """bf16 concat of several weights along dim 0 (one cached tensor per tuple of parameters)."""
        key = (tuple(id(q) for q in ps), tag)
        ver = tuple((q._version, q.data_ptr(), tuple(q.shape)) for q in ps)
        hit = self._c.get(key)
        if self._alive(hit, ps) and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            t = torch.cat([self.get(q) for q in ps], dim=0).contiguous()
        self._put(key, ver, t, tuple(ps))
        return t

    def clear(self):
        self._c.clear()

    def plain_copy(self, p):
        """The cached plain bf16 copy of parameter `p` if it is current (used by the fused optimizer)."""
        hit = self._c.get((id(p), ""))
        if self._alive(hit, (p,)) and hit[0] == (p._version, p.data_ptr(), tuple(p.shape)):
            return hit[1]
        return None

    def clear_derived(self):
        for k in [k for k in self._c if not (isinstance(k, tuple) and len(k) == 2 and k[1] == "" and isinstance(k[0], int))]:
            del self._c[k]


WEIGHTS = _WeightCache()


def w16(p):
    return WEIGHTS.get(p)


def _cat16(tag, *ps):
    """bf16 concat of several weights along dim 0 (Lens to_q | to_kv -> one QKV GEMM)."""
    return WEIGHTS.get_cat(tag, ps)


def _conv_w16(p, kpad):
    def make(w):
        o = w.shape[0]
        flat = w.detach().reshape(o, -1)
        if flat.shape[1] != kpad:
            pad = torch.zeros((o, kpad), device=w.device, dtype=F32)
            pad[:, : flat.shape[1]] = flat
            flat = pad
        return _ops.cast_bf16(flat)

    return WEIGHTS.get(p, f"conv{kpad}", make)


def _wgrad(dy, x):
    """dW[out, in] = dy^T x  (both operands MN-major: the reduction runs over tokens)."""
    return _ops.gemm(dy, x, a_t=True, b_t=True, out_dtype=F32, accumulate=True)


def _wgrad_bias(dy, x, want_w=True, want_b=True):
    """(dW, db) of a Linear: dW = dy^T x and db = column sums of dy.  When the weight-gradient GEMM runs on the CTA-pair
    kernel the bias gradient rides on it (row sums of its A operand, VlGemmArgs.rowsum_out) instead of a second pass over dy."""
    if want_w and want_b and _ops.rowsum_fusable(dy.shape[1], x.shape[1]):
        return _ops.gemm(dy, x, a_t=True, b_t=True, out_dtype=F32, accumulate=True, want_rowsum=True)
    return (_wgrad(dy, x) if want_w else None), (_ops.colsum(dy) if want_b else None)


def _dgrad(dy, w, **kw):
    """dx[T, in] = dy[T, out] @ W[out, in]  (W consumed in place as an MN-major B operand)."""
    return _ops.gemm(dy, w, b_t=True, **kw)


def _need(ctx, i):
    return ctx.needs_input_grad[i]


# ----------------------------------------------------------------------------- ResidualAttentionBlock
class VitBlockFn(torch.autograd.Function):
    """ResidualAttentionBlock.forward (transformer.py:254-272):
    x + out_proj(MHSA(ln_1(x))) then + c_proj(act(c_fc(ln_2(.)))), as 4 GEMMs with fused
    bias / residual / GELU epilogues, 2 LayerNorm kernels and 1 attention kernel."""

    @staticmethod
    def forward(ctx, x, ln1w, ln1b, inw, inb, outw, outb, ln2w, ln2b, fcw, fcb, pjw, pjb, B, N, H, causal, quick):
        D = x.shape[1]
        train_w = any(ctx.needs_input_grad[1:13])
        need_bwd = train_w or ctx.needs_input_grad[0]
        xn1, m1, r1 = _ops.layernorm_fwd(x, ln1w, ln1b, want_stats=need_bwd)
        qkv = _ops.gemm(xn1, w16(inw), bias=inb)
        o, lse = _ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B=B, H=H, nq=N, nk=N, causal=causal)
        x1 = _ops.gemm(o, w16(outw), bias=outb, epilogue=_ops.EPI_RESIDUAL, aux_in=x)
        xn2, m2, r2 = _ops.layernorm_fwd(x1, ln2w, ln2b, want_stats=need_bwd)
        h, u = _ops.gemm(xn2, w16(fcw), bias=fcb, epilogue=_ops.EPI_GELU, want_aux_out=True, act_quick=quick)
        y = _ops.gemm(h, w16(pjw), bias=pjb, epilogue=_ops.EPI_RESIDUAL, aux_in=x1)
        if need_bwd:
            ctx.cfg = (B, N, H, causal, quick, train_w)
            keep_w = (xn1, xn2, h) if train_w else (None, None, None)
            ctx.save_for_backward(x, m1, r1, qkv, o, lse, x1, m2, r2, u, ln1w, inw, outw, ln2w, fcw, pjw, *keep_w)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, N, H, causal, quick, train_w = ctx.cfg
        x, m1, r1, qkv, o, lse, x1, m2, r2, u, ln1w, inw, outw, ln2w, fcw, pjw, xn1, xn2, h = ctx.saved_tensors
        D = x.shape[1]
        dy = dy.contiguous()
        g = [None] * 18
        # ---- MLP
        g[11], g[12] = _wgrad_bias(dy, h, _need(ctx, 11), _need(ctx, 12))
        du = _dgrad(dy, w16(pjw), epilogue=_ops.EPI_GELU_BWD, aux_in=u, act_quick=quick)
        g[9], g[10] = _wgrad_bias(du, xn2, _need(ctx, 9), _need(ctx, 10))
        dxn2 = _dgrad(du, w16(fcw))
        want_ln2 = _need(ctx, 7) or _need(ctx, 8)
        dx1, g7, g8 = _ops.layernorm_bwd(dxn2, x1, ln2w, m2, r2, dres=dy, want_wgrad=want_ln2)
        g[7], g[8] = (g7, g8) if want_ln2 else (None, None)
        # ---- attention
        g[5], g[6] = _wgrad_bias(dx1, o, _need(ctx, 5), _need(ctx, 6))
        do = _dgrad(dx1, w16(outw))
        dqkv = torch.empty_like(qkv)
        _ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                           B=B, H=H, nq=N, nk=N, causal=causal)
        g[3], g[4] = _wgrad_bias(dqkv, xn1, _need(ctx, 3), _need(ctx, 4))
        if _need(ctx, 0) or _need(ctx, 1) or _need(ctx, 2):
            dxn1 = _dgrad(dqkv, w16(inw))
            want_ln1 = _need(ctx, 1) or _need(ctx, 2)
            dx, g1, g2 = _ops.layernorm_bwd(dxn1, x, ln1w, m1, r1, dres=dx1, want_wgrad=want_ln1)
            g[0] = dx
            g[1], g[2] = (g1, g2) if want_ln1 else (None, None)
        return tuple(g)


# ----------------------------------------------------------------------------- stand-alone LayerNorm (ln_pre / ln_post / ln_final / gathers)
class LayerNormFn(torch.autograd.Function):
    """F.layer_norm on bf16 rows; with row_index it is the fused `ln(x[rows])` used for cls / EOT
    pooling (transformer.py:653-657,783; model.py:537-540)."""

    @staticmethod
    def forward(ctx, x, w, b, row_index):
        need = any(ctx.needs_input_grad[:3])
        y, m, r = _ops.layernorm_fwd(x, w, b, row_index=row_index, want_stats=need)
        if need:
            ctx.save_for_backward(x, w, m, r, row_index if row_index is not None else torch.empty(0))
            ctx.has_index = row_index is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, m, r, idx = ctx.saved_tensors
        want_w = _need(ctx, 1) or _need(ctx, 2)
        dx, dw, db = _ops.layernorm_bwd(dy.contiguous(), x, w, m, r, row_index=idx if ctx.has_index else None, want_wgrad=want_w)
        return (dx if _need(ctx, 0) else None), (dw if _need(ctx, 1) else None), (db if _need(ctx, 2) else None), None


# ----------------------------------------------------------------------------- patch embed (conv1 as GEMM)
class PatchEmbedFn(torch.autograd.Function):
    """Bias-free strided Conv2d -> [B*L, width] tokens (transformer.py:464-470,674-676;
    AST_tokenizer.py:44-52; DepthTokenizer.py:50-53): a gather kernel builds the bf16 patch matrix,
    the tcgen05 GEMM multiplies it with the flattened kernel.  `geom` = dict of vl_patchify args."""

    @staticmethod
    def forward(ctx, inp, weight, geom):
        K = weight[0].numel()
        kpad = (K + 7) // 8 * 8
        cols = _ops.patchify(inp, Kpad=kpad, **geom)
        tok = _ops.gemm(cols, _conv_w16(weight, kpad))
        if ctx.needs_input_grad[1]:
            ctx.save_for_backward(cols)
            ctx.wshape = tuple(weight.shape)
        return tok

    @staticmethod
    def backward(ctx, dtok):
        (cols,) = ctx.saved_tensors
        dw = _wgrad(dtok.contiguous(), cols)  # [width, Kpad]
        O = ctx.wshape[0]
        K = 1
        for s in ctx.wshape[1:]:
            K *= s
        return None, dw[:, :K].reshape(ctx.wshape), None


class PatchEmbedBiasFn(torch.autograd.Function):
    """Conv1d-with-bias patch embed of the EEG tokenizer (modal_eeg/models/EEG_tokenizer.py:16-21,35-36)."""

    @staticmethod
    def forward(ctx, inp, weight, bias, geom):
        K = weight[0].numel()
        kpad = (K + 7) // 8 * 8
        cols = _ops.patchify(inp, Kpad=kpad, **geom)
        tok = _ops.gemm(cols, _conv_w16(weight, kpad), bias=bias)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            ctx.save_for_backward(cols)
            ctx.wshape = tuple(weight.shape)
        return tok

    @staticmethod
    def backward(ctx, dtok):
        (cols,) = ctx.saved_tensors
        dtok = dtok.contiguous()
        K = 1
        for s in ctx.wshape[1:]:
            K *= s
        dw = _wgrad(dtok, cols)[:, :K].reshape(ctx.wshape) if _need(ctx, 1) else None
        db = _ops.colsum(dtok) if _need(ctx, 2) else None
        return None, dw, db, None


# ----------------------------------------------------------------------------- cls / positional assembly
class AssembleFn(torch.autograd.Function):
    """cat(cls, tokens) + positional_embedding (transformer.py:756-768) or tokens + pos
    (adapter output, transformer.py:743) in one pass.  tok: [B*L, D] bf16 -> [B*(L+has_cls), D]."""

    @staticmethod
    def forward(ctx, tok, cls, pos, B, L):
        D = tok.shape[1]
        ctx.cfg = (B, L, D, cls is not None)
        return _ops.assemble_tokens(tok.contiguous(), cls, pos, B=B, L_=L, D=D)

    @staticmethod
    def backward(ctx, dx):
        B, L, D, has_cls = ctx.cfg
        dtok, dpos, dcls = _ops.assemble_tokens_bwd(dx.contiguous(), B=B, L_=L, D=D, has_cls=has_cls, want_tok=_need(ctx, 0),
                                                    want_pos=_need(ctx, 2), want_cls=_need(ctx, 1))
        return dtok, dcls, dpos, None, None


class BroadcastRowsFn(torch.autograd.Function):
    """repeat(latents, 'n d -> b n d') (perceiver.py:316) as a bf16 broadcast; backward sums over the batch."""

    @staticmethod
    def forward(ctx, latents, B):
        n, D = latents.shape
        ctx.cfg = (B, n, D)
        zero = torch.zeros((B * n, D), device=latents.device, dtype=BF16)
        return _ops.assemble_tokens(zero, None, latents, B=B, L_=n, D=D)

    @staticmethod
    def backward(ctx, dx):
        B, n, D = ctx.cfg
        _, dpos, _ = _ops.assemble_tokens_bwd(dx.contiguous(), B=B, L_=n, D=D, has_cls=False, want_tok=False, want_pos=True, want_cls=False)
        return dpos, None


# ----------------------------------------------------------------------------- projection head
class ProjFn(torch.autograd.Function):
    """`pooled @ proj` (transformer.py:786-787; text_projection model.py:540): bf16 rows x fp32 [D, E]
    parameter -> fp32 features."""

    @staticmethod
    def forward(ctx, pooled, proj):
        ctx.save_for_backward(pooled, proj)
        return _ops.gemm(pooled, w16(proj), b_t=True, out_dtype=F32)

    @staticmethod
    def backward(ctx, dfeat):
        pooled, proj = ctx.saved_tensors
        d16 = _ops.cast_bf16(dfeat)
        dpooled = _ops.gemm(d16, w16(proj)) if _need(ctx, 0) else None
        dproj = _ops.gemm(pooled, d16, a_t=True, b_t=True, out_dtype=F32, accumulate=True) if _need(ctx, 1) else None
        return dpooled, dproj


class L2NormFn(torch.autograd.Function):
    """F.normalize(dim=-1) on fp32 features (model.py:295,307,522,526,540)."""

    @staticmethod
    def forward(ctx, x):
        y, inv = _ops.l2norm_fwd(x.contiguous())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        return _ops.l2norm_bwd(dy.contiguous(), y, inv)


class TextEmbedFn(torch.autograd.Function):
    """token_embedding(text) + positional_embedding (model.py:530-532)."""

    @staticmethod
    def forward(ctx, ids, table, pos):
        ctx.save_for_backward(ids)
        ctx.vocab = table.shape[0]
        return _ops.embed_tokens(ids, table, pos)

    @staticmethod
    def backward(ctx, dx):
        (ids,) = ctx.saved_tensors
        dt, dp = _ops.embed_tokens_bwd(ids, dx.contiguous(), vocab=ctx.vocab, want_table=_need(ctx, 1), want_pos=_need(ctx, 2))
        return None, dt, dp


# ----------------------------------------------------------------------------- the Lens (Perceiver) stages
class LensAttnFn(torch.autograd.Function):
    """PreNorm(Attention) + residual (perceiver.py:67-82,105-154,319/323).  `ctx_rows is None` ->
    latent self-attention (context = normed x); otherwise cross-attention over `data` with its own
    norm_context.  to_q / to_kv are bias-free, to_out has a bias."""

    @staticmethod
    def forward(ctx, x, data, nw, nb, cw, cb, wq, wkv, wo, bo, B, n, M, heads):
        cross = data is not None
        inner = wq.shape[0]
        need = any(ctx.needs_input_grad)
        xn, mx, rx = _ops.layernorm_fwd(x, nw, nb, want_stats=need)
        if cross:
            cn, mc, rc = _ops.layernorm_fwd(data, cw, cb, want_stats=need)
            q = _ops.gemm(xn, w16(wq))
            kv = _ops.gemm(cn, w16(wkv))
            k, v = kv[:, :inner], kv[:, inner:]
        else:
            cn = mc = rc = None
            qkv = _ops.gemm(xn, _cat16("qkv", wq, wkv))
            q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
            kv = qkv
        o, lse = _ops.attention_fwd(q, k, v, B=B, H=heads, nq=n, nk=M if cross else n)
        y = _ops.gemm(o, w16(wo), bias=bo, epilogue=_ops.EPI_RESIDUAL, aux_in=x)
        if need:
            ctx.cfg = (cross, B, n, M, heads, inner)
            dummy = torch.empty(0, device=x.device)
            ctx.save_for_backward(x, xn, mx, rx, q if cross else dummy, kv, o, lse, nw, wq, wkv, wo,
                                  *( (data, cn, mc, rc, cw) if cross else (dummy,) * 5))
        return y

    @staticmethod
    def backward(ctx, dy):
        cross, B, n, M, heads, inner = ctx.cfg
        x, xn, mx, rx, q, kv, o, lse, nw, wq, wkv, wo, data, cn, mc, rc, cw = ctx.saved_tensors
        dy = dy.contiguous()
        g = [None] * 14
        if _need(ctx, 8):
            g[8] = _wgrad(dy, o)
        if _need(ctx, 9):
            g[9] = _ops.colsum(dy)
        do = _dgrad(dy, w16(wo))
        if cross:
            dq = torch.empty_like(q)
            dkv = torch.empty_like(kv)
            _ops.attention_bwd(q, kv[:, :inner], kv[:, inner:], o, do, lse, dq, dkv[:, :inner], dkv[:, inner:], B=B, H=heads, nq=n, nk=M)
            if _need(ctx, 6):
                g[6] = _wgrad(dq, xn)
            if _need(ctx, 7):
                g[7] = _wgrad(dkv, cn)
            dxn = _dgrad(dq, w16(wq))
            if _need(ctx, 1) or _need(ctx, 4) or _need(ctx, 5):
                dcn = _dgrad(dkv, w16(wkv))
                want = _need(ctx, 4) or _need(ctx, 5)
                ddata, g4, g5 = _ops.layernorm_bwd(dcn, data, cw, mc, rc, want_wgrad=want)
                g[1] = ddata if _need(ctx, 1) else None
                g[4], g[5] = (g4, g5) if want else (None, None)
        else:
            qkv = kv
            dqkv = torch.empty_like(qkv)
            _ops.attention_bwd(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], o, do, lse, dqkv[:, :inner],
                               dqkv[:, inner:2 * inner], dqkv[:, 2 * inner:], B=B, H=heads, nq=n, nk=n)
            if _need(ctx, 6) or _need(ctx, 7):
                dw = _wgrad(dqkv, xn)  # [3*inner, D]
                g[6], g[7] = dw[:inner], dw[inner:]
            dxn = _dgrad(dqkv, _cat16("qkv", wq, wkv))
        want = _need(ctx, 2) or _need(ctx, 3)
        dx, g2, g3 = _ops.layernorm_bwd(dxn, x, nw, mx, rx, dres=dy, want_wgrad=want)
        g[0] = dx if _need(ctx, 0) else None
        g[2], g[3] = (g2, g3) if want else (None, None)
        return tuple(g)


class LensFFFn(torch.autograd.Function):
    """PreNorm(FeedForward) + residual (perceiver.py:85-102,320/324): Linear(d, 8d) -> GEGLU -> Linear(4d, d)."""

    @staticmethod
    def forward(ctx, x, nw, nb, w0, b0, w2, b2):
        need = any(ctx.needs_input_grad)
        xn, m, r = _ops.layernorm_fwd(x, nw, nb, want_stats=need)
        if _ops.geglu_fusable(xn.shape[0], w0.shape[0] // 2):
            # one GEMM: the epilogue pairs every value column with its gate column (weight rows permuted once per weight version)
            wp = WEIGHTS.get(w0, "geglu", lambda w: _ops.cast_bf16(_ops.geglu_permute_rows(w.detach())))
            bp = WEIGHTS.get(b0, "geglu", lambda b: _ops.geglu_permute_rows(b.detach().float()))
            gg, h = _ops.gemm_geglu(xn, wp, bp)
        else:
            h = _ops.gemm(xn, w16(w0), bias=b0)
            gg = _ops.geglu_fwd(h)
        y = _ops.gemm(gg, w16(w2), bias=b2, epilogue=_ops.EPI_RESIDUAL, aux_in=x)
        if need:
            ctx.save_for_backward(x, xn, m, r, h, gg, nw, w0, w2)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, xn, m, r, h, gg, nw, w0, w2 = ctx.saved_tensors
        dy = dy.contiguous()
        g = [None] * 7
        if _need(ctx, 5):
            g[5] = _wgrad(dy, gg)
        if _need(ctx, 6):
            g[6] = _ops.colsum(dy)
        dg = _dgrad(dy, w16(w2))
        dh = _ops.geglu_bwd(h, dg)
        if _need(ctx, 3):
            g[3] = _wgrad(dh, xn)
        if _need(ctx, 4):
            g[4] = _ops.colsum(dh)
        dxn = _dgrad(dh, w16(w0))
        want = _need(ctx, 1) or _need(ctx, 2)
        dx, g1, g2 = _ops.layernorm_bwd(dxn, x, nw, m, r, dres=dy, want_wgrad=want)
        g[0] = dx if _need(ctx, 0) else None
        g[1], g[2] = (g1, g2) if want else (None, None)
        return tuple(g)


# ----------------------------------------------------------------------------- contrastive loss
class ContrastiveFn(torch.autograd.Function):
    """One feature pair of ClipLoss / TriClipLoss (loss.py:116-163, 346-385), local rows only:

        value = [ sum_i CE_i(s * X_loc @ all_Y^T) + sum_i CE_i(s * Y_loc @ all_X^T) ] / (2 * val_rows)

    with labels = arange + label_off.  Two tcgen05 GEMMs with the row-LSE epilogue produce it
    without writing logits to HBM.  Backward re-runs the GEMMs with the gradient epilogue
        g = gscale * (softmax_row + [col_term] softmax_col - k * onehot)     (bf16 [B_loc x B_all] scratch)
    followed by dX = s * g @ all_Y.  softmax_col needs the LSE of every *column*, i.e. the row LSEs
    of the opposite direction from all ranks: `gather_lse` all-gathers a [B_loc] vector when
    world_size > 1.  col_term=False is the `local_loss=True, gather_with_grad=False` gradient.
    `ds_post` post-processes d(loss)/d(scale) (cross-rank sum where the reference computes the
    full matrix on every rank).  `ds_rows_only` (local_loss + gather_with_grad): the column term of g carries the OTHER
    ranks' losses (it is what all_gather's backward sends home), so d(this rank's loss)/d(scale) sums the row term alone."""

    @staticmethod
    def forward(ctx, x, y, all_x, all_y, scale, label_off, val_rows, grad_rows, col_term, gather_lse, ds_post, ds_rows_only=False, masks=None):
        x16, y16 = _ops.cast_bf16(x.detach()), _ops.cast_bf16(y.detach())
        # all_x / all_y: None (single process), a gathered [B_all, E] tensor, or ops.PeerRows -- the rows of rank p read in place
        # from p's peer arena inside the GEMMs (the all-gather of loss.py:55-76 fused into the logits kernel)
        peer = isinstance(all_x, _PeerRows)
        ax16 = all_x if peer else (x16 if all_x is None else _ops.cast_bf16(all_x.detach()))
        ay16 = all_y if peer else (y16 if all_y is None else _ops.cast_bf16(all_y.detach()))
        s = scale.detach().float().reshape(1).contiguous()  # stays on the device: no host sync in the step
        # masks = (mask_x [B_loc, B_all], mask_y [B_loc, B_all]) uint8 or None: `logits * mask` of the mask loss variants
        mkw_x = {} if masks is None else dict(mask=masks[0])
        mkw_y = {} if masks is None else dict(mask=masks[1])
        lse_x, sum_x = _ops.rowlse(x16, ay16, alpha=s, label_off=label_off, **mkw_x)
        lse_y, sum_y = _ops.rowlse(y16, ax16, alpha=s, label_off=label_off, **mkw_y)
        loss = (sum_x + sum_y) / (2.0 * val_rows)
        empty = torch.empty(0, device=x.device)
        col_x = col_y = empty
        if col_term:
            col_x = gather_lse(lse_y) if gather_lse is not None else lse_y  # columns of the x-direction = rows of the y-direction
            col_y = gather_lse(lse_x) if gather_lse is not None else lse_x
        ctx.peer = (ax16, ay16) if peer else None
        ctx.save_for_backward(x16, y16, empty if peer else ax16, empty if peer else ay16, lse_x, lse_y, col_x, col_y, s)
        ctx.cfg = (label_off, grad_rows, col_term, ds_post, ds_rows_only)
        ctx.masks = masks
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        x16, y16, ax16, ay16, lse_x, lse_y, col_x, col_y, s = ctx.saved_tensors
        label_off, grad_rows, col_term, ds_post, ds_rows_only = ctx.cfg
        if ctx.peer is not None:
            ax16, ay16 = ctx.peer
            if ax16.arena is not None and not ax16.arena.features_alive(ax16.ticket):
                raise RuntimeError("contrastive backward: the peer arena's feature ring has been overwritten since this loss's forward "
                                   "(more than two later loss forwards); run backward earlier or set VL_COMM=nccl")
        gs = 1.0 / (2.0 * grad_rows)
        dl = dloss.detach().float().reshape(1).contiguous()  # upstream gradient stays on the device too
        mkw_x = {} if ctx.masks is None else dict(mask=ctx.masks[0])
        mkw_y = {} if ctx.masks is None else dict(mask=ctx.masks[1])
        kw = dict(alpha=s, label_off=label_off, gscale=gs, gscale_dev=dl, ds_row_only=ds_rows_only)
        if _ops.clip_backward_fusable(x16.shape[1], ay16) and _ops.clip_backward_fusable(x16.shape[1], ax16):
            # one launch per direction: the gradient tile g = d loss / d logits never leaves the SM (vl_clip_backward)
            dx, ds_x = _ops.clip_backward(x16, ay16, row_lse=lse_x, col_lse=col_x if col_term else None, **kw, **mkw_x)
            dy, ds_y = _ops.clip_backward(y16, ax16, row_lse=lse_y, col_lse=col_y if col_term else None, **kw, **mkw_y)
            dx = dx if _need(ctx, 0) else None
            dy = dy if _need(ctx, 1) else None
        else:
            gx, ds_x = _ops.clipgrad(x16, ay16, row_lse=lse_x, col_lse=col_x if col_term else None, **kw, **mkw_x)
            gy, ds_y = _ops.clipgrad(y16, ax16, row_lse=lse_y, col_lse=col_y if col_term else None, **kw, **mkw_y)
            dx = _ops.gemm(gx, ay16, b_t=True, out_dtype=F32, alpha_dev=s) if _need(ctx, 0) else None
            dy = _ops.gemm(gy, ax16, b_t=True, out_dtype=F32, alpha_dev=s) if _need(ctx, 1) else None
        dscale = None
        if _need(ctx, 4):
            # with the column term every logit's gradient appears in both directions
            dscale = (ds_x + ds_y) * (0.5 if (col_term and not ds_rows_only) else 1.0)
            if ds_post is not None:
                dscale = ds_post(dscale)
            dscale = dscale.reshape(())
        return dx, dy, None, None, dscale, None, None, None, None, None, None, None, None


class AddFn(torch.autograd.Function):
    """bf16 a + b (token features + per-sample positional tokens, transformer.py:743)."""

    @staticmethod
    def forward(ctx, a, b):
        return _ops.add_bf16(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, d):
        return d, d


# ----------------------------------------------------------------------------- point-cloud tokenizer
def _bn_scale(bn_w, rv, eps):
    return torch.rsqrt(rv + eps) * bn_w


def _allreduce_sum(t, group):
    import torch.distributed as dist

    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class PointTokenizerFn(torch.autograd.Function):
    """PointTokenizer.forward (point_encoder.py:350-362): FPS -> kNN groups -> mini-PointNet (dvae.py:196-212) ->
    reduce_dim, plus pos = MLP(centres).  Returns (tokens, pos, bn_stats): tokens / pos bf16 [B*G, trans_dim].

    BatchNorm1d, eval mode (`train` False): running statistics, i.e. a per-channel affine folded into the neighbouring 1x1
    convolutions (its weight / bias still receive gradients).

    BatchNorm1d, training mode (`train` True; what model.train() gives the reference, dvae.py:185-193): batch statistics
    over all B*G*k positions.  first_conv.0 is linear in the 3-d neighbourhood points, so the mean / variance of its 128
    outputs come in closed form from the points' 3-vector mean and 3x3 second-moment matrix (vl_moments3) and the layer
    still runs as one fused conv+BN+ReLU kernel; its backward needs no per-element pass either (everything reduces to
    `wgrad3`, a column sum and the same moments).  second_conv.0's pre-BN output is materialised once (bf16), reduced with
    vl_colsum2 (sum, sum of squares), and normalised + ReLU'd by vl_col_affine; backward applies the batch-norm correction
    dz = s (dy - mean dy - xhat mean(dy xhat)) with the same kernel.  `bn_stats` = [mean0 | var0 | mean1 | var1 | n]
    (biased variances) feeds the running-statistics update.  `sync_group`: SyncBatchNorm (pc_tri_main.py:372-373) -- the
    sums are all-reduced over that process group in forward and backward."""

    @staticmethod
    def forward(ctx, pts, fps_start, w0, b0, g0, be0, rm0, rv0, w3, b3, w10, b10, g1, be1, rm1, rv1, w13, b13, wr, br, wp0, bp0, wp2, bp2,
                G, k, eps0, eps1, train=False, sync_group=None):
        need = any(ctx.needs_input_grad)
        _, centers = _ops.fps(pts, fps_start, G)
        nb = _ops.knn_group(pts, centers, G, k)
        W0 = w0.reshape(w0.shape[0], 3)
        w10m = w10.reshape(w10.shape[0], -1)
        half = w10m.shape[1] // 2
        stats = torch.empty(0, device=pts.device)
        if not train:
            s0 = _bn_scale(g0, rv0, eps0)
            t0 = (b0 - rm0) * s0 + be0
        else:
            mom = _ops.moments3(nb).double()
            cnt = torch.tensor([float(nb.shape[0])], device=pts.device, dtype=torch.float64)
            if sync_group is not None:
                packed = _allreduce_sum(torch.cat([mom, cnt]), sync_group)
                mom, cnt = packed[:12], packed[12:]
            m1 = mom[:3] / cnt
            M2 = mom[3:].view(3, 3) / cnt
            W0d, b0d = W0.double(), b0.double()
            mu0 = W0d @ m1 + b0d
            var0 = ((W0d @ (M2 - torch.outer(m1, m1))) * W0d).sum(1).clamp_min(0)
            rstd0 = torch.rsqrt(var0 + eps0)
            s0 = (g0.double() * rstd0).float()
            t0 = ((b0d - mu0) * g0.double() * rstd0 + be0.double()).float()
        f1 = _ops.linear3(nb, W0, s0, t0, 1)                                             # conv + BN + ReLU  [R,128]
        w3_16 = w16(w3)
        f2 = _ops.gemm(f1, w3_16, bias=b3)                                               # [R,256]
        g1f, arg1 = _ops.group_max(f2, k, want_arg=True)                                 # [BG,256]
        y1 = None
        if not train:
            s1 = _bn_scale(g1, rv1, eps1)
            t1 = (b10 - rm1) * s1 + be1
            w10f = w10m * s1[:, None]                                                    # BN scale folded into the conv rows
            wg = _ops.cast_bf16(w10f[:, :half].contiguous())
            wl = _ops.cast_bf16(w10f[:, half:].contiguous())
            gp = _ops.gemm(g1f, wg, bias=t1)                                             # global half + shift, per group
            f3 = _ops.gemm_grouped_residual_relu(f2, wl, gp, k)                          # [R,512]
        else:
            wg = _ops.cast_bf16(w10m[:, :half].contiguous())
            wl = _ops.cast_bf16(w10m[:, half:].contiguous())
            gp = _ops.gemm(g1f, wg, bias=b10)
            y1 = _ops.gemm_grouped_residual_relu(f2, wl, gp, k, relu=False)              # pre-BN conv output [R,512]
            sy, syy = _ops.colsum2(y1, y1)
            if sync_group is not None:
                packed = _allreduce_sum(torch.cat([sy, syy]), sync_group)
                sy, syy = packed[: sy.numel()], packed[sy.numel():]
            n = cnt.float()
            mu1 = sy / n
            var1 = (syy / n - mu1 * mu1).clamp_min(0)
            rstd1 = torch.rsqrt(var1 + eps1)
            s1 = g1 * rstd1
            f3 = _ops.col_affine(y1, s1, be1 - mu1 * s1, relu=True)                      # normalise + affine + ReLU
            stats = torch.cat([mu0.float(), var0.float(), mu1, var1, n])
        w13_16 = w16(w13)
        f4 = _ops.gemm(f3, w13_16, bias=b13)                                             # [R,enc]
        tokf, arg2 = _ops.group_max(f4, k, want_arg=True)                                # [BG,enc]
        tok = _ops.gemm(tokf, w16(wr), bias=br)
        ones = torch.ones_like(bp0)
        p1, up = _ops.linear3(centers, wp0, ones, bp0, 2, want_pre=True)                 # Linear(3,128) + GELU
        pos = _ops.gemm(p1, w16(wp2), bias=bp2)
        if need:
            ctx.cfg = (G, k, train, sync_group)
            dummy = torch.empty(0, device=pts.device)
            extra = (y1, mom.float(), cnt.float(), mu0.float(), rstd0.float(), mu1, rstd1, W0, b0) if train else (dummy,) * 9
            ctx.save_for_backward(nb, f1, f2, arg1, g1f, f3, arg2, tokf, centers, up, p1, wg, wl, s0, s1, g0, be0, g1, be1, w3, w13, wr, wp2, *extra)
        ctx.mark_non_differentiable(stats)
        return tok, pos, stats

    @staticmethod
    def backward(ctx, dtok, dpos, _dstats):
        G, k, train, sync_group = ctx.cfg
        (nb, f1, f2, arg1, g1f, f3, arg2, tokf, centers, up, p1, wg, wl, s0, s1, g0, be0, g1, be1, w3, w13, wr, wp2,
         y1, mom, cnt, mu0, rstd0, mu1, rstd1, W0, b0) = ctx.saved_tensors
        dtok, dpos = dtok.contiguous(), dpos.contiguous()
        # reduce_dim
        g_wr, g_br = _wgrad(dtok, tokf), _ops.colsum(dtok)
        dtokf = _dgrad(dtok, w16(wr))
        df4 = _ops.group_max_bwd(dtokf, arg2, k)
        # second_conv.3
        g_w13, g_b13 = _wgrad(df4, f3).unsqueeze(-1), _ops.colsum(df4)
        dy3 = _dgrad(df4, w16(w13), epilogue=_ops.EPI_GELU_BWD, aux_in=f3, act_quick=2)  # through the ReLU (f3 > 0)
        # second_conv.1 (BatchNorm) and second_conv.0 on cat(global, local)
        if not train:   # BatchNorm as a fixed affine: its scale rides in wg / wl
            sdy, sdya = _ops.colsum2(dy3, f3)
            g_be1 = sdy
            g_g1 = (sdya - be1 * sdy) / g1
            g_b10 = s1 * sdy
            dz, wscale = dy3, s1[:, None]
        else:           # batch statistics: dz = s (dy - mean dy - xhat mean(dy xhat)),  xhat = (y1 - mu) rstd
            sdy, sdyy = _ops.colsum2(dy3, y1)
            g_be1 = sdy                                                                  # gamma / beta: this rank's rows only (as torch's
            g_g1 = rstd1 * (sdyy - mu1 * sdy)                                            # SyncBatchNorm; DDP averages them afterwards)
            if sync_group is not None:
                packed = _allreduce_sum(torch.cat([sdy, sdyy]), sync_group)
                sdy, sdyy = packed[: sdy.numel()], packed[sdy.numel():]
            g_b10 = torch.zeros_like(sdy)                                                # a bias in front of a batch-norm has no gradient
            c1 = s1 * rstd1 * rstd1 * (sdyy - mu1 * sdy) / cnt
            dz = _ops.col_affine(dy3, s1, -(s1 * sdy / cnt - mu1 * c1), b=y1, p1=-c1)
            wscale = 1.0
        gs = _ops.group_sum(dz, k)
        g_w10 = (torch.cat([_wgrad(gs, g1f), _wgrad(dz, f2)], dim=1) * wscale).unsqueeze(-1)
        dg1 = _dgrad(gs, wg)
        df2 = _dgrad(dz, wl, epilogue=_ops.EPI_RESIDUAL, aux_in=_ops.group_max_bwd(dg1, arg1, k))
        # first_conv.3
        g_w3, g_b3 = _wgrad(df2, f1).unsqueeze(-1), _ops.colsum(df2)
        dy1 = _dgrad(df2, w16(w3), epilogue=_ops.EPI_GELU_BWD, aux_in=f1, act_quick=2)
        # first_conv.1 (BatchNorm) and first_conv.0 (3 -> 128)
        if not train:
            sdy, sdya = _ops.colsum2(dy1, f1)
            g_be0 = sdy
            g_g0 = (sdya - be0 * sdy) / g0
            g_b0 = s0 * sdy
            g_w0 = (_ops.wgrad3(dy1, nb) * s0[:, None]).unsqueeze(-1)
        else:
            # y0 = W0 x + b0 is never materialised: with A = sum_r dy_r x_r^T (wgrad3), sum dy and the points' moments,
            #   sum dy xhat = rstd ((W0 * A).sum(1) + (b0 - mu) sum dy)
            #   dW0 = s (A - mean(dy) sum x^T - mean(dy xhat) rstd (W0 sum x x^T + (b0 - mu) sum x^T))
            A = _ops.wgrad3(dy1, nb)
            sdy = _ops.colsum(dy1)
            sdyx = rstd0 * ((W0 * A).sum(1) + (b0 - mu0) * sdy)
            sx, sxx = mom[:3], mom[3:].view(3, 3)            # global sums under SyncBN ...
            n = cnt
            if sync_group is not None:                        # ... but dW0 sums over the local rows only
                loc = _ops.moments3(nb)
                sx_l, sxx_l = loc[:3], loc[3:].view(3, 3)
                packed = _allreduce_sum(torch.cat([sdy, sdyx]), sync_group)
                sdy_g, sdyx_g = packed[: sdy.numel()], packed[sdy.numel():]
            else:
                sx_l, sxx_l, sdy_g, sdyx_g = sx, sxx, sdy, sdyx
            g_be0 = sdy
            g_g0 = sdyx
            g_b0 = torch.zeros_like(sdy)
            corr = (sdy_g / n)[:, None] * sx_l[None, :] + (sdyx_g / n * rstd0)[:, None] * (W0 @ sxx_l + (b0 - mu0)[:, None] * sx_l[None, :])
            g_w0 = (s0[:, None] * (A - corr)).unsqueeze(-1)
        # pos_embed
        g_wp2, g_bp2 = _wgrad(dpos, p1), _ops.colsum(dpos)
        dup = _dgrad(dpos, w16(wp2), epilogue=_ops.EPI_GELU_BWD, aux_in=up)
        g_wp0, g_bp0 = _ops.wgrad3(dup, centers), _ops.colsum(dup)
        grads = (None, None, g_w0, g_b0, g_g0, g_be0, None, None, g_w3, g_b3, g_w10, g_b10, g_g1, g_be1, None, None, g_w13, g_b13, g_wr, g_br,
                 g_wp0, g_bp0, g_wp2, g_bp2, None, None, None, None, None, None)
        return tuple(g if (g is None or ctx.needs_input_grad[i]) else None for i, g in enumerate(grads))


def _bn_sync_group(bn):
    """The process group a converted SyncBatchNorm layer reduces over (None: plain BatchNorm or a single process)."""
    import torch.distributed as dist

    if not isinstance(bn, torch.nn.SyncBatchNorm) or not (dist.is_available() and dist.is_initialized()):
        return None
    group = bn.process_group if bn.process_group is not None else dist.group.WORLD
    return group if dist.get_world_size(group) > 1 else None


def _bn_update_running(bn, mean, var_biased, n):
    """nn.BatchNorm1d's running-statistics update (momentum None = cumulative average), unbiased variance."""
    with torch.no_grad():
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        if bn.running_mean is None:
            return
        m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        bn.running_mean.mul_(1 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
        bn.running_var.mul_(1 - m).add_((var_biased * (n / (n - 1).clamp_min(1))).to(bn.running_var.dtype), alpha=m)


def point_tokenizer_forward(tk, pts, fps_start):
    enc = tk.encoder
    c0, bn0, c3 = enc.first_conv[0], enc.first_conv[1], enc.first_conv[3]
    c10, bn1, c13 = enc.second_conv[0], enc.second_conv[1], enc.second_conv[3]
    # nn.BatchNorm1d semantics: batch statistics in training mode (or when the layer tracks no running statistics)
    train = bn0.training or bn0.running_mean is None
    assert train == (bn1.training or bn1.running_mean is None), "the two BatchNorm layers of the point tokenizer must be in the same mode"
    zeros = lambda bn: torch.zeros_like(bn.weight)  # noqa: E731
    tok, pos, stats = PointTokenizerFn.apply(
        pts, fps_start, c0.weight, c0.bias, bn0.weight, bn0.bias,
        bn0.running_mean if bn0.running_mean is not None else zeros(bn0), bn0.running_var if bn0.running_var is not None else zeros(bn0),
        c3.weight, c3.bias, c10.weight, c10.bias, bn1.weight, bn1.bias,
        bn1.running_mean if bn1.running_mean is not None else zeros(bn1), bn1.running_var if bn1.running_var is not None else zeros(bn1),
        c13.weight, c13.bias, tk.reduce_dim.weight, tk.reduce_dim.bias, tk.pos_embed[0].weight, tk.pos_embed[0].bias,
        tk.pos_embed[2].weight, tk.pos_embed[2].bias, tk.num_group, tk.group_size, bn0.eps, bn1.eps, train, _bn_sync_group(bn0) if train else None)
    if train and bn0.training:
        c0n, c1n = bn0.num_features, bn1.num_features
        n = stats[-1]
        _bn_update_running(bn0, stats[:c0n], stats[c0n:2 * c0n], n)
        _bn_update_running(bn1, stats[2 * c0n:2 * c0n + c1n], stats[2 * c0n + c1n:2 * c0n + 2 * c1n], n)
    return tok, pos
