"""CPU oracle for the ViT-Lens hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``vit-lens_b200/``) never imports it and never falls back to it.

This is a plain fp32 restatement, in elementary torch-CPU ops (matmul / softmax /
mean / erf ...), of the reference's algorithm for the path named by
``BASELINE.json.north_star``.  It is *functional*: every function takes the
reference's ``state_dict`` (same key names / shapes as the reference modules
own) plus input tensors.  Gradients come from torch autograd over
these elementary ops.

Parity pinning: the reference ships no tests/golden vectors for this path
 -- "parity unpinned by the reference's own tests".  Instead
``oracle/make_golden.py`` imports the real reference from ``/root/reference``
(in the build container), runs it next to this restatement on identical seeded
weights + inputs, asserts agreement, and commits the reference's outputs as
fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks
this file against those fixtures everywhere (no reference needed).

All ``file:line`` citations are relative to ``/root/reference/vitlens/src/open_clip``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- leaf math
def layer_norm(x, w, b, eps: float = 1e-5):
    """transformer.py:28-34 (LayerNorm), perceiver.py:71-72 (nn.LayerNorm): eps 1e-5, affine,
    biased variance over the last dim."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * w + b


def gelu(x):
    """nn.GELU() exact-erf form; model.py:130 selects it when quick_gelu=False."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def quick_gelu(x):
    """transformer.py:37-40."""
    return x * torch.sigmoid(1.702 * x)


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def l2_normalize(x, eps: float = 1e-12):
    """F.normalize(dim=-1) as called at model.py:295,307,522,526,540: x / max(||x||, eps)."""
    n = x.pow(2).sum(dim=-1, keepdim=True).sqrt().clamp_min(eps)
    return x / n


def mha_packed(x, in_w, in_b, out_w, out_b, heads: int, attn_mask=None):
    """nn.MultiheadAttention self-attention with packed in_proj (transformer.py:215,241-252).

    x: [B, N, D] (the reference runs LND; batch-first here is the same arithmetic).
    in_proj_weight rows are ordered q,k,v; q is scaled by hd**-0.5; additive mask."""
    B, N, D = x.shape
    hd = D // heads
    qkv = linear(x, in_w, in_b)
    q, k, v = qkv.split(D, dim=-1)

    def split(t):
        return t.reshape(B, N, heads, hd).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    s = (q * (hd ** -0.5)) @ k.transpose(-1, -2)
    if attn_mask is not None:
        s = s + attn_mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, N, D)
    return linear(o, out_w, out_b)


def resblock(sd: SD, pre: str, x, heads: int, act=gelu, attn_mask=None):
    """ResidualAttentionBlock.forward transformer.py:254-272 (ls_1/ls_2 are Identity for
    every shipped config: ls_init_value=None)."""
    h = layer_norm(x, sd[pre + "ln_1.weight"], sd[pre + "ln_1.bias"])
    x = x + mha_packed(
        h,
        sd[pre + "attn.in_proj_weight"],
        sd[pre + "attn.in_proj_bias"],
        sd[pre + "attn.out_proj.weight"],
        sd[pre + "attn.out_proj.bias"],
        heads,
        attn_mask,
    )
    h = layer_norm(x, sd[pre + "ln_2.weight"], sd[pre + "ln_2.bias"])
    h = act(linear(h, sd[pre + "mlp.c_fc.weight"], sd[pre + "mlp.c_fc.bias"]))
    x = x + linear(h, sd[pre + "mlp.c_proj.weight"], sd[pre + "mlp.c_proj.bias"])
    return x


def n_resblocks(sd: SD, pre: str) -> int:
    n = 0
    while f"{pre}resblocks.{n}.ln_1.weight" in sd:
        n += 1
    return n


def transformer(sd: SD, pre: str, x, heads: int, act=gelu, attn_mask=None, taps: Optional[dict] = None):
    """Transformer.forward transformer.py:364-371."""
    for i in range(n_resblocks(sd, pre)):
        x = resblock(sd, f"{pre}resblocks.{i}.", x, heads, act, attn_mask)
        if taps is not None:
            taps[f"{pre}resblocks.{i}"] = x
    return x


# --------------------------------------------------------------------------- patch / token embeds
def conv_patch_embed(x, w, stride: Tuple[int, int]):
    """Bias-free Conv2d as an explicit unfold + GEMM (transformer.py:464-470,674-676;
    AST_tokenizer.py:22-28; DepthTokenizer.py:22-28) followed by the reshape/permute to
    [B, L, width] (row-major over output (h, w))."""
    B, C, H, W = x.shape
    O, _, kh, kw = w.shape
    cols = torch.nn.functional.unfold(x, kernel_size=(kh, kw), stride=stride)  # [B, C*kh*kw, L]
    return cols.transpose(1, 2) @ w.reshape(O, -1).t()  # [B, L, O]


def audio_adapter(sd: SD, pre: str, x, fstride: int = 10, tstride: int = 10):
    """AST_tokenizer.forward AST_tokenizer.py:44-57: x [B, T, F] -> unsqueeze(1).transpose(2,3)
    -> conv (k=patch, stride (fstride, tstride)) -> [B, fdim*tdim, width]; returns (x, pos)."""
    img = x.unsqueeze(1).transpose(2, 3)
    tok = conv_patch_embed(img, sd[pre + "conv1.weight"], (fstride, tstride))
    return tok, sd[pre + "pos_emb"]


def depth_adapter(sd: SD, pre: str, x):
    """DepthTokenizer.forward DepthTokenizer.py:50-60 (input_patchnorm=False branch)."""
    w = sd[pre + "conv1.weight"]
    tok = conv_patch_embed(x, w, (w.shape[2], w.shape[3]))
    return tok, sd[pre + "pos_emb"]


def eeg_adapter(sd: SD, pre: str, x, stride: int):
    """PatchEmbed1D.forward modal_eeg/models/EEG_tokenizer.py:35-42: Conv1d(in_chans, width, k=window, stride) WITH bias over
    time as an explicit window gather + GEMM, transposed to [B, L, width]; returns (x, pos)."""
    w, b = sd[pre + "proj.weight"], sd[pre + "proj.bias"]  # [O, C, k], [O]
    O, C, k = w.shape
    win = x.unfold(2, k, stride)  # [B, C, L, k]
    cols = win.permute(0, 2, 1, 3).reshape(x.shape[0], win.shape[2], C * k)
    return cols @ w.reshape(O, C * k).t() + b, sd[pre + "pos_emb"]


def fps_indices(xyz, npoint: int, start: torch.Tensor):
    """misc.fps modal_3d/models/pointbert/misc.py:48-68 with the random start index
    (misc.py:60) passed in so the result is comparable."""
    B, N, _ = xyz.shape
    cent = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.full((B, N), 1e10, dtype=xyz.dtype)
    far = start.clone()
    bi = torch.arange(B)
    for i in range(npoint):
        cent[:, i] = far
        c = xyz[bi, far, :].view(B, 1, 3)
        d = ((xyz - c) ** 2).sum(-1)
        distance = torch.minimum(distance, d)
        far = distance.max(-1)[1]
    return cent


def batchnorm1d_eval(x, sd: SD, pre: str, eps: float = 1e-5):
    """nn.BatchNorm1d in eval mode on [B, C, L]."""
    m = sd[pre + "running_mean"].view(1, -1, 1)
    v = sd[pre + "running_var"].view(1, -1, 1)
    return (x - m) * torch.rsqrt(v + eps) * sd[pre + "weight"].view(1, -1, 1) + sd[pre + "bias"].view(1, -1, 1)


def batchnorm1d_train(x, sd: SD, pre: str, eps: float = 1e-5, momentum: float = 0.1, new_stats=None):
    """nn.BatchNorm1d in training mode on [B, C, L] (what model.train() gives dvae.py:185-193): biased batch statistics over
    (B, L); `new_stats[pre]` receives the updated (running_mean, running_var) -- unbiased variance, torch's momentum rule."""
    m = x.mean(dim=(0, 2))
    v = x.var(dim=(0, 2), unbiased=False)
    if new_stats is not None:
        n = x.shape[0] * x.shape[2]
        new_stats[pre] = ((1 - momentum) * sd[pre + "running_mean"] + momentum * m.detach(),
                          (1 - momentum) * sd[pre + "running_var"] + momentum * v.detach() * n / max(n - 1, 1))
    return (x - m.view(1, -1, 1)) * torch.rsqrt(v.view(1, -1, 1) + eps) * sd[pre + "weight"].view(1, -1, 1) + sd[pre + "bias"].view(1, -1, 1)


def point_adapter(sd: SD, pre: str, pts, fps_start, num_group: int = 512, group_size: int = 32, bn_train: bool = False, new_stats=None):
    """PointTokenizer.forward point_encoder.py:350-362 -> Group.forward dvae.py:150-176
    (fps, knn_point dvae.py:107-118, centre subtraction) -> Encoder.forward dvae.py:196-212
    (BatchNorm in eval mode, or with batch statistics when bn_train) -> reduce_dim; pos = MLP(centres)."""
    bn = (lambda x, q: batchnorm1d_train(x, sd, q, new_stats=new_stats)) if bn_train else (lambda x, q: batchnorm1d_eval(x, sd, q))
    B, N, _ = pts.shape
    cidx = fps_indices(pts, num_group, fps_start)
    center = torch.gather(pts, 1, cidx.unsqueeze(-1).expand(-1, -1, 3))
    d = -2 * center @ pts.transpose(1, 2)
    d = d + (center ** 2).sum(-1, keepdim=True) + (pts ** 2).sum(-1).unsqueeze(1)
    idx = torch.topk(d, group_size, dim=-1, largest=False, sorted=False)[1]
    nb = torch.gather(pts.unsqueeze(1).expand(-1, num_group, -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, 3))
    nb = nb - center.unsqueeze(2)
    g = nb.reshape(B * num_group, group_size, 3).transpose(1, 2)  # [BG, 3, n]
    e = pre + "encoder."
    f = torch.einsum("oc,bcn->bon", sd[e + "first_conv.0.weight"].squeeze(-1), g) + sd[e + "first_conv.0.bias"].view(1, -1, 1)
    f = torch.relu(bn(f, e + "first_conv.1."))
    f = torch.einsum("oc,bcn->bon", sd[e + "first_conv.3.weight"].squeeze(-1), f) + sd[e + "first_conv.3.bias"].view(1, -1, 1)
    fg = f.max(dim=2, keepdim=True)[0]
    f = torch.cat([fg.expand(-1, -1, group_size), f], dim=1)
    f = torch.einsum("oc,bcn->bon", sd[e + "second_conv.0.weight"].squeeze(-1), f) + sd[e + "second_conv.0.bias"].view(1, -1, 1)
    f = torch.relu(bn(f, e + "second_conv.1."))
    f = torch.einsum("oc,bcn->bon", sd[e + "second_conv.3.weight"].squeeze(-1), f) + sd[e + "second_conv.3.bias"].view(1, -1, 1)
    tok = f.max(dim=2)[0].reshape(B, num_group, -1)
    tok = linear(tok, sd[pre + "reduce_dim.weight"], sd[pre + "reduce_dim.bias"])
    pos = linear(gelu(linear(center, sd[pre + "pos_embed.0.weight"], sd[pre + "pos_embed.0.bias"])),
                 sd[pre + "pos_embed.2.weight"], sd[pre + "pos_embed.2.bias"])
    return tok, pos


# --------------------------------------------------------------------------- the Lens (Perceiver)
def lens_attention(sd: SD, pre: str, x, context, heads: int):
    """perceiver.Attention.forward perceiver.py:120-154 (non-xformers branch): bias-free to_q /
    to_kv, sim = q k^T * dim_head**-0.5, softmax, to_out with bias."""
    B, n, _ = x.shape
    q = linear(x, sd[pre + "to_q.weight"])
    kv = linear(context, sd[pre + "to_kv.weight"])
    k, v = kv.chunk(2, dim=-1)
    hd = q.shape[-1] // heads

    def split(t):
        return t.reshape(B, t.shape[1], heads, hd).permute(0, 2, 1, 3)

    q, k, v = split(q), split(k), split(v)
    p = torch.softmax((q @ k.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, n, heads * hd)
    return linear(o, sd[pre + "to_out.weight"], sd[pre + "to_out.bias"])


def lens_ff(sd: SD, pre: str, x):
    """FeedForward/GEGLU perceiver.py:85-102: Linear(d, 8d) -> chunk -> x * gelu(gates) -> Linear(4d, d)."""
    h = linear(x, sd[pre + "net.0.weight"], sd[pre + "net.0.bias"])
    a, g = h.chunk(2, dim=-1)
    return linear(a * gelu(g), sd[pre + "net.2.weight"], sd[pre + "net.2.bias"])


def perceiver(sd: SD, pre: str, data, cross_heads: int = 1, latent_heads: int = 16):
    """Perceiver.forward perceiver.py:289-332 with fourier_encode_data=False and
    return_embeddings=True (how transformer.py:753 calls it). PreNorm: perceiver.py:67-82."""
    B = data.shape[0]
    x = sd[pre + "latents"].unsqueeze(0).expand(B, -1, -1)
    d = 0
    while f"{pre}layers.{d}.0.norm.weight" in sd:
        lp = f"{pre}layers.{d}."
        xn = layer_norm(x, sd[lp + "0.norm.weight"], sd[lp + "0.norm.bias"])
        cn = layer_norm(data, sd[lp + "0.norm_context.weight"], sd[lp + "0.norm_context.bias"])
        x = lens_attention(sd, lp + "0.fn.", xn, cn, cross_heads) + x
        x = lens_ff(sd, lp + "1.fn.", layer_norm(x, sd[lp + "1.norm.weight"], sd[lp + "1.norm.bias"])) + x
        s = 0
        while f"{lp}2.{s}.0.norm.weight" in sd:
            sp = f"{lp}2.{s}."
            xn = layer_norm(x, sd[sp + "0.norm.weight"], sd[sp + "0.norm.bias"])
            x = lens_attention(sd, sp + "0.fn.", xn, xn, latent_heads) + x
            x = lens_ff(sd, sp + "1.fn.", layer_norm(x, sd[sp + "1.norm.weight"], sd[sp + "1.norm.bias"])) + x
            s += 1
        d += 1
    return x


# --------------------------------------------------------------------------- towers
def vit_trunk(sd: SD, pre: str, tokens, heads: int, act=gelu, taps: Optional[dict] = None):
    """The tail of VisionTransformer.forward transformer.py:755-787: prepend cls, add positional
    embedding, ln_pre, N blocks, ln_post on the cls row, @ proj."""
    B = tokens.shape[0]
    cls = sd[pre + "class_embedding"].view(1, 1, -1).expand(B, -1, -1)
    x = torch.cat([cls, tokens], dim=1) + sd[pre + "positional_embedding"]
    x = layer_norm(x, sd[pre + "ln_pre.weight"], sd[pre + "ln_pre.bias"])
    if taps is not None:
        taps[pre + "ln_pre"] = x
    x = transformer(sd, pre + "transformer.", x, heads, act, None, taps)
    pooled = layer_norm(x[:, 0], sd[pre + "ln_post.weight"], sd[pre + "ln_post.bias"])
    return pooled @ sd[pre + "proj"]


def image_tower(sd: SD, pre: str, image, heads: int, act=gelu, taps: Optional[dict] = None):
    """VisionTransformer.forward for visual_modality_type == 'image' (transformer.py:714-718,
    img_adapter_forawrd :659-677)."""
    w = sd[pre + "conv1.weight"]
    tok = conv_patch_embed(image, w, (w.shape[2], w.shape[3]))
    if taps is not None:
        taps[pre + "conv1"] = tok
    return vit_trunk(sd, pre, tok, heads, act, taps)


def lens_tower(sd: SD, pre: str, x, modality: str, heads: int, act=gelu,
               perceiver_as_identity: bool = False, latent_heads: int = 16, cross_heads: int = 1,
               fps_start=None, taps: Optional[dict] = None, perceiver_as_transformer: bool = False, **adapter_kw):
    """VisionTransformer.forward for audio / depth / pc / eeg (transformer.py:724-753): adapter ->
    x + pos -> Lens (or Identity, or a plain Transformer) -> ViT trunk; tactile takes the image path (transformer.py:715-718)."""
    ap = pre + "visual_adapter."
    if modality == "tactile":
        return image_tower(sd, pre, x, heads, act, taps)
    if modality == "eeg":
        tok, pos = eeg_adapter(sd, ap, x, **adapter_kw)
    elif modality == "audio":
        tok, pos = audio_adapter(sd, ap, x, **adapter_kw)
    elif modality == "depth":
        tok, pos = depth_adapter(sd, ap, x)
    elif modality in ("pc", "3dpc"):
        tok, pos = point_adapter(sd, ap, x, fps_start, **adapter_kw)
    else:
        raise NotImplementedError(modality)
    t = tok + pos
    if taps is not None:
        taps[pre + "adapter"] = t
    if perceiver_as_transformer and not perceiver_as_identity:
        # perceiver.py:372-381 builds transformer.Transformer and transformer.py:751 calls it on the BATCH-FIRST tokens
        # [B, M, C]; nn.MultiheadAttention (batch_first=False) therefore attends over dim 0 -- across the samples of the batch,
        # separately for every token position.  Restated as is: swap the two leading dims around a batch-first transformer.
        t = transformer(sd, pre + "perceiver.", t.transpose(0, 1), heads, act).transpose(0, 1)
    elif not perceiver_as_identity:
        t = perceiver(sd, pre + "perceiver.", t, cross_heads, latent_heads)
        if taps is not None:
            taps[pre + "perceiver"] = t
    return vit_trunk(sd, pre, t, heads, act, taps)


def causal_mask(n: int):
    """TextTransformer.build_attention_mask transformer.py:870-876."""
    return torch.full((n, n), float("-inf")).triu_(1)


def text_tower(sd: SD, text, heads: int, act=gelu):
    """CLIP.encode_text model.py:297-307 / TriCLIP.encode_text model.py:528-540."""
    x = sd["token_embedding.weight"][text] + sd["positional_embedding"]
    x = transformer(sd, "transformer.", x, heads, act, causal_mask(text.shape[1]))
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)]
    return x @ sd["text_projection"]


# --------------------------------------------------------------------------- losses
def cross_entropy_arange(logits, offset: int = 0):
    """F.cross_entropy(logits, arange(n) + offset), mean reduction."""
    n = logits.shape[0]
    lse = torch.logsumexp(logits, dim=-1)
    tgt = logits[torch.arange(n), torch.arange(n) + offset]
    return (lse - tgt).mean()


def clip_loss(x, y, logit_scale):
    """ClipLoss.forward loss.py:372-385 at world_size == 1: logits = (s*x) @ y^T (precedence of
    `logit_scale * x @ y.T`), both directions, /2."""
    lx = (logit_scale * x) @ y.t()
    ly = (logit_scale * y) @ x.t()
    return (cross_entropy_arange(lx) + cross_entropy_arange(ly)) / 2


def tri_clip_loss(image, text, visual, logit_scale):
    """TriClipLoss.forward loss.py:140-165: pairs (image, visual) and (text, visual), 4 CE terms / 2."""
    return clip_loss(image, visual, logit_scale) + clip_loss(text, visual, logit_scale)


def sim_mask(all_x, sim_thres: float):
    """ClipLossSimMask.get_logits loss.py:535-542 / 566-573: pairs whose teacher features are at least `sim_thres` similar are
    masked out, the diagonal is always kept."""
    sim = all_x.detach() @ all_x.detach().t()
    return torch.logical_or(torch.logical_not(sim >= sim_thres), torch.eye(all_x.shape[0], dtype=torch.bool))


def label_mask(all_x_labels, all_y_labels):
    """ClipLossLabelMask.get_logits loss.py:667-679 / 706-717: same-class pairs are masked out, the diagonal is kept."""
    if all_x_labels.ndim == 1:
        all_x_labels, all_y_labels = all_x_labels.unsqueeze(0), all_y_labels.unsqueeze(0)
    return torch.logical_or(torch.logical_not(all_x_labels.t() == all_y_labels), torch.eye(all_x_labels.shape[1], dtype=torch.bool))


def clip_loss_sharded(xs: Sequence[torch.Tensor], ys: Sequence[torch.Tensor], logit_scale, rank: int,
                      local_loss: bool, gather_with_grad: bool, mask=None):
    """What rank `rank` computes in ClipLoss / gather_features (loss.py:20-78, 346-385) at
    world_size == len(xs); xs[r], ys[r] are rank r's local feature blocks.  Non-grad gather:
    remote blocks are detached; the local block keeps grad only when not local_loss
    (loss.py:63-76).  `mask` (bool [B_all, B_all]): the mask variants' `logits * mask` (loss.py:544-575, 681-722) -- the
    x-direction logits take mask rows, the y-direction logits mask.T rows."""
    W = len(xs)
    if gather_with_grad:
        ax = torch.cat(list(xs), 0)
        ay = torch.cat(list(ys), 0)
    else:
        keep = not local_loss
        ax = torch.cat([xs[r] if (r == rank and keep) else xs[r].detach() for r in range(W)], 0)
        ay = torch.cat([ys[r] if (r == rank and keep) else ys[r].detach() for r in range(W)], 0)
    if local_loss:
        lx = (logit_scale * xs[rank]) @ ay.t()
        ly = (logit_scale * ys[rank]) @ ax.t()
        bl = xs[rank].shape[0]
        off = bl * rank
        if mask is not None:
            lx = lx * mask[off:off + bl]
            ly = ly * mask.t()[off:off + bl]
        return (cross_entropy_arange(lx, off) + cross_entropy_arange(ly, off)) / 2
    lx = (logit_scale * ax) @ ay.t()
    if mask is not None:
        lx = lx * mask
    return (cross_entropy_arange(lx) + cross_entropy_arange(lx.t())) / 2


# --------------------------------------------------------------------------- whole models
def heads_of(width: int, head_width: int = 64) -> int:
    return width // head_width


def clip_forward(sd: SD, image, text, vision_heads: int, text_heads: int, act=gelu):
    """CLIP.forward model.py:309-326 -> (image_features, text_features, logit_scale.exp())."""
    fi = l2_normalize(image_tower(sd, "visual.", image, vision_heads, act))
    ft = l2_normalize(text_tower(sd, text, text_heads, act))
    return fi, ft, sd["logit_scale"].exp()


def triclip_forward(sd: SD, image, text, visual_x, modality: str, vision_heads: int, text_heads: int,
                    act=gelu, **lens_kw):
    """TriCLIP.forward model.py:542-621 ('original impl' branch)."""
    if image.ndim == 5:  # model.py:591-604: normalise every frame's features, average over frames, normalise again
        b, t = image.shape[:2]
        fi = l2_normalize(image_tower(sd, "image.", image.reshape((b * t,) + tuple(image.shape[2:])), vision_heads, act))
        fi = l2_normalize(fi.reshape(b, t, -1).mean(1))
    else:
        fi = l2_normalize(image_tower(sd, "image.", image, vision_heads, act))
    ft = l2_normalize(text_tower(sd, text, text_heads, act))
    fv = l2_normalize(lens_tower(sd, "visual.", visual_x, modality, vision_heads, act, **lens_kw))
    return fi, ft, fv, sd["logit_scale"].exp()


# --------------------------------------------------------------------------- input pipelines
def kaldi_fbank(waveform, sample_rate: float = 16000.0, n_mel: int = 128, frame_length_ms: float = 25.0, frame_shift_ms: float = 10.0,
                preemphasis: float = 0.97, low_freq: float = 20.0):
    """The arithmetic lives in a THIRD-PARTY dependency: torchaudio.compliance.kaldi.fbank (reference requirements: torchaudio,
    unpinned), called at modal_audio/processors/at_processor.py:854-863 with htk_compat=True, use_energy=False,
    window_type="hanning", dither=0.0, frame_shift=10.  Restated from the published Kaldi algorithm (feature-window.cc,
    mel-computations.cc): snip_edges framing, per-frame DC removal, pre-emphasis with the first sample against itself, Hann window
    (symmetric), zero-pad to 512, power spectrum, triangular filters on the HTK mel scale over [20 Hz, Nyquist], log(max(., eps)).
    waveform [samples] -> [frames, n_mel].  tests/ pin this against torchaudio itself wherever torchaudio is importable."""
    x = waveform.double()
    flen, shift = int(sample_rate * frame_length_ms * 0.001), int(sample_rate * frame_shift_ms * 0.001)
    n_frames = 1 + (x.numel() - flen) // shift
    fr = x.unfold(0, flen, shift)[:n_frames]
    fr = fr - fr.mean(dim=1, keepdim=True)
    prev = torch.cat([fr[:, :1], fr[:, :-1]], dim=1)
    fr = fr - preemphasis * prev
    n = torch.arange(flen, dtype=torch.float64)
    fr = fr * (0.5 - 0.5 * torch.cos(2 * math.pi * n / (flen - 1)))
    nfft = 512
    fr = torch.nn.functional.pad(fr, (0, nfft - flen))
    k = torch.arange(nfft // 2 + 1, dtype=torch.float64)
    ang = -2 * math.pi * k[:, None] * torch.arange(nfft, dtype=torch.float64)[None, :] / nfft
    power = (fr @ torch.cos(ang).t()) ** 2 + (fr @ torch.sin(ang).t()) ** 2  # explicit DFT
    mel = lambda f: 1127.0 * torch.log(1.0 + f / 700.0)  # noqa: E731
    lo, hi = mel(torch.tensor(low_freq, dtype=torch.float64)), mel(torch.tensor(0.5 * sample_rate, dtype=torch.float64))
    delta = (hi - lo) / (n_mel + 1)
    b = torch.arange(n_mel, dtype=torch.float64)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
    m = mel(sample_rate / nfft * torch.arange(nfft // 2, dtype=torch.float64))[None, :]
    bins = torch.clamp(torch.minimum((m - left) / (center - left), (right - m) / (right - center)), min=0)
    bins = torch.nn.functional.pad(bins, (0, 1))
    return torch.log(torch.clamp(power @ bins.t(), min=1.1920929e-07)).float()


def ast_clip(waveform, target_length: int = 512, mean: float = -4.2677393, std: float = 4.5689974, **kw):
    """convert2fbank + transform of AudioASTProcessorEval (at_processor.py:845-872): fbank, zero-pad / crop to target_length
    frames, (x - mean) / std."""
    fb = kaldi_fbank(waveform, **kw)
    p = target_length - fb.shape[0]
    fb = torch.nn.functional.pad(fb, (0, 0, 0, p)) if p > 0 else fb[:target_length]
    return (fb - mean) / std


def pc_norm(pc):
    """modal_3d/processors/pc_processor.py:32-38 (xyz in the first three channels)."""
    xyz = pc[:, :3] - pc[:, :3].mean(dim=0)
    xyz = xyz / xyz.pow(2).sum(dim=1).sqrt().max()
    return torch.cat([xyz, pc[:, 3:]], dim=1)


def depth_norm(depth, max_depth: float = 75.0, min_depth: float = 0.01, clamp_max_before_scale: bool = True, mean: float = 0.0418, std: float = 0.0295):
    """modal_depth/processors/transforms_rgbd.py:393-413 (DepthNorm) followed by transforms.Normalize on the depth channel
    (vt_processor.py:311-322,170-171)."""
    d = depth.clamp(min=min_depth)
    if clamp_max_before_scale:
        d = d.clamp(max=max_depth)
    return (d / max_depth - mean) / std
