"""ctypes binding of libvitlens_b200.so (include/vitlens_b200.h).

torch is used only for device memory (``tensor.data_ptr()``) and the current CUDA stream.
There is NO fallback: if the shared library is missing or a call fails this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import build as _build

_lock = threading.Lock()
_lib = None
launch_count = 0  # number of vl_* kernel-launching calls issued (bench.py reports it)
# bench.py roofline instrumentation: when a list, every GEMM / attention launch is bracketed by CUDA events on
# the launching stream and (start, end, algorithmic flops | kind, bytes) is appended.
GEMM_TIMING = None
ATTN_TIMING = None
# bench.py's per-kernel breakdown pass: list of (entry point, start event, end event, bound, algorithmic amount) with bound
# "tensor" (amount = FLOPs) / "hbm" (amount = bytes) / None
CALL_TIMING = None


def _timed(store, payload, fn):
    if store is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    store.append((e0, e1) + payload)


class VlError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
        ("a_mn", C.c_int32), ("b_mn", C.c_int32),
        ("d_f32", C.c_int32), ("accumulate", C.c_int32), ("split_k", C.c_int32),
        ("epilogue", C.c_int32), ("act_quick", C.c_int32), ("alpha", C.c_float),
        ("bias", C.c_void_p), ("aux_in", C.c_void_p), ("aux_out", C.c_void_p), ("ldaux", C.c_int64),
        ("row_vec", C.c_void_p), ("col_vec", C.c_void_p), ("out_vec0", C.c_void_p), ("out_vec1", C.c_void_p),
        ("out_vec2", C.c_void_p), ("scalar_out", C.c_void_p), ("iparam", C.c_int32), ("fparam", C.c_float),
        ("alpha_dev", C.c_void_p), ("fparam_dev", C.c_void_p), ("aux_row_div", C.c_int32), ("relu", C.c_int32), ("rowsum_out", C.c_void_p),
        ("loss_flags", C.c_int32),
        ("b_peers", C.c_void_p), ("b_npeers", C.c_int32), ("b_peer_rows", C.c_int32), ("peer_flags", C.c_void_p), ("peer_flag_value", C.c_int32),
        ("mask", C.c_void_p), ("ldmask", C.c_int64),
    ]


EPI_LINEAR, EPI_GELU, EPI_RESIDUAL, EPI_GELU_BWD, EPI_GEGLU, EPI_ROWLSE, EPI_CLIPGRAD = 0, 1, 2, 3, 4, 5, 6
ABI_VERSION = 2  # include/vitlens_b200.h: VL_ABI_VERSION


def lib_path() -> str:
    # VL_LIB_PATH: load another build of the same ABI (A/B timing of two kernel versions on one box)
    return os.environ.get("VL_LIB_PATH") or _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the in-tree .so is absent/stale and nvcc exists)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if build_if_missing and not os.environ.get("VL_LIB_PATH") and _build.needs_build():
            try:
                _build.build()
            except Exception as e:
                # A box without nvcc (the GPU box ships the prebuilt .so whose mtime the snapshot may have reordered) can still
                # run an existing library -- the ABI check below rejects one that does not match this binding.  A compile
                # ERROR with nvcc present is never swallowed: the sources and the library would disagree.
                if not os.path.exists(path) or "nvcc not found" not in str(e):
                    raise VlError(f"libvitlens_b200.so is stale or missing and could not be built: {e}") from e
        if not os.path.exists(path):
            raise VlError(f"{path} not found: run `python __graft_entry__.py` (build()) first; there is no fallback path")
        lib = C.CDLL(path)
        lib.vl_last_error.restype = C.c_char_p
        lib.vl_abi_version.restype = C.c_int
        if lib.vl_abi_version() != ABI_VERSION:
            raise VlError(f"{path} implements ABI {lib.vl_abi_version()}, this binding needs {ABI_VERSION}: rebuild (python __graft_entry__.py)")
        _lib = lib
        # bring-up knobs for A/B runs (kernel variants; see vl_debug_set call sites in csrc/): VL_DEBUG="13=1,12=1"
        for kv in os.environ.get("VL_DEBUG", "").split(","):
            if "=" in kv:
                k, v = kv.split("=")
                lib.vl_debug_set(int(k), int(v))
        return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().vl_last_error().decode(errors="replace")
        raise VlError(f"{what} failed (rc={rc}): {msg}")


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def debug_set(key: int, value: int):
    _check(load().vl_debug_set(int(key), int(value)), "vl_debug_set")


def _count():
    global launch_count
    launch_count += 1


def gemm(a, b, d, *, M, N, K, lda, ldb, ldd, a_mn=False, b_mn=False, epilogue=EPI_LINEAR, bias=None,
         aux_in=None, aux_out=None, ldaux=0, alpha=1.0, accumulate=False, split_k=1, act_quick=False,
         row_vec=None, col_vec=None, out_vec0=None, out_vec1=None, out_vec2=None, scalar_out=None, iparam=0, fparam=0.0,
         alpha_dev=None, fparam_dev=None, aux_row_div=0, relu=False, rowsum_out=None, loss_flags=0,
         b_peers=None, b_peer_rows=0, peer_flags=0, peer_flag_value=0, mask=None):
    """Raw GEMM call; see include/vitlens_b200.h.  a, b bf16; d bf16 or fp32; bias fp32."""
    assert a.dtype == torch.bfloat16 and (b is None or b.dtype == torch.bfloat16)
    assert d is None or d.dtype in (torch.bfloat16, torch.float32)
    peer_arr = None
    if b_peers is not None:  # raw device addresses of B's row blocks in the peers' arenas
        peer_arr = (C.c_void_p * len(b_peers))(*[int(q) for q in b_peers])
    assert bias is None or bias.dtype == torch.float32
    args = GemmArgs(
        _ptr(a), _ptr(b), _ptr(d), M, N, K, lda, ldb, ldd, int(a_mn), int(b_mn),
        int(d is not None and d.dtype == torch.float32), int(accumulate), int(split_k), int(epilogue), int(act_quick), float(alpha),
        _ptr(bias), _ptr(aux_in), _ptr(aux_out), ldaux,
        _ptr(row_vec), _ptr(col_vec), _ptr(out_vec0), _ptr(out_vec1), _ptr(out_vec2), _ptr(scalar_out), int(iparam), float(fparam),
        _ptr(alpha_dev), _ptr(fparam_dev), int(aux_row_div), int(relu), _ptr(rowsum_out), int(loss_flags),
        C.cast(peer_arr, C.c_void_p) if peer_arr is not None else C.c_void_p(0), len(b_peers) if b_peers is not None else 0, int(b_peer_rows),
        C.c_void_p(int(peer_flags)), int(peer_flag_value), _ptr(mask), 0 if mask is None else mask.stride(0))
    _count()
    if CALL_TIMING is not None:
        kind = "wgrad" if (a_mn and b_mn) else ("dgrad" if b_mn else "fwd")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _check(_fn("vl_gemm_bf16")(C.cast(C.pointer(args), C.c_void_p), _stream()), "vl_gemm_bf16")
        e1.record()
        CALL_TIMING.append((f"vl_gemm_bf16/{kind}/epi{int(epilogue)}", e0, e1, "tensor", 2.0 * M * N * K))
        return
    _timed(GEMM_TIMING, (2.0 * M * N * K, (M, N, K, int(epilogue), int(a_mn), int(b_mn))),
           lambda: _check(_fn("vl_gemm_bf16")(C.cast(C.pointer(args), C.c_void_p), _stream()), "vl_gemm_bf16"))


# ----------------------------------------------------------------------------- prototypes
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_PROTOS = {
    "vl_gemm_bf16": [_P, _P],
    "vl_clip_backward": [_P, _P],
    "vl_attention_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _L, _F, _I, _P],
    "vl_attention_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _L, _L, _L, _L, _L, _F, _I, _P],
    "vl_layernorm_fwd": [_P, _L, _P, _P, _P, _P, _L, _P, _P, _I, _I, _F, _P],
    "vl_layernorm_bwd": [_P, _L, _P, _L, _P, _P, _P, _P, _P, _L, _P, _L, _P, _P, _P, _I, _I, _P],
    "vl_colsum_bf16": [_P, _L, _P, _I, _I, _P],
    "vl_patchify": [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _L, _L, _I, _P],
    "vl_assemble_tokens": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vl_assemble_tokens_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vl_embed_tokens": [_P, _P, _P, _P, _L, _I, _I, _P],
    "vl_embed_tokens_bwd": [_P, _P, _P, _P, _L, _I, _I, _P],
    "vl_l2norm_fwd": [_P, _P, _P, _I, _I, _F, _P],
    "vl_l2norm_bwd": [_P, _P, _P, _P, _I, _I, _P],
    "vl_geglu_fwd": [_P, _P, _L, _I, _P],
    "vl_geglu_bwd": [_P, _P, _P, _L, _I, _P],
    "vl_cast_f32_bf16": [_P, _P, _L, _P],
    "vl_add_bf16": [_P, _P, _P, _L, _P],
    "vl_adamw_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P],
    "vl_adamw_multi": [_P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _I, _F, _I, _L, _P],
    "vl_multi_sqnorm": [_P, _P, _P, _I, _P, _I, _L, _P],
    "vl_adamw_multi_clip": [_P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _I, _F, _P, _F, _I, _L, _P],
    "vl_comm_signal": [_I, _I, _P],
    "vl_comm_wait": [_I, _I, _P],
    "vl_comm_peer_reduce_f32": [_L, _L, _P, _P],
    "vl_comm_peer_gather_f32": [_L, _L, _P, _P],
    "vl_allgather_features": [_L, _L, _I, _I, _P],
    "vl_allreduce_grads": [_L, _L, _I, _I, _P],
    "vl_lse_combine": [_P, _P, _P, _I, _I, _P, _P, _P],
    "vl_fps": [_P, _P, _I, _I, _I, _P, _P, _P],
    "vl_knn_group": [_P, _P, _I, _I, _I, _I, _P, _P, _P],
    "vl_linear3": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "vl_group_max_bwd": [_P, _P, _P, _L, _I, _I, _P],
    "vl_group_sum": [_P, _P, _L, _I, _I, _P],
    "vl_colsum2_bf16": [_P, _P, _P, _P, _L, _I, _P],
    "vl_wgrad3": [_P, _P, _P, _L, _I, _P],
    "vl_moments3": [_P, _P, _L, _P],
    "vl_col_affine_bf16": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "vl_group_max": [_P, _P, _P, _L, _I, _I, _P],
    "vl_fbank": [_P, _L, _I, _I, _I, _I, _P, _P, _I, _F, _I, _F, _F, _P, _P],
    "vl_pc_norm": [_P, _P, _I, _I, _I, _P],
    "vl_depth_norm": [_P, _P, _L, _F, _F, _I, _F, _F, _P],
    "vl_template_mean": [_P, _P, _I, _I, _I, _L, _I, _P],
    "vl_topk_rows": [_P, _L, _I, _I, _I, _P, _P, _P],
    "vl_average_precision": [_P, _L, _P, _L, _I, _I, _I, _P, _P, _P],
}


def _fn(name):
    lib = load()
    f = getattr(lib, name)
    if not getattr(f, "_vl_ready", False):
        f.argtypes = _PROTOS[name]
        f.restype = C.c_int
        f._vl_ready = True
    return f


def _call(name, *args, work=None):
    _count()
    if CALL_TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _fn(name)(*args, _stream())
        e1.record()
        CALL_TIMING.append((name, e0, e1) + (work if work is not None else (None, 0.0)))
    else:
        rc = _fn(name)(*args, _stream())
    if rc != 0:
        _check(rc, name)


def _p(t):
    return None if t is None else t.data_ptr()


BF16, F32 = torch.bfloat16, torch.float32


def attention_fwd(q, k, v, o, lse, *, B, H, nq, nk, ldq, ldk, ldv, ldo, scale, causal=False):
    byts = 2.0 * B * H * 64 * (2 * nq + 2 * nk)  # read Q,K,V + write O (bf16)
    _timed(ATTN_TIMING, ("fwd", byts),
           lambda: _call("vl_attention_fwd", _p(q), _p(k), _p(v), _p(o), _p(lse), B, H, nq, nk, ldq, ldk, ldv, ldo, float(scale), int(causal),
                         work=("hbm", byts)))


def attention_bwd(q, k, v, o, dout, lse, dq, dk, dv, *, B, H, nq, nk, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, scale, causal=False):
    byts = 2.0 * B * H * 64 * (4 * nq + 4 * nk)  # read Q,K,V,O,dO + write dQ,dK,dV (bf16)
    _timed(ATTN_TIMING, ("bwd", byts),
           lambda: _call("vl_attention_bwd", _p(q), _p(k), _p(v), _p(o), _p(dout), _p(lse), _p(dq), _p(dk), _p(dv), B, H, nq, nk,
                         ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, float(scale), int(causal), work=("hbm", byts)))


def layernorm_fwd(x, w, b, y, mean, rstd, *, T, D, ldx, ldy, row_index=None, eps=1e-5):
    _call("vl_layernorm_fwd", _p(x), ldx, _p(row_index), _p(w), _p(b), _p(y), ldy, _p(mean), _p(rstd), T, D, float(eps),
          work=("hbm", 2.0 * T * D * 2))  # read x, write y (bf16)


def layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db, *, T, D, lddy, ldx, lddx, dres=None, lddres=0, row_index=None, dres_sum=None):
    _call("vl_layernorm_bwd", _p(dy), lddy, _p(x), ldx, _p(row_index), _p(w), _p(mean), _p(rstd), _p(dres), lddres,
          _p(dx), lddx, _p(dw), _p(db), _p(dres_sum), T, D,
          work=("hbm", (3.0 + (dres is not None)) * T * D * 2))  # read dy, x (+ dres), write dx (bf16)


def colsum(dy, db, *, T, N, ld):
    _call("vl_colsum_bf16", _p(dy), ld, _p(db), T, N)


def patchify(inp, out, *, B, C, OH, OW, kh, kw, stride_h, stride_w, sb, sc, sh, sw, Kpad):
    assert inp.dtype in (F32, BF16)
    _call("vl_patchify", _p(inp), int(inp.dtype == BF16), _p(out), B, C, OH, OW, kh, kw, stride_h, stride_w, sb, sc, sh, sw, Kpad)


def assemble_tokens(tok, cls, pos, out, *, B, L, D, has_cls):
    _call("vl_assemble_tokens", _p(tok), _p(cls), _p(pos), _p(out), B, L, D, int(has_cls))


def assemble_tokens_bwd(dx, dtok, dpos, dcls, *, B, L, D, has_cls):
    _call("vl_assemble_tokens_bwd", _p(dx), _p(dtok), _p(dpos), _p(dcls), B, L, D, int(has_cls))


def embed_tokens(ids, table, pos, out, *, rows, ctx, D):
    _call("vl_embed_tokens", _p(ids), _p(table), _p(pos), _p(out), rows, ctx, D)


def embed_tokens_bwd(ids, dx, dtable, dpos, *, rows, ctx, D):
    _call("vl_embed_tokens_bwd", _p(ids), _p(dx), _p(dtable), _p(dpos), rows, ctx, D)


def l2norm_fwd(x, y, inv_norm, *, B, E, eps=1e-12):
    _call("vl_l2norm_fwd", _p(x), _p(y), _p(inv_norm), B, E, float(eps))


def l2norm_bwd(dy, y, inv_norm, dx, *, B, E):
    _call("vl_l2norm_bwd", _p(dy), _p(y), _p(inv_norm), _p(dx), B, E)


def geglu_fwd(h, out, *, M, F):
    _call("vl_geglu_fwd", _p(h), _p(out), M, F, work=("hbm", 3.0 * M * F * 2))


def geglu_bwd(h, dout, dh, *, M, F):
    _call("vl_geglu_bwd", _p(h), _p(dout), _p(dh), M, F, work=("hbm", 5.0 * M * F * 2))


def cast_f32_bf16(inp, out):
    _call("vl_cast_f32_bf16", _p(inp), _p(out), inp.numel())


def add_bf16(a, b, out):
    _call("vl_add_bf16", _p(a), _p(b), _p(out), a.numel())


def adamw_step(p, g, m, v, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    _call("vl_adamw_step", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step, grad_scale)


def rowlse_parts(N: int) -> int:
    f = load().vl_gemm_rowlse_parts
    f.argtypes, f.restype = [C.c_int32], C.c_int
    return f(N)


class ClipBwdArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("y_peers", C.c_void_p), ("y_npeers", C.c_int32), ("y_peer_rows", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("E", C.c_int32), ("ldx", C.c_int64), ("ldy", C.c_int64),
        ("dx", C.c_void_p), ("lddx", C.c_int64), ("row_lse", C.c_void_p), ("col_lse", C.c_void_p), ("label_off", C.c_int32),
        ("alpha_dev", C.c_void_p), ("gscale", C.c_float), ("gscale_dev", C.c_void_p), ("ds_out", C.c_void_p), ("ds_row_only", C.c_int32),
        ("mask", C.c_void_p), ("ldmask", C.c_int64),
    ]


def clip_backward(x, y, dx, *, M, N, E, ldx, ldy, row_lse, col_lse, label_off, alpha_dev, gscale, gscale_dev=None, ds_out=None,
                  ds_row_only=False, mask=None, y_peers=None, y_peer_rows=0):
    """vl_clip_backward; see include/vitlens_b200.h.  y: bf16 [N, E] or None with y_peers (device addresses of the ranks' row blocks)."""
    peer_arr = None
    if y_peers is not None:
        peer_arr = (C.c_void_p * len(y_peers))(*[int(a) for a in y_peers])
    args = ClipBwdArgs(
        _ptr(x), _ptr(y), C.cast(peer_arr, C.c_void_p) if peer_arr is not None else C.c_void_p(0), len(y_peers) if y_peers is not None else 0,
        int(y_peer_rows), int(M), int(N), int(E), int(ldx), int(ldy), _ptr(dx), int(dx.stride(0)), _ptr(row_lse), _ptr(col_lse), int(label_off),
        _ptr(alpha_dev), float(gscale), _ptr(gscale_dev), _ptr(ds_out), int(ds_row_only), _ptr(mask), 0 if mask is None else mask.stride(0))
    _call("vl_clip_backward", C.cast(C.pointer(args), C.c_void_p), work=("tensor", 2.0 * M * N * E * 2))


def lse_combine(part_max, part_sum, diag, lse, loss_sum, *, M, nparts):
    _call("vl_lse_combine", _p(part_max), _p(part_sum), _p(diag), M, nparts, _p(lse), _p(loss_sum))


def fps(xyz, start, idx_out, centers, *, B, N, npoint):
    _call("vl_fps", _p(xyz), _p(start), B, N, npoint, _p(idx_out), _p(centers))


def knn_group(xyz, centers, nb_out, idx_out, *, B, N, G, k):
    _call("vl_knn_group", _p(xyz), _p(centers), B, N, G, k, _p(nb_out), _p(idx_out))


def linear3(x, w, scale, shift, out, *, R, C, act, pre_out=None):
    _call("vl_linear3", _p(x), _p(w), _p(scale), _p(shift), _p(out), _p(pre_out), R, C, act)


def group_max_bwd(dout, arg, dx, *, groups, G, C):
    _call("vl_group_max_bwd", _p(dout), _p(arg), _p(dx), groups, G, C)


def group_sum(x, out, *, groups, G, C):
    _call("vl_group_sum", _p(x), _p(out), groups, G, C)


def colsum2(a, b, s1, s2, *, T, N):
    _call("vl_colsum2_bf16", _p(a), _p(b), _p(s1), _p(s2), T, N)


def wgrad3(dy, x, dw, *, R, C):
    _call("vl_wgrad3", _p(dy), _p(x), _p(dw), R, C)


def moments3(x, out12, *, R):
    _call("vl_moments3", _p(x), _p(out12), R)


def col_affine(a, b, p0, p1, p2, out, *, R, C, act):
    _call("vl_col_affine_bf16", _p(a), _p(b), _p(p0), _p(p1), _p(p2), _p(out), R, C, act)


def group_max(x, out, arg, *, groups, G, C):
    _call("vl_group_max", _p(x), _p(out), _p(arg), groups, G, C)


def fbank(wav, window, mel, out, *, clip_stride, n_clips, n_samples, frame_len, frame_shift, n_mel, preemph, target_len, mean, std):
    _call("vl_fbank", _p(wav), clip_stride, n_clips, n_samples, frame_len, frame_shift, _p(window), _p(mel), n_mel, float(preemph), target_len,
          float(mean), float(std), _p(out))


def pc_norm(inp, out, *, B, N, C):
    _call("vl_pc_norm", _p(inp), _p(out), B, N, C)


def depth_norm(inp, out, *, n, min_depth, max_depth, clamp_max, mean, std):
    _call("vl_depth_norm", _p(inp), _p(out), n, float(min_depth), float(max_depth), int(clamp_max), float(mean), float(std))


def template_mean(x, out, *, G, T, E, ldo, transpose_out):
    _call("vl_template_mean", _p(x), _p(out), G, T, E, ldo, int(transpose_out))


def topk_rows(scores, idx_out, val_out, *, ld, rows, cols, k):
    _call("vl_topk_rows", _p(scores), ld, rows, cols, k, _p(idx_out), _p(val_out))


def average_precision(scores, targets, ap_out, npos_out, *, lds, ldt, N, C, apply_sigmoid):
    _call("vl_average_precision", _p(scores), lds, _p(targets), ldt, N, C, int(apply_sigmoid), _p(ap_out), _p(npos_out))


ADAM_CHUNK = 16384


def adamw_multi(ptrs, sizes, wds, chunk_tab, *, n_chunks, lr, beta1, beta2, eps, step, grad_scale=1.0, sumsq=None, max_norm=None, lrs=None,
                n_src=1, src_stride=0):
    if sumsq is None:
        _call("vl_adamw_multi", _p(ptrs), _p(sizes), _p(wds), _p(lrs), _p(chunk_tab), n_chunks, lr, beta1, beta2, eps, step, grad_scale,
              int(n_src), int(src_stride))
    else:
        _call("vl_adamw_multi_clip", _p(ptrs), _p(sizes), _p(wds), _p(lrs), _p(chunk_tab), n_chunks, lr, beta1, beta2, eps, step, grad_scale,
              _p(sumsq), float(max_norm), int(n_src), int(src_stride))


def multi_sqnorm(ptrs, sizes, chunk_tab, sumsq, *, n_chunks, n_src=1, src_stride=0):
    _call("vl_multi_sqnorm", _p(ptrs), _p(sizes), _p(chunk_tab), n_chunks, _p(sumsq), int(n_src), int(src_stride))


# ----------------------------------------------------------------------------- peer memory (csrc/comm.cu)
def comm_init(rank: int, world: int, arena_bytes: int):
    """-> (64-byte IPC handle, address of this rank's arena)"""
    lib = load()
    handle = C.create_string_buffer(64)
    arena = C.c_void_p()
    lib.vl_comm_init.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.vl_comm_init.restype = C.c_int
    _check(lib.vl_comm_init(rank, world, arena_bytes, handle, C.byref(arena)), "vl_comm_init")
    return handle.raw, arena.value


def comm_connect(all_handles: bytes):
    lib = load()
    lib.vl_comm_connect.argtypes = [C.c_char_p]
    lib.vl_comm_connect.restype = C.c_int
    _check(lib.vl_comm_connect(all_handles), "vl_comm_connect")


def comm_peer_ptr(peer: int) -> int:
    lib = load()
    out = C.c_void_p()
    lib.vl_comm_peer_ptr.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    lib.vl_comm_peer_ptr.restype = C.c_int
    _check(lib.vl_comm_peer_ptr(peer, C.byref(out)), "vl_comm_peer_ptr")
    return out.value


def comm_destroy():
    lib = load()
    lib.vl_comm_destroy.restype = C.c_int
    _check(lib.vl_comm_destroy(), "vl_comm_destroy")


def comm_signal(flag_idx, value):
    _call("vl_comm_signal", int(flag_idx), int(value))


def comm_wait(flag_idx, value):
    _call("vl_comm_wait", int(flag_idx), int(value))


def comm_peer_reduce(offset, n, out):
    _call("vl_comm_peer_reduce_f32", int(offset), int(n), _p(out))


def comm_peer_gather(offset, n, out):
    _call("vl_comm_peer_gather_f32", int(offset), int(n), _p(out))


def allgather_features(offset, nbytes, flag_idx, ticket):
    _call("vl_allgather_features", int(offset), int(nbytes), int(flag_idx), int(ticket))


def allreduce_grads(offset, nbytes, flag_idx, ticket):
    _call("vl_allreduce_grads", int(offset), int(nbytes), int(flag_idx), int(ticket))
