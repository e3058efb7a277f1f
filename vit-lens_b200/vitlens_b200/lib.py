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
    ]


EPI_LINEAR, EPI_GELU, EPI_RESIDUAL, EPI_GELU_BWD, EPI_GEGLU = 0, 1, 2, 3, 4


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the in-tree .so is absent/stale and nvcc exists)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if build_if_missing and _build.needs_build():
            try:
                _build.build()
            except Exception as e:  # stale .so is still better than nothing only if it exists
                if not os.path.exists(path):
                    raise VlError(f"libvitlens_b200.so is missing and could not be built: {e}") from e
        if not os.path.exists(path):
            raise VlError(f"{path} not found: run `python __graft_entry__.py` (build()) first; there is no fallback path")
        lib = C.CDLL(path)
        lib.vl_last_error.restype = C.c_char_p
        lib.vl_abi_version.restype = C.c_int
        _lib = lib
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
         aux_in=None, aux_out=None, ldaux=0, alpha=1.0, accumulate=False, split_k=1, act_quick=False):
    """Raw GEMM call; see include/vitlens_b200.h.  a, b bf16; d bf16 or fp32; bias fp32."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert d.dtype in (torch.bfloat16, torch.float32)
    assert bias is None or bias.dtype == torch.float32
    args = GemmArgs(
        _ptr(a), _ptr(b), _ptr(d), M, N, K, lda, ldb, ldd, int(a_mn), int(b_mn),
        int(d.dtype == torch.float32), int(accumulate), int(split_k), int(epilogue), int(act_quick), float(alpha),
        _ptr(bias), _ptr(aux_in), _ptr(aux_out), ldaux)
    _count()
    _check(load().vl_gemm_bf16(C.byref(args), _stream()), "vl_gemm_bf16")
