"""The C-ABI shared library loads without a GPU / driver and exports every symbol include/vitlens_b200.h declares."""
import ctypes
import os
import re

from tests.common import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "vitlens_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vl_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from vitlens_b200 import lib

    h = lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/vitlens_b200.h but not exported"
    assert h.vl_abi_version() == 2


def test_argument_errors_are_reported_without_a_gpu():
    from vitlens_b200 import lib

    h = lib.load()
    h.vl_last_error.restype = ctypes.c_char_p
    rc = h.vl_colsum_bf16(None, ctypes.c_int64(8), None, 4, 8, None)
    assert rc == -1 and b"vl_colsum_bf16" in h.vl_last_error()


def test_product_has_no_cpu_path():
    import pytest
    import torch

    from vitlens_b200 import ops

    with pytest.raises(AssertionError, match="CUDA"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
