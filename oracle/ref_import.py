"""Import the *real* reference (read-only at /root/reference) on CPU  --  TEST INFRASTRUCTURE.

Only usable in the build container (``/root/reference`` does not travel to the GPU box).
Installs in-memory stubs for the pure-Python wheels the reference imports but that are
not installed here: easydict, timm.models.{hub,layers},
ftfy, matplotlib / mpl_toolkits, termcolor.  None of them does hot-path arithmetic.
Must run in a process that has NOT imported this repo's own ``open_clip`` package.
"""
from __future__ import annotations

import os
import sys
import types

REF_SRC = "/root/reference/vitlens/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "open_clip"))


def _install_stubs():
    import torch
    import transformers  # noqa: F401  (must precede the timm stub: its lazy loader probes timm.__spec__)

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            d = dict(d or {})
            d.update(kw)
            for k, v in d.items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setattr__(k, v)
            super().__setitem__(k, v)

        __setitem__ = __setattr__

        def update(self, e=None, **f):
            d = dict(e or {})
            d.update(f)
            for k in d:
                setattr(self, k, d[k])

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("easydict", EasyDict=EasyDict)
    mod("timm")
    mod("timm.models")
    mod("timm.models.hub", get_cache_dir=lambda *a, **k: "/tmp", download_cached_file=lambda *a, **k: None)
    drop_path = type("DropPath", (torch.nn.Identity,), {"__init__": lambda s, p=0.0: torch.nn.Identity.__init__(s)})
    mod("timm.models.layers", DropPath=drop_path)
    mod("ftfy", fix_text=lambda s: s)
    mod("termcolor", colored=lambda s, *a, **k: s)
    mod("matplotlib")
    mod("matplotlib.pyplot")
    mod("mpl_toolkits")
    mod("mpl_toolkits.mplot3d", Axes3D=object)
    return EasyDict


def import_reference():
    """Returns (open_clip module of the reference, fetch_model_cfg, EasyDict)."""
    assert available(), "reference tree not present"
    assert "open_clip" not in sys.modules, "another open_clip is already imported in this process"
    EasyDict = _install_stubs()
    sys.path.insert(0, REF_SRC)
    import open_clip  # noqa: E402
    from mm_vit_lens.model_cfg import fetch_model_cfg  # noqa: E402

    return open_clip, fetch_model_cfg, EasyDict


def modality_args(fetch_model_cfg, modality: str, **overrides):
    """The ``args`` object the reference's tri_create_model needs (factory.py:246-258,348;
    transformer.py:593,604-606; model.py:549-552)."""
    cfg = fetch_model_cfg(modality=modality)
    cfg.pretrained = None
    for k in ("unlock_from_head", "vid_use_fpos", "vid_use_ltpos", "vid_distill_tokens"):
        setattr(cfg, k, False)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg
