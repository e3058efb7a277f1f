"""Pin mm_vit_lens.ViTLens.encode against the real reference  --  TEST INFRASTRUCTURE (build container only).

    python oracle/make_golden_api.py

Runs the REFERENCE's own ViTLens class (mm_vit_lens/vitlens.py:21-189, unmodified): its `_init_modality_module` builds the
vitlensL towers with its own factory, its `encode` does the audio clip mean (vitlens.py:175-183), the text closure and the
final normalisation.  Two things it cannot do offline are supplied by the harness: `fetch_model_cfg` is wrapped so that
`pretrained` is None (the tag it ships needs a download), and the file -> tensor processors (PIL / torchaudio / numpy loaders,
mm_vit_lens/data_processors.py, which need omegaconf and media files) are replaced by pass-through callables that accept the
tensors those processors would return.  Weights: synth_state_dict over the ViTLens-level state_dict (keys `vitlens.<modality>.*`,
the release checkpoint's wire format, vitlens.py:153-159), so tests/ load the very same dict into this repo's ViTLens with
strict=True.  Output: tests/golden/vitlens_encode.pt (features per modality, normalised and raw).
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cases as C  # noqa: E402
from oracle import ref_import  # noqa: E402

MODALITIES = ["image", "text", "audio", "depth"]
OUT = os.path.join(C.GOLDEN_DIR, "vitlens_encode.pt")


def api_inputs():
    """Tensors in the layout the reference's processors return: image [B,3,224,224], text ids [B,77], audio [B, S clips, T, F]
    (vitlens.py:175-178), depth [B,1,224,224]."""
    s = C._synth()
    return {
        "image": s.synth_normal("api_image", (2, 3, 224, 224), seed=7),
        "text": s.synth_text(2, 77, 49408, seed=7),
        "audio": s.synth_normal("api_audio", (2, 2, 512, 128), seed=7),
        "depth": s.synth_normal("api_depth", (2, 1, 224, 224), seed=7),
    }


class _PassThrough:
    def __call__(self, x, device=None):
        return x.to(device) if device is not None else x

    def set_image_transform(self, t):
        self.image_transform = t


def main():
    assert ref_import.available(), "needs /root/reference"
    open_clip, fetch_model_cfg, _ = ref_import.import_reference()
    import mm_vit_lens.vitlens as RV

    def cfg_offline(*a, **k):
        cfg = fetch_model_cfg(*a, **k)
        cfg.pretrained = None
        for f in ("unlock_from_head", "vid_use_fpos", "vid_use_ltpos", "vid_distill_tokens"):
            setattr(cfg, f, False)
        return cfg

    RV.fetch_model_cfg = cfg_offline
    torch.set_num_threads(os.cpu_count())
    m = RV.ViTLens.__new__(RV.ViTLens)  # __init__ would import the file processors; everything below is the class's own code
    torch.nn.Module.__init__(m)
    m.model_var, m.modality_loaded = "vitlensL", list(MODALITIES)
    m.processors = {k: _PassThrough() for k in MODALITIES}
    m.vitlens = torch.nn.ModuleDict()
    for k in MODALITIES:
        m._init_modality_module(k)
    synth = C._synth()
    sd = synth.synth_state_dict(m.state_dict(), seed=3)
    m.load_state_dict(sd, strict=True)
    m.eval()
    inp = api_inputs()
    with torch.no_grad():
        out = m.encode(inp, normalize=True)
        raw = m.encode(inp, normalize=False)
    fx = {"chk_weights": torch.tensor(sum(float(v.double().abs().sum()) for v in sd.values() if v.is_floating_point())),
          "n_keys": torch.tensor(len(sd)), "keys_crc": torch.tensor(__import__("zlib").crc32("\n".join(sorted(sd)).encode()))}
    for k in MODALITIES:
        fx[k] = out[k].clone()
        fx["raw_" + k] = raw[k].clone()
        print(k, tuple(out[k].shape), float(out[k].norm(dim=-1).mean()), float(raw[k].norm(dim=-1).mean()))
    torch.save(fx, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; state_dict keys:", len(sd))


if __name__ == "__main__":
    main()
