"""ViTLens.encode parity shared by the CPU (emulated kernels) and GPU tiers."""
import zlib

import torch

from tests.common import cosine


def check_vitlens_encode(device):
    from mm_vit_lens import ViTLens
    from oracle import cases as C
    from oracle.make_golden_api import MODALITIES, OUT, api_inputs
    from vitlens_b200 import synth

    gold = torch.load(OUT, map_location="cpu", weights_only=True)
    m = ViTLens(modality_loaded=list(MODALITIES), device="cpu")
    like = m.state_dict()
    assert len(like) == int(gold["n_keys"]) and zlib.crc32("\n".join(sorted(like)).encode()) == int(gold["keys_crc"]), "state_dict schema differs from the reference's ViTLens"
    sd = synth.synth_state_dict(like, seed=3)
    assert abs(sum(float(v.double().abs().sum()) for v in sd.values() if v.is_floating_point()) - float(gold["chk_weights"])) < 1e-6 * float(gold["chk_weights"])
    m.load_state_dict(sd, strict=True)
    m.eval().to(device)
    inp = api_inputs()
    rows = {}
    with torch.no_grad():
        out = m.encode(inp, normalize=True)
        raw = m.encode(inp, normalize=False)
    for k in MODALITIES:
        c, cr = cosine(out[k].cpu(), gold[k]), cosine(raw[k].cpu(), gold["raw_" + k])
        nr = float(raw[k].float().norm()) / float(gold["raw_" + k].norm())
        rows["cos_" + k], rows["cos_raw_" + k], rows["norm_ratio_raw_" + k] = round(c, 6), round(cr, 6), round(nr, 5)
        assert tuple(out[k].shape) == tuple(gold[k].shape)
        assert c > 0.999 and cr > 0.999 and abs(nr - 1) < 2e-2, (k, c, cr, nr)
        assert float((out[k].float().norm(dim=-1) - 1).abs().max()) < 1e-4
    return rows
