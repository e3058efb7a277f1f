"""How far does the REFERENCE's own mixed-precision path drift from its fp32 path?  Runs the CPU oracle under
torch.autocast(bfloat16) (the arithmetic `--precision amp_bf16` gives the reference: bf16 GEMM operands / outputs, fp32
LayerNorm / softmax / residual adds through autocast's op lists) against the fp32 golden fixtures and prints one JSON line per
case -- the yardstick for this repo's bf16 tolerances (SURVEY 7 / VERDICT r1 item 9).  CPU only, test infrastructure.

    python tools/autocast_deviation.py [case ...] > profiles/r02_autocast_deviation.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.common import C, build_model, cosine, run_oracle  # noqa: E402


def main(names):
    for name in names:
        case = C.CASES[name]
        gold = C.load_golden(name)
        model, sd, args = build_model(case)
        inp = C.build_inputs(case, args)
        if "fps_start" in gold:
            inp["fps_start"] = gold["fps_start"]
        keys = sorted(k for k, p in model.named_parameters() if p.requires_grad)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            feats, ls, loss, grads = run_oracle(case, sd, args, inp, set(keys))
        row = {"case": name, "vs": "fp32 reference fixture", "what": "oracle under torch.autocast(bfloat16)"}
        for k, v in feats.items():
            row["cos_" + k] = round(cosine(v.detach().float(), gold[k]), 6)
        row["loss"], row["loss_ref"] = float(loss), float(gold["loss"])
        worst, wn = (1.0, ""), (0.0, "")
        for i, k in enumerate(keys):
            gn = float(gold["grad_norms"][i])
            if gn > 1e-4 and k != "logit_scale":
                wn = max(wn, (abs(float(grads[k].float().norm()) - gn) / gn, k))
            if "grad:" + k in gold and gn > 1e-4 and gold["grad:" + k].numel() > 1:
                worst = min(worst, (cosine(grads[k].float(), gold["grad:" + k]), k))
        row.update(min_grad_cos=round(worst[0], 5), min_grad_cos_key=worst[1], max_grad_norm_relerr=round(wn[0], 5), max_grad_norm_key=wn[1])
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["tiny_clip", "tiny_tri_audio", "tiny_tri_depth", "tiny_tri_pc", "tiny_tri_pc_bntrain", "vitb32_clip_bs8",
                          "vitl14_audio128_bs2", "vitl14_depth_bs2"])
