"""vitlens_b200.train_step.TrainStep (training/train.py:131-249) on CPU with emulated kernels: the accum_freq > 1 scheme
(cached no-grad features, per-micro-batch re-forward) must reproduce the gradients of one step on the whole batch, the
gradient-norm clip and the logit_scale clamp must act as in the reference loop."""
import math

import torch

from tests.common import C, build_model


class _Recorder(torch.optim.SGD):
    """SGD with lr 0 that records the gradients it was handed."""

    def __init__(self, params):
        super().__init__(params, lr=0.0)
        self.seen = None

    def step(self):
        self.seen = {id(p): p.grad.detach().clone() for g in self.param_groups for p in g["params"] if p.grad is not None}
        return super().step()


def _setup(output_dict=True):
    import open_clip
    from vitlens_b200.train_step import TrainStep

    case = C.CASES["tiny_tri_audio"]
    model, sd, args = build_model(case)
    model.output_dict = True
    inp = C.build_inputs(case, args)
    params = [p for p in model.parameters() if p.requires_grad]
    return case, model, inp, params, open_clip, TrainStep


def test_accumulated_step_equals_whole_batch_step(emu):
    case, model, inp, params, open_clip, TrainStep = _setup()
    loss = open_clip.TriClipLoss()
    opt = _Recorder(params)
    TrainStep(model, loss, opt, accum_freq=1)(inp["image"], inp["text"], inp["visual"])
    whole = opt.seen
    opt2 = _Recorder(params)
    step = TrainStep(model, loss, opt2, accum_freq=2)
    B = case.batch
    assert not step(inp["image"][: B // 2], inp["text"][: B // 2], inp["visual"][: B // 2])
    assert opt2.seen is None
    assert step(inp["image"][B // 2:], inp["text"][B // 2:], inp["visual"][B // 2:])
    assert set(opt2.seen) == set(whole)
    for k, g in whole.items():
        if float(g.abs().max()) < 1e-6:
            continue
        # logit_scale is live in every one of the accum_freq re-forwards (train.py:183), so the reference loop -- and this one --
        # accumulates its gradient accum_freq times; every other parameter only sees its own micro-batch's features live
        mult = 2.0 if k == id(model.logit_scale) else 1.0
        c = float(torch.nn.functional.cosine_similarity(opt2.seen[k].flatten(), g.flatten(), dim=0))
        assert c > 0.99, c
        assert abs(float(opt2.seen[k].norm()) - mult * float(g.norm())) < 0.05 * mult * float(g.norm()) + 1e-6


def test_clip_and_logit_scale_clamp(emu):
    case, model, inp, params, open_clip, TrainStep = _setup()
    with torch.no_grad():
        model.logit_scale.fill_(9.0)  # above ln(100): must come back clamped
    opt = _Recorder(params)
    TrainStep(model, open_clip.TriClipLoss(), opt, grad_clip_norm=1e-3)(inp["image"], inp["text"], inp["visual"])
    total = math.sqrt(sum(float(g.pow(2).sum()) for g in opt.seen.values()))
    assert total <= 1e-3 * (1 + 1e-4)
    assert abs(float(model.logit_scale) - math.log(100)) < 1e-6
