"""Top-1 accuracy accumulated on the device (reference open_clip/metrics/accuracy.py): counters stay on the GPU, one
all-reduce at merge time, no host sync per batch."""
import torch
import torch.distributed as dist

from vitlens_b200 import ops as _ops

from .base_metric import BaseMetric, all_gather_cat


class Accuracy(BaseMetric):
    def initialize(self, device="cuda"):
        self.score_sum = torch.zeros(1, device=device, dtype=torch.float32)
        self.score_cnt = torch.zeros(1, device=device, dtype=torch.int32)
        self.ids = torch.zeros(0, device=device, dtype=torch.long)
        self.hyps = torch.zeros(0, device=device, dtype=torch.long)

    def compute(self, ids, logits, targets):
        predict_labels = _ops.topk_rows(logits.float().contiguous(), 1)[:, 0].long()
        if targets.dim() == 2:  # multi-hot targets: correct when the predicted class is one of the labels
            n_correct = targets.gather(1, predict_labels.unsqueeze(1)).sum()
        else:
            n_correct = predict_labels.eq(targets).sum()
        self.score_sum += n_correct
        self.score_cnt += logits.size(0)
        self.ids = torch.cat([self.ids, ids], dim=0)
        self.hyps = torch.cat([self.hyps, predict_labels], dim=0)

    def merge_results(self, output_predict=False):
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(self.score_sum, op=dist.ReduceOp.SUM)
            dist.all_reduce(self.score_cnt, op=dist.ReduceOp.SUM)
        ids, hyps = all_gather_cat(self.ids), all_gather_cat(self.hyps)
        predict_results = dict(zip(ids.cpu().tolist(), hyps.cpu().tolist())) if output_predict else {}
        score_sum, score_cnt = self.score_sum.item(), self.score_cnt.item()
        return {"accuracy": score_sum / score_cnt, "score_sum": score_sum, "score_cnt": score_cnt, "predict_results": predict_results}
