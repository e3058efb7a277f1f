"""Mean average precision (reference open_clip/metrics/map.py): logits and multi-hot targets are gathered on the device and the
per-class average precision -- sklearn.metrics.average_precision_score's definition, ties included -- is computed by
vl_average_precision (one CTA per class) instead of a host round trip through numpy / scikit-learn."""
import torch

from vitlens_b200 import ops as _ops

from .base_metric import BaseMetric, all_gather_cat


class MAP(BaseMetric):
    def initialize(self, device="cuda"):
        self.logits = torch.zeros(0, device=device, dtype=torch.float32)
        self.targets = torch.zeros(0, device=device, dtype=torch.float32)
        self.ids = torch.zeros(0, device=device, dtype=torch.long)

    def compute(self, ids, logits, targets):
        self.ids = torch.cat([self.ids, ids], dim=0)
        self.logits = torch.cat([self.logits, logits.float()], dim=0)
        self.targets = torch.cat([self.targets, targets.float()], dim=0)

    def merge_results(self, output_predict=False):
        ids, preds, targets = all_gather_cat(self.ids), all_gather_cat(self.logits), all_gather_cat(self.targets)
        if targets.ndim != preds.ndim:
            targets = targets.reshape(preds.shape)
        ap, npos = _ops.average_precision(preds, targets, apply_sigmoid=True)  # map.py:36 applies a sigmoid before scoring
        predict_results = {}
        if output_predict:
            predict_results = dict(zip(ids.cpu().tolist(), torch.sigmoid(preds).cpu().tolist()))
        return {"map": float(ap.double().mean()), "map_cnt": int(targets.shape[0]), "predict_results": predict_results,
                "ap_per_class": ap, "positives_per_class": npos}
