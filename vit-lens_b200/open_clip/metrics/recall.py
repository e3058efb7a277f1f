"""Retrieval recall@1/5/10 in both directions (reference open_clip/metrics/recall.py): the [images x texts] similarity matrix
comes from the tcgen05 GEMM (fp32 output), the two top-10 rankings from vl_topk_rows; only the twelve counters reach the host."""
import torch

from vitlens_b200 import ops as _ops

from .base_metric import BaseMetric, all_gather_cat


class Recall(BaseMetric):
    def initialize(self, text_ids, text_logits):
        self.text_ids = text_ids
        self.text_logits = text_logits
        self.image_ids_list = []
        self.image_logits_list = []

    def compute(self, image_ids, image_logits):
        self.image_ids_list.append(image_ids)
        self.image_logits_list.append(image_logits)

    def merge_results(self, output_predict=False):
        self.image_ids = all_gather_cat(torch.cat(self.image_ids_list, dim=0))
        self.image_logits = all_gather_cat(torch.cat(self.image_logits_list, dim=0))
        sim_i2t = _ops.similarity(self.image_logits.float(), gallery=self.text_logits.float())
        sim_t2i = _ops.similarity(self.text_logits.float(), gallery=self.image_logits.float())  # == sim_i2t.t(), written row-major
        return self.retrieval_eval(sim_i2t, sim_t2i, output_predict)

    def retrieval_eval(self, scores_i2t, scores_t2i, output_predict=False):
        k_t, k_i = min(10, scores_i2t.size(1)), min(10, scores_t2i.size(1))
        rank_txt = _ops.topk_rows(scores_i2t.contiguous(), k_t).long()
        predict_txt = self.text_ids[None, :].expand(rank_txt.size(0), -1).gather(1, rank_txt)
        i2t = [predict_txt[:, :r].eq(self.image_ids[:, None]).any(1).sum().item() for r in (1, 5, 10)]
        rank_img = _ops.topk_rows(scores_t2i.contiguous(), k_i).long()
        predict_img = self.image_ids[None, :].expand(rank_img.size(0), -1).gather(1, rank_img)
        t2i = [predict_img[:, :r].eq(self.text_ids[:, None]).any(1).sum().item() for r in (1, 5, 10)]
        n_i, n_t = scores_i2t.size(0), scores_t2i.size(0)
        tr = [100.0 * c / n_i for c in i2t]
        ir = [100.0 * c / n_t for c in t2i]
        tr_mean, ir_mean = sum(tr) / 3, sum(ir) / 3
        predict_txt_results, predict_img_results = {}, {}
        if output_predict:
            predict_txt_results = dict(zip(self.image_ids.cpu().tolist(), predict_txt.cpu().tolist()))
            predict_img_results = dict(zip(self.text_ids.cpu().tolist(), predict_img.cpu().tolist()))
        return {"txt_r1": tr[0], "txt_r5": tr[1], "txt_r10": tr[2], "txt_r_mean": tr_mean, "img_count": n_i,
                "img_r1": ir[0], "img_r5": ir[1], "img_r10": ir[2], "img_r_mean": ir_mean, "r_mean": (tr_mean + ir_mean) / 2,
                "txt_count": n_t, "predict_txt": predict_txt_results, "predict_img": predict_img_results}
