class BaseMetric(object):
    """reference open_clip/metrics/base_metric.py"""

    def initialize(self):
        raise NotImplementedError

    def compute(self, *args, **kwargs):
        raise NotImplementedError

    def merge_results(self, output_predict=False):
        raise NotImplementedError


def all_gather_cat(t):
    """Concatenate a tensor over ranks (sizes may differ per rank), reference open_clip/utils.py:295-330 `all_gather`."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.long)
    sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: int(s)] for b, s in zip(bufs, sizes)], dim=0)
