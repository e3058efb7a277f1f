"""Point-cloud tokenizer (reference modal_3d/models/pointbert/point_encoder.py:299-362, dvae.py:107-212,
misc.py:48-68): FPS -> kNN grouping -> mini-PointNet -> Linear, pos = MLP(centres).

Parameter tree and state_dict keys match the reference.  The forward runs vl_fps / vl_knn_group / vl_linear3 /
vl_group_max and the tcgen05 GEMM (BatchNorm folded into the neighbouring 1x1 convs), forward and backward
(engine.PointTokenizerFn).  BatchNorm follows nn.BatchNorm1d: running statistics in eval mode (an affine folded into the
convs), batch statistics + running-statistics update in training mode, and cross-rank statistics when the layers were
converted with torch.nn.SyncBatchNorm.convert_sync_batchnorm (--use-bn-sync, pc_tri_main.py:372-373)."""
import torch
import torch.nn as nn

from vitlens_b200 import engine as E

from ....transformer import TokenMat
from ....util.Sample import Sample


class Encoder(nn.Module):
    def __init__(self, encoder_channel):
        super().__init__()
        self.encoder_channel = encoder_channel
        self.first_conv = nn.Sequential(nn.Conv1d(3, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True), nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True), nn.Conv1d(512, self.encoder_channel, 1))


class Group(nn.Module):
    def __init__(self, num_group, group_size):
        super().__init__()
        self.num_group = num_group
        self.group_size = group_size


class PointTokenizer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.trans_dim = config.trans_dim
        self.group_size = config.group_size
        self.num_group = config.num_group
        self.group_divider = Group(num_group=self.num_group, group_size=self.group_size)
        self.encoder_dims = config.encoder_dims
        self.encoder = Encoder(encoder_channel=self.encoder_dims)
        self.reduce_dim = nn.Linear(self.encoder_dims, self.trans_dim)
        self.pos_embed = nn.Sequential(nn.Linear(3, 128), nn.GELU(), nn.Linear(128, self.trans_dim))

    def forward(self, pts, fps_start=None):
        """pts [B, N, 3] fp32.  `fps_start` [B] int64: first FPS index per sample (the reference draws it with
        torch.randint, misc.py:60); drawn the same way when omitted."""
        if self.group_size > 32:
            raise NotImplementedError("group_size > 32")
        B, N, _ = pts.shape
        if fps_start is None:
            fps_start = torch.randint(0, N, (B,), dtype=torch.long, device=pts.device)
        tok, pos = E.point_tokenizer_forward(self, pts.float(), fps_start.to(pts.device))
        return Sample({"x": TokenMat(tok, B, self.num_group), "pos": TokenMat(pos, B, self.num_group)})
