"""Point-cloud tokenizer (reference modal_3d/models/pointbert/point_encoder.py:299-362, dvae.py:107-212,
misc.py:48-68): FPS -> kNN grouping -> mini-PointNet -> Linear, pos = MLP(centres).

Parameter tree and state_dict keys match the reference.  The FPS / kNN / grouped-PointNet kernels
are the next row of the coverage table (BASELINE config 5); until they land the
forward raises instead of silently running a non-native path."""
import torch
import torch.nn as nn


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
        raise NotImplementedError(
            "PointTokenizer.forward: FPS / kNN / grouped-PointNet sm_100a kernels are not built yet "
            "(next coverage row, DESIGN.md); there is deliberately no PyTorch fallback.")
