"""Dispatcher for the per-modality patch / token embeds (reference open_clip/visual_adapter.py:7-69)."""
from torch import nn


def get_visual_adapter(cfg, **kwargs):
    vtype = cfg.visual_modality_type
    if vtype in ("3dpc", "pc", "pointcloud", "point_cloud", "point cloud"):
        if cfg.pc_tokenizer == "pointbert":
            from .modal_3d.models.pointbert.point_encoder import PointTokenizer

            return PointTokenizer(config=cfg)
        # 'pnsa' (pointnet_util.py:345-368) is unreachable in the reference too: PointNSATokenizer.forward needs xyz=, which
        # TriCLIP.encode_visual never passes (model.py:524-525)
        raise NotImplementedError(f"pc_tokenizer={cfg.pc_tokenizer!r}: only 'pointbert' (vitlensL) is on the covered path")
    if vtype == "3dpc_raw":
        return nn.Identity()
    if vtype == "depth":
        from .modal_depth.models.DepthTokenizer import DepthTokenizer

        return DepthTokenizer(grid_size=kwargs["grid_size"], patch_size=kwargs["patch_size"], width=kwargs["width"],
                              input_patchnorm=kwargs["input_patchnorm"])
    if vtype == "audio":
        from .modal_audio.models.AST_tokenizer import AST_tokenizer

        exp_args = kwargs["exp_args"]
        return AST_tokenizer(fstride=exp_args.audio_fstride, tstride=exp_args.audio_tstride, input_fdim=exp_args.audio_mel_bins,
                             input_tdim=exp_args.audio_target_length, patch_size=kwargs["patch_size"], width=kwargs["width"])
    if vtype == "tactile":
        return None
    if vtype == "eeg":
        from .modal_eeg.models.EEG_tokenizer import PatchEmbed1D

        exp_args = kwargs["exp_args"]
        return PatchEmbed1D(time_len=exp_args.eeg_time_len, in_chans=exp_args.eeg_chans, window_size=exp_args.eeg_window_size,
                            stride=exp_args.eeg_stride, width=kwargs["width"])
    raise NotImplementedError(vtype)
