"""Config helpers with the reference's field names (open_clip/module_cfg.py:12-92).  `args` is any
attribute object (argparse Namespace / EasyDict / mm_vit_lens.model_cfg.AttrDict)."""
import copy


class AttrDict(dict):
    """dict with attribute access (stands in for easydict.EasyDict, which is not a dependency here)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


edict = AttrDict

_IMAGE_DEFAULTS = {
    "use_perceiver": False,
    "use_visual_adapter": False,
    "visual_modality_type": "image",
    "perceiver_cfg": None,
    "visual_adapter_cfg": None,
    "unlock_cls": False,
    "skip_trans_first_n_layers": None,
    "unlock_trans_first_n_layers": None,
}


def get_default_image_cfg():
    return AttrDict({k: _IMAGE_DEFAULTS[k] for k in ("use_perceiver", "use_visual_adapter", "visual_modality_type", "perceiver_cfg", "visual_adapter_cfg")})


def set_default_image_cfg(cfg):
    """Plain-CLIP image tower settings for TriCLIP.image (module_cfg.py:17-34)."""
    cfg_ = copy.deepcopy(cfg)
    if isinstance(cfg_, dict):
        cfg_.update(_IMAGE_DEFAULTS)
    else:
        for k, v in _IMAGE_DEFAULTS.items():
            if hasattr(cfg_, k):
                setattr(cfg_, k, v)
    return cfg_


def get_perceiver_cfg(args):
    keys = ["input_chan", "input_axis", "num_freq_bands", "max_freq", "depth", "num_latents", "latent_dim", "cross_heads",
            "latent_heads", "cross_dim_head", "latent_dim_head", "num_classes", "attn_dropout", "ff_dropout",
            "weight_tie_layers", "fourier_encode_data", "self_per_cross_attn"]
    cfg = AttrDict(use_perceiver=args.use_perceiver)
    for k in keys:
        cfg[k] = getattr(args, "perceiver_" + k)
    return cfg


def get_input_adapter_cfg(args):
    cfg = AttrDict(use_visual_adapter=args.use_visual_adapter, visual_modality_type=args.visual_modality_type,
                   disable_orig_pos=args.disable_orig_pos)
    vt = args.visual_modality_type
    if vt in ("3dpc", "pc", "point cloud", "pointcloud"):
        cfg.pc_tokenizer = args.pc_tokenizer
        cfg.trans_dim = args.pc_trans_dim
        cfg.group_size = args.pc_group_size
        cfg.num_group = args.pc_num_group
        cfg.encoder_dims = args.pc_encoder_dims
        cfg.radius = args.pc_radius
        cfg.in_dim = args.pc_in_channel
    elif vt in ("image", "3dpc_raw", "video", "depth", "audio", "tactile", "eeg"):
        pass
    else:
        raise NotImplementedError(vt)
    return cfg
