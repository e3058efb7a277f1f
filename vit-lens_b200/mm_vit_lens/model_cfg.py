"""Per-modality configuration of the released ViT-Lens-L models with the reference's field names
(reference mm_vit_lens/model_cfg.py:9-197).  These are hyper-parameter tables, not code."""
from copy import deepcopy

from open_clip.constants import CKPT_CACHE_DIR
from open_clip.module_cfg import AttrDict

default_cfg = AttrDict(
    audio_clip_duration=5.0,
    audio_fstride=10,
    audio_mel_bins=128,
    audio_sampling_rate=16000,
    audio_target_length=512,
    audio_tstride=10,
    aug_cfg={},
    cache_dir=CKPT_CACHE_DIR,
    dataset_type='image',
    device='cpu',
    disable_orig_pos=False,
    disable_pt_vit=False,
    disable_visual_adapter_pos=False,
    eeg_chans=128,
    eeg_stride=1,
    eeg_time_len=512,
    eeg_window_size=1,
    force_custom_text=False,
    force_image_size=None,
    force_patch_dropout=None,
    force_quick_gelu=False,
    image_mean=None,
    image_std=None,
    load_ckpt_strict=False,
    model='ViT-L-14',
    pc_encoder_dims=256,
    pc_group_size=32,
    pc_in_channel=3,
    pc_npoints=8192,
    pc_num_group=512,
    pc_radius=0.2,
    pc_tokenizer='pointbert',
    pc_trans_dim=384,
    perceiver_as_identity=False,
    perceiver_as_transformer=False,
    perceiver_attn_dropout=0.0,
    perceiver_cross_dim_head=64,
    perceiver_cross_heads=1,
    perceiver_depth=1,
    perceiver_ff_dropout=0.0,
    perceiver_fourier_encode_data=False,
    perceiver_input_axis=1,
    perceiver_input_chan=1024,
    perceiver_latent_dim=1024,
    perceiver_latent_dim_head=64,
    perceiver_latent_heads=16,
    perceiver_max_freq=10.0,
    perceiver_num_classes=1000,
    perceiver_num_freq_bands=32,
    perceiver_num_latents=256,
    perceiver_self_per_cross_attn=1,
    perceiver_weight_tie_layers=False,
    precision='fp32',
    pretrained='datacomp_xl_s13b_b90k',
    pretrained_image=False,
    skip_trans_first_n_layers=None,
    torchcompile=False,
    torchscript=False,
    trace=False,
    use_bn_sync=False,
    use_bnb_linear=None,
    use_eva_pt_lin=False,
    use_openclip_transform=False,
    use_perceiver=False,
    use_visual_adapter=False,
    v_key='image',
    visual_arch='perceiver_vit',
    visual_modality_type='image',
)

_MODALITIES_L = AttrDict(
    pc=AttrDict(
        ckpt_pth='/PATH_TO/vitlensL_pc.pt',
        pc_encoder_dims=256,
        pc_group_size=32,
        pc_npoints=8192,
        pc_num_group=512,
        pc_trans_dim=384,
        perceiver_attn_dropout=0.0,
        perceiver_cross_dim_head=64,
        perceiver_cross_heads=1,
        perceiver_depth=4,
        perceiver_ff_dropout=0.0,
        perceiver_fourier_encode_data=False,
        perceiver_input_axis=1,
        perceiver_input_chan=384,
        perceiver_latent_dim=1024,
        perceiver_latent_dim_head=64,
        perceiver_latent_heads=16,
        perceiver_num_latents=256,
        perceiver_self_per_cross_attn=1,
        perceiver_weight_tie_layers=False,
        use_perceiver=True,
        use_visual_adapter=True,
        v_key='pc',
        visual_modality_type='3dpc',
    ),
    audio=AttrDict(
        audio_clip_duration=5.0,
        audio_fstride=10,
        audio_mel_bins=128,
        audio_sampling_rate=16000,
        audio_target_length=512,
        audio_tstride=10,
        ckpt_pth='/PATH_TO/vitlensL_audio.pt',
        perceiver_attn_dropout=0.0,
        perceiver_cross_dim_head=64,
        perceiver_cross_heads=1,
        perceiver_depth=2,
        perceiver_ff_dropout=0.0,
        perceiver_fourier_encode_data=False,
        perceiver_input_axis=1,
        perceiver_input_chan=1024,
        perceiver_latent_dim=1024,
        perceiver_latent_dim_head=64,
        perceiver_latent_heads=16,
        perceiver_num_latents=256,
        perceiver_self_per_cross_attn=3,
        perceiver_weight_tie_layers=False,
        use_perceiver=True,
        use_visual_adapter=True,
        v_key='audio',
        visual_modality_type='audio',
    ),
    depth=AttrDict(
        ckpt_pth='/PATH_TO/vitlensL_depth.pt',
        perceiver_as_identity=True,
        use_perceiver=True,
        use_visual_adapter=True,
        v_key='depth',
        visual_modality_type='depth',
    ),
    tactile=AttrDict(
        ckpt_pth='/PATH_TO/vitlensL_tactile.pt',
        use_perceiver=False,
        use_visual_adapter=False,
        v_key='tactile',
        visual_modality_type='tactile',
    ),
    eeg=AttrDict(
        ckpt_pth='/PATH_TO/vitlensL_eeg.pt',
        eeg_chans=128,
        eeg_stride=1,
        eeg_time_len=512,
        eeg_window_size=1,
        perceiver_as_transformer=False,
        perceiver_attn_dropout=0.0,
        perceiver_cross_dim_head=64,
        perceiver_cross_heads=1,
        perceiver_depth=1,
        perceiver_ff_dropout=0.0,
        perceiver_fourier_encode_data=False,
        perceiver_input_axis=1,
        perceiver_input_chan=1024,
        perceiver_latent_dim=1024,
        perceiver_latent_dim_head=64,
        perceiver_latent_heads=16,
        perceiver_max_freq=10.0,
        perceiver_self_per_cross_attn=1,
        perceiver_weight_tie_layers=False,
        use_perceiver=True,
        use_visual_adapter=True,
        v_key='eeg',
        visual_modality_type='eeg',
    ),
)

vitlens_model_cfg = AttrDict(
    vitlensL=AttrDict(model='ViT-L-14', pretrained='datacomp_xl_s13b_b90k', **_MODALITIES_L),
    vitlensB=None,
)


def fetch_model_cfg(model_keys=("model", "pretrained"), modality="pc", model_option="vitlensL"):
    """Defaults overlaid with the model keys and the modality block (model_cfg.py:185-197)."""
    base_cfg = deepcopy(default_cfg)
    model_cfg = vitlens_model_cfg[model_option]
    for k in model_keys:
        setattr(base_cfg, k, model_cfg[k])
    if modality not in ("image", "video", "text"):
        base_cfg.update(model_cfg[modality])
    return base_cfg


def training_args(modality="pc", model_option="vitlensL", **overrides):
    """fetch_model_cfg + the four flags the training stack adds (training/params.py:339,849-867) that
    VisionTransformer.lock / TriCLIP.forward read; pretrained=None (no network)."""
    cfg = fetch_model_cfg(modality=modality, model_option=model_option)
    cfg.pretrained = None
    for k in ("unlock_from_head", "vid_use_fpos", "vid_use_ltpos", "vid_distill_tokens"):
        setattr(cfg, k, False)
    cfg.update(overrides)
    return cfg
