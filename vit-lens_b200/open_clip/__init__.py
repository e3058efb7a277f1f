"""Drop-in `open_clip` surface of TencentARC/ViT-Lens for its data-parallel hot path, running on
hand-written sm_100a kernels (vitlens_b200).  Export list follows reference open_clip/__init__.py:1-48
for the in-scope symbols."""
from .constants import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, ModalityType
from .factory import (add_model_config, create_loss, create_model, create_model_and_transforms, get_model_config,
                      get_tokenizer, list_models, load_checkpoint, tri_create_model, tri_create_model_and_transforms)
from .loss import ClipLoss, ClipLossGeneral, ClipLossLabelMask, ClipLossSimMask, TriClipLoss, TriClipLossLabelMask, gather_features
from .model import (CLIP, CLIPTextCfg, CLIPVisionCfg, TriCLIP, convert_weights_to_lp, get_cast_dtype, get_input_dtype,
                    trace_model)
from .tokenizer import tokenize
from .transform import AugmentationCfg, image_transform
from .zero_shot_classifier import build_zero_shot_classifier, build_zero_shot_classifier_legacy

__version__ = "2.20.0+vitlens_b200"
