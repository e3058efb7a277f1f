from .vitlens import ViTLens  # noqa: F401  (reference mm_vit_lens/__init__.py:1)
