"""vitlens_b200: hand-written sm_100a kernels (csrc/, C ABI in include/vitlens_b200.h) and the autograd
engine that the reference-compatible `open_clip` / `mm_vit_lens` packages in this directory run on."""
__version__ = "0.1.0"
