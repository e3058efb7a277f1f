"""Input pipelines feeding the towers, on the device: waveform -> AST log-mel clips, point-cloud
normalisation (+ FPS resampling), depth (disparity) normalisation.  Tensor in, tensor out; file decoding (torchaudio.load,
PIL, numpy loaders) stays with the caller."""
from __future__ import annotations

import math
from functools import lru_cache

import torch

from . import lib as L

AST_AS_MEAN, AST_AS_STD = -4.2677393, 4.5689974  # reference modal_audio/processors/at_processor.py (AudioSet statistics)
F32 = torch.float32


def _mel_scale(f):
    return 1127.0 * math.log(1.0 + f / 700.0)


@lru_cache(maxsize=8)
def _fbank_tables(device_str: str, sample_rate: float, frame_len: int, n_mel: int, low_freq: float, high_freq: float):
    """(hanning window [frame_len], mel matrix [n_mel, 257]) -- Kaldi's triangular filters on the HTK mel scale, built the way
    torchaudio.compliance.kaldi.get_mel_banks does (fp32 tensor arithmetic), padded with a zero column for the Nyquist bin."""
    dev = torch.device(device_str)
    n_fft = 512
    window = torch.hann_window(frame_len, periodic=False, dtype=F32)
    nyquist = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_rate / n_fft
    mel_low, mel_high = _mel_scale(low_freq), _mel_scale(high_freq)
    delta = (mel_high - mel_low) / (n_mel + 1)
    b = torch.arange(n_mel).unsqueeze(1)
    left, center, right = mel_low + b * delta, mel_low + (b + 1.0) * delta, mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + fft_bin_width * torch.arange(n_fft // 2) / 700.0).log()).unsqueeze(0)
    up, down = (mel - left) / (center - left), (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    bins = torch.nn.functional.pad(bins, (0, 1), value=0.0).to(F32).contiguous()
    return window.to(dev), bins.to(dev)


def fbank(waveform: torch.Tensor, *, sample_rate: float = 16000.0, n_mel: int = 128, target_length: int = 512, frame_length_ms: float = 25.0,
          frame_shift_ms: float = 10.0, preemphasis: float = 0.97, low_freq: float = 20.0, high_freq: float = 0.0, mean: float = AST_AS_MEAN,
          std: float = AST_AS_STD) -> torch.Tensor:
    """AudioASTProcessor*.convert2fbank + transform (at_processor.py:845-872) for a batch of clips: waveform fp32 [clips, samples]
    (or [samples]) on the device -> [clips, target_length, n_mel] normalised log-mel spectrograms, the layout AST_tokenizer takes.
    mean = 0, std = 1 gives the raw padded / cropped fbank."""
    assert waveform.is_cuda, "vitlens_b200 kernels need CUDA tensors (there is no CPU path)"
    w = waveform.to(F32)
    if w.dim() == 1:
        w = w[None]
    w = w.contiguous()
    frame_len = int(sample_rate * frame_length_ms * 0.001)
    shift = int(sample_rate * frame_shift_ms * 0.001)
    window, mel = _fbank_tables(str(w.device), float(sample_rate), frame_len, n_mel, float(low_freq), float(high_freq))
    out = torch.empty((w.shape[0], target_length, n_mel), device=w.device, dtype=F32)
    L.fbank(w, window, mel, out, clip_stride=w.stride(0), n_clips=w.shape[0], n_samples=w.shape[1], frame_len=frame_len, frame_shift=shift,
            n_mel=n_mel, preemph=preemphasis, target_len=target_length, mean=mean, std=std)
    return out


def pc_norm(pc: torch.Tensor) -> torch.Tensor:
    """pc_processor.pc_norm (pc_processor.py:32-38) per cloud: [B, N, C >= 3] (or [N, C]) -> xyz centred and scaled into the unit
    sphere, extra channels untouched."""
    assert pc.is_cuda
    x = pc.to(F32)
    single = x.dim() == 2
    if single:
        x = x[None]
    x = x.contiguous()
    out = torch.empty_like(x)
    L.pc_norm(x, out, B=x.shape[0], N=x.shape[1], C=x.shape[2])
    return out[0] if single else out


def pc_resample_fps(pc: torch.Tensor, npoints: int, start: torch.Tensor = None) -> torch.Tensor:
    """farthest_point_sample of the loader (pc_processor.py:8-29) on the device: [B, N, C] -> [B, npoints, C].  `start`: first
    index per cloud (the reference draws it with np.random.randint)."""
    from . import ops

    B, N, C = pc.shape
    if start is None:
        start = torch.randint(0, N, (B,), device=pc.device, dtype=torch.long)
    idx, _ = ops.fps(pc[..., :3].contiguous().float(), start, npoints)
    return torch.gather(pc, 1, idx.unsqueeze(-1).expand(-1, -1, C))


def depth_norm(depth: torch.Tensor, *, max_depth: float = 75.0, min_depth: float = 0.01, clamp_max_before_scale: bool = True,
               mean: float = 0.0418, std: float = 0.0295) -> torch.Tensor:
    """DepthNorm + Normalize of the disparity channel (transforms_rgbd.py:393-413; statistics vt_processor.py:170-171)."""
    assert depth.is_cuda
    x = depth.to(F32).contiguous()
    out = torch.empty_like(x)
    L.depth_norm(x, out, n=x.numel(), min_depth=min_depth, max_depth=max_depth, clamp_max=clamp_max_before_scale, mean=mean, std=std)
    return out
