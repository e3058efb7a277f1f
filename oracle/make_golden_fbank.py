"""Golden log-mel vectors from torchaudio.compliance.kaldi.fbank with the reference's call-site arguments
(modal_audio/processors/at_processor.py:854-863)  --  TEST INFRASTRUCTURE.  python oracle/make_golden_fbank.py"""
import os
import sys

import torch
import torchaudio

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases as C  # noqa: E402

OUT = os.path.join(C.GOLDEN_DIR, "fbank.pt")


def waveform(seed=0, seconds=1.3):
    g = torch.Generator().manual_seed(seed)
    n = int(16000 * seconds)
    t = torch.arange(n) / 16000.0
    w = 0.3 * torch.sin(2 * torch.pi * 440.0 * t) + 0.1 * torch.sin(2 * torch.pi * 3000.0 * t * (1 + 0.2 * t)) + 0.05 * torch.randn(n, generator=g)
    return w - w.mean()  # audio_get_clip(sub_mean=True), at_processor.py:221-222


if __name__ == "__main__":
    w = waveform()
    fb = torchaudio.compliance.kaldi.fbank(w[None], htk_compat=True, sample_frequency=16000, use_energy=False, window_type="hanning",
                                           num_mel_bins=128, dither=0.0, frame_shift=10)
    torch.save({"fbank": fb.clone(), "torchaudio": torchaudio.__version__}, OUT)
    print(tuple(fb.shape), "wrote", OUT, os.path.getsize(OUT))
