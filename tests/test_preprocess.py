"""Input pipelines: the oracle's restatement of Kaldi's fbank against torchaudio's golden output (and against
torchaudio itself where importable), then the CUDA kernels against the oracle."""
import pytest
import torch

from tests.common import C, O, ROOT  # noqa: F401


def _golden():
    import os

    return torch.load(os.path.join(C.GOLDEN_DIR, "fbank.pt"), map_location="cpu", weights_only=True)


def test_oracle_fbank_matches_torchaudio_golden():
    from oracle.make_golden_fbank import waveform

    gold = _golden()["fbank"]
    ours = O.kaldi_fbank(waveform())
    assert ours.shape == gold.shape
    assert float((ours - gold).abs().max()) < 2e-3  # log-mel values in [-16, 16]; fp64 restatement vs torchaudio's fp32 FFT
    try:
        import torchaudio
    except Exception:
        return
    w = waveform(seed=3, seconds=0.7)
    ref = torchaudio.compliance.kaldi.fbank(w[None], htk_compat=True, sample_frequency=16000, use_energy=False, window_type="hanning",
                                            num_mel_bins=128, dither=0.0, frame_shift=10)
    assert float((O.kaldi_fbank(w) - ref).abs().max()) < 2e-3


def test_oracle_pc_and_depth_norm():
    g = torch.Generator().manual_seed(0)
    pc = torch.randn(500, 6, generator=g) * 3 + 1
    out = O.pc_norm(pc)
    assert float(out[:, :3].mean(0).abs().max()) < 1e-5 and abs(float(out[:, :3].norm(dim=1).max()) - 1) < 1e-5
    assert torch.equal(out[:, 3:], pc[:, 3:])
    d = torch.tensor([-1.0, 0.0, 0.5, 75.0, 200.0])
    want = (torch.tensor([0.01, 0.01, 0.5, 75.0, 75.0]) / 75.0 - 0.0418) / 0.0295
    assert torch.allclose(O.depth_norm(d), want)


@pytest.mark.gpu
def test_fbank_kernel_vs_oracle_and_torchaudio():
    from oracle.make_golden_fbank import waveform
    from vitlens_b200 import preprocess as P

    gold = _golden()["fbank"]
    w = waveform().cuda()
    raw = P.fbank(w, target_length=gold.shape[0], mean=0.0, std=1.0)[0].cpu()
    assert float((raw - gold).abs().max()) < 2e-3, float((raw - gold).abs().max())
    # batch of clips, pad (target longer than the clip) and crop (shorter), AST normalisation
    clips = torch.stack([waveform(seed=s, seconds=1.3) for s in (1, 2, 3)]).cuda()
    for tl in (512, 64):
        got = P.fbank(clips, target_length=tl).cpu()
        for i in range(3):
            want = O.ast_clip(clips[i].cpu(), target_length=tl)
            assert float((got[i] - want).abs().max()) < 1e-3, (tl, i, float((got[i] - want).abs().max()))
    assert torch.equal(P.fbank(clips, target_length=512), P.fbank(clips, target_length=512))  # deterministic


@pytest.mark.gpu
def test_pc_and_depth_kernels_vs_oracle():
    from vitlens_b200 import preprocess as P

    g = torch.Generator().manual_seed(1)
    pc = (torch.randn(3, 4000, 6, generator=g) * 2 + 0.5)
    got = P.pc_norm(pc.cuda()).cpu()
    for b in range(3):
        assert float((got[b] - O.pc_norm(pc[b])).abs().max()) < 1e-5
    start = torch.tensor([5, 17, 0])
    sub = P.pc_resample_fps(pc.cuda(), 256, start.cuda()).cpu()
    idx = O.fps_indices(pc[..., :3], 256, start)
    assert torch.equal(sub, torch.gather(pc, 1, idx.unsqueeze(-1).expand(-1, -1, 6)))  # bit-exact FPS indices
    d = torch.randn(2, 1, 224, 224, generator=g) * 60
    assert float((P.depth_norm(d.cuda()).cpu() - O.depth_norm(d)).abs().max()) < 1e-5
