"""Self-consistency of the restated torchlibrosa front-end (parity UNPINNED at this boundary:
SURVEY.md section 8c).  CPU only."""
import numpy as np
import torch

from oracle import frontend as fe
from oracle import sed


def test_mel_bank_facts():
    bank = fe.slaney_mel_filterbank(32000, 1024, 64, 50, 14000)
    assert bank.shape == (64, 513) and bank.dtype == np.float32
    nz = np.nonzero(bank)
    assert len(nz[0]) == 866                              # SURVEY.md Appendix A.3 probe facts
    assert nz[1].min() == 2 and nz[1].max() == 447
    assert (bank != 0).sum(axis=0).max() <= 2
    taps = (bank != 0).sum(axis=1)
    assert taps.min() == 3 and taps.max() == 47
    assert abs(float(bank.sum()) - 2.048126) < 1e-5
    assert abs(float(bank.max()) - 0.018150) < 1e-6


def test_dft_weights():
    wr, wi = fe.hann_dft_conv_weights(1024, 1024)
    assert wr.shape == wi.shape == (513, 1, 1024) and wr.dtype == np.float32
    n = np.arange(1024)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / 1024)
    for k in (0, 1, 7, 256, 512):
        np.testing.assert_allclose(wr[k, 0], win * np.cos(2 * np.pi * k * n / 1024), atol=1e-6)
        np.testing.assert_allclose(wi[k, 0], -win * np.sin(2 * np.pi * k * n / 1024), atol=1e-6)


def test_logmel_vs_float64_broadband():
    _, wave, _ = sed.synthetic_batch(2, 32000, seed=1234)
    wave = torch.from_numpy(wave)
    spec = fe.Spectrogram(n_fft=1024, hop_length=320, win_length=1024)
    mel = fe.LogmelFilterBank(sr=32000, n_fft=1024, n_mels=64, fmin=50, fmax=14000, top_db=None)
    with torch.no_grad():
        p = spec(wave)
        x = mel(p)
    assert p.shape == (2, 1, 101, 513) and x.shape == (2, 1, 101, 64)
    ref = fe.logmel_float64_reference(wave)
    err = (x[:, 0].double() - ref).abs().max().item()
    assert err < 1e-4, err                                 # fp32 noise floor ~1.5e-5 dB


def test_silence_is_minus_100_db():
    spec = fe.Spectrogram(n_fft=1024, hop_length=320, win_length=1024)
    mel = fe.LogmelFilterBank(sr=32000, n_fft=1024, n_mels=64, fmin=50, fmax=14000, top_db=None)
    with torch.no_grad():
        x = mel(spec(torch.zeros(1, 3200)))
    assert x.shape == (1, 1, 11, 64)
    assert torch.all(x == -100.0)


def test_specaug_rng_protocol_and_eval_noop():
    aug = fe.SpecAugmentation(64, 2, 8, 2)
    x = torch.ones(3, 1, 101, 64)
    aug.eval()
    assert aug(x) is x and torch.all(x == 1)
    aug.train()
    torch.manual_seed(1)
    y = aug(x.clone())
    torch.manual_seed(1)
    t = fe.draw_stripes(3, 101, 64, 2)       # all time stripes first ...
    f = fe.draw_stripes(3, 64, 8, 2)         # ... then all freq stripes
    exp = torch.ones(3, 1, 101, 64)
    for n in range(3):
        for s in range(2):
            exp[n, :, int(t[n, s, 0]):int(t[n, s, 0] + t[n, s, 1]), :] = 0
            exp[n, :, :, int(f[n, s, 0]):int(f[n, s, 0] + f[n, s, 1])] = 0
    assert torch.equal(y, exp)
    assert int(t[:, :, 1].max()) < 64 and int(f[:, :, 1].max()) < 8
