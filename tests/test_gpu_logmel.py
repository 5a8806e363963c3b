"""-m gpu: the fused log-mel kernel (through the C ABI) against the CPU oracle.
Tolerance (BASELINE.json north_star): log-mel <= 1e-4 dB abs on broadband input."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_DB = 1e-4


def _oracle_logmel(wave):
    from oracle import frontend as ofe
    spec = ofe.Spectrogram(n_fft=1024, hop_length=320, win_length=1024)
    mel = ofe.LogmelFilterBank(sr=32000, n_fft=1024, n_mels=64, fmin=50, fmax=14000, top_db=None)
    with torch.no_grad():
        p = spec(wave)
        return p, mel(p)


def _bank(dev):
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    melW = torch.from_numpy(fe.mel_weight_matrix(32000, 1024, 64, 50, 14000)).to(dev)
    return fe.MelBankCSR(melW)


@pytest.mark.parametrize('n_samples', [32000, 320000, 513 + 7, 31999])
def test_logmel_f32_matches_oracle(n_samples):
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    _, wave, _ = sed.synthetic_batch(3, n_samples, seed=1234)
    wave = torch.from_numpy(wave)
    _, ref = _oracle_logmel(wave)
    got = fe.logmel(wave.cuda(), 320, _bank('cuda')).cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    assert err <= TOL_DB, err


def test_logmel_int16_input_equals_float_path():
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    pcm, wave, _ = sed.synthetic_batch(2, 64000, seed=7)
    a = fe.logmel(torch.from_numpy(wave).cuda(), 320, _bank('cuda'))
    b = fe.logmel(torch.from_numpy(pcm).cuda(), 320, _bank('cuda'))
    assert torch.equal(a, b)            # x/32767 fused into the gather is bit-identical


def test_int16_conversion_on_device_equals_reference_for_every_value():
    """Every int16 value goes through the kernel's x/32767 (csrc/logmel.cu Sample<int16_t>::cvt, a multiply + two
    FMAs): the log-mel of PCM input is bit-identical to that of the reference-converted fp32 waveform
    (utils/utilities.py:66-67) on clips that contain all 65536 values, in order and shuffled."""
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    allv = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    rs = np.random.RandomState(3)
    pcm = np.stack([allv, rs.permutation(allv), allv[::-1].copy()])
    wave = (pcm / 32767.).astype(np.float32)
    a = fe.logmel(torch.from_numpy(wave).cuda(), 320, _bank('cuda'))
    b = fe.logmel(torch.from_numpy(pcm).cuda(), 320, _bank('cuda'))
    assert torch.equal(a, b)


def test_logmel_gaussian_input():
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    rs = np.random.RandomState(5)
    pcm = np.clip(np.round(rs.randn(2, 96000) * 0.1 * 32767), -32768, 32767).astype(np.int16)
    wave = torch.from_numpy((pcm / 32767.).astype(np.float32))
    _, ref = _oracle_logmel(wave)
    got = fe.logmel(wave.cuda(), 320, _bank('cuda')).cpu()
    assert (got - ref).abs().max().item() <= TOL_DB


def test_silence_is_exactly_minus_100_db():
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    got = fe.logmel(torch.zeros(2, 6400, device='cuda'), 320, _bank('cuda'))
    assert got.shape == (2, 1, 21, 64)
    assert torch.all(got == -100.0)


def test_empty_batch_and_bad_args():
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    out = fe.logmel(torch.zeros(0, 6400, device='cuda'), 320, _bank('cuda'))
    assert out.shape == (0, 1, 21, 64)
    with pytest.raises(RuntimeError):
        fe.logmel(torch.zeros(1, 400, device='cuda'), 320, _bank('cuda'))     # shorter than the reflect pad
    with pytest.raises(RuntimeError):
        fe.logmel(torch.zeros(1, 6400), 320, _bank('cpu'))                    # no CPU path


def test_unfused_seam_power_and_mel():
    """Spectrogram alone (513-bin power) and LogmelFilterBank alone, as at models.py:199-200."""
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import frontend as fe
    _, wave, _ = sed.synthetic_batch(2, 32000, seed=3)
    wave = torch.from_numpy(wave)
    p_ref, x_ref = _oracle_logmel(wave)
    p = fe.stft_power(wave.cuda(), 320)
    assert p.shape == p_ref.shape
    rel = ((p.cpu() - p_ref).abs() / p_ref.abs().clamp_min(1e-3)).max().item()
    assert rel < 2e-4, rel
    x = fe.mel_db(p, _bank('cuda')).cpu()
    assert (x - x_ref).abs().max().item() <= TOL_DB


def test_torchlibrosa_dropin_modules_fuse_and_materialize():
    import os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), 'sound_event_detection_dcase2017_task4_b200', 'dropin'))
    for k in [k for k in sys.modules if k == 'torchlibrosa' or k.startswith('torchlibrosa.')]:
        del sys.modules[k]
    from torchlibrosa.stft import Spectrogram, LogmelFilterBank, LazyPowerSpectrogram
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import _lib
    _, wave, _ = sed.synthetic_batch(2, 32000, seed=11)
    wave = torch.from_numpy(wave)
    p_ref, x_ref = _oracle_logmel(wave)
    spec = Spectrogram(n_fft=1024, hop_length=320, win_length=1024, window='hann', center=True,
                       pad_mode='reflect', freeze_parameters=True).cuda()
    mel = LogmelFilterBank(sr=32000, n_fft=1024, n_mels=64, fmin=50, fmax=14000, ref=1.0, amin=1e-10,
                           top_db=None, freeze_parameters=True).cuda()
    assert sorted(spec.state_dict()) == ['stft.conv_imag.weight', 'stft.conv_real.weight']
    assert list(mel.state_dict()) == ['melW']
    n0 = _lib.launch_count()
    s = spec(wave.cuda())
    assert isinstance(s, LazyPowerSpectrogram) and s.shape == (2, 1, 101, 513)
    x = mel(s)
    assert _lib.launch_count() - n0 == 1                   # ONE fused kernel for both module calls
    assert (x.cpu() - x_ref).abs().max().item() <= TOL_DB
    dense = spec(wave.cuda()) * 1.0                        # any other use materialises the spectrogram
    assert type(dense) is torch.Tensor
    rel = ((dense.cpu() - p_ref).abs() / p_ref.abs().clamp_min(1e-3)).max().item()
    assert rel < 2e-4
