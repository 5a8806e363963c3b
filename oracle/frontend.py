"""CPU restatement of the torchlibrosa==0.0.4 front-end the reference imports.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the package is pinned at
/root/reference/requirements.txt:2 but is not vendored, installed or downloadable, so this
file restates its published behaviour and is anchored on the
reference's call sites:

* ctor  /root/reference/pytorch/models.py:166-168  Spectrogram(n_fft, hop_length, win_length,
        window, center, pad_mode, freeze_parameters)
* ctor  /root/reference/pytorch/models.py:171-173  LogmelFilterBank(sr, n_fft, n_mels, fmin,
        fmax, ref, amin, top_db, freeze_parameters)
* ctor  /root/reference/pytorch/models.py:176-177  SpecAugmentation(time_drop_width,
        time_stripes_num, freq_drop_width, freq_stripes_num)
* calls /root/reference/pytorch/models.py:199-200, :206-207 (and the six sibling classes)

The arithmetic is the reference's own: a windowed DFT evaluated as two fp32 ``conv1d`` with
frozen DFT-matrix weights, ``real**2 + imag**2``, a dense fp32 matmul with the Slaney mel
bank and ``10*log10(clamp(., amin))``.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- mel bank
def _hz_to_slaney_mel(hz):
    """Slaney (auditory toolbox) mel scale: linear below 1 kHz, log above."""
    hz = np.asarray(hz, dtype=np.float64)
    lin_step = 200.0 / 3.0
    knee_hz = 1000.0
    knee_mel = knee_hz / lin_step
    log_step = math.log(6.4) / 27.0
    mel = hz / lin_step
    above = hz >= knee_hz
    mel = np.where(above, knee_mel + np.log(np.maximum(hz, 1e-30) / knee_hz) / log_step, mel)
    return mel


def _slaney_mel_to_hz(mel):
    mel = np.asarray(mel, dtype=np.float64)
    lin_step = 200.0 / 3.0
    knee_hz = 1000.0
    knee_mel = knee_hz / lin_step
    log_step = math.log(6.4) / 27.0
    hz = mel * lin_step
    above = mel >= knee_mel
    hz = np.where(above, knee_hz * np.exp(log_step * (mel - knee_mel)), hz)
    return hz


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """(n_mels, n_fft//2+1) float32 triangular bank, area-normalised (librosa ``filters.mel``
    defaults ``htk=False, norm=1`` of the 0.7 era the reference used).

    The triangles are evaluated in float64, stored to float32, and only then scaled by the
    float64 area norm (in-place float32 multiply) -- that order is part of the published
    behaviour and decides the last bit of every tap.
    """
    n_bins = n_fft // 2 + 1
    bin_hz = np.linspace(0.0, sr / 2.0, n_bins)
    edges_mel = np.linspace(_hz_to_slaney_mel(fmin), _hz_to_slaney_mel(fmax), n_mels + 2)
    edges_hz = _slaney_mel_to_hz(edges_mel)
    width = np.diff(edges_hz)
    dist = edges_hz[:, None] - bin_hz[None, :]            # (n_mels+2, n_bins)
    bank = np.zeros((n_mels, n_bins), dtype=np.float32)
    for m in range(n_mels):
        rising = -dist[m] / width[m]
        falling = dist[m + 2] / width[m + 1]
        bank[m] = np.maximum(0.0, np.minimum(rising, falling))
    area = 2.0 / (edges_hz[2:n_mels + 2] - edges_hz[:n_mels])
    bank *= area[:, None]
    return bank


# ----------------------------------------------------------------------------- STFT
def hann_dft_conv_weights(n_fft, win_length):
    """Frozen conv1d weights of the windowed DFT: two (n_fft//2+1, 1, n_fft) float32 arrays.

    window = periodic Hann of ``win_length`` centre-padded to ``n_fft``; DFT matrix built in
    complex128 as integer powers of exp(-2*pi*i/n_fft); the imaginary weights carry the
    negative sine.
    """
    n = np.arange(win_length, dtype=np.float64)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)       # scipy get_window('hann', fftbins=True)
    if win_length < n_fft:
        lpad = (n_fft - win_length) // 2
        win = np.pad(win, (lpad, n_fft - win_length - lpad))
    n_out = n_fft // 2 + 1
    idx = np.arange(n_fft)
    omega = np.exp(-2.0j * np.pi / n_fft)
    dft = np.power(omega, np.outer(idx, idx[:n_out]))            # (n_fft, n_out) complex128
    dft = dft * win[:, None]
    w_real = np.ascontiguousarray(np.real(dft).T[:, None, :]).astype(np.float32)
    w_imag = np.ascontiguousarray(np.imag(dft).T[:, None, :]).astype(np.float32)
    return w_real, w_imag


class STFT(nn.Module):
    """Windowed DFT as two frozen Conv1d(1, n_fft//2+1, n_fft, stride=hop).  Registers
    ``conv_real.weight`` / ``conv_imag.weight`` exactly like upstream so that state-dict
    keys and torch RNG consumption under ``torch.manual_seed`` match.
    """

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window='hann',
                 center=True, pad_mode='reflect', freeze_parameters=True):
        super().__init__()
        assert pad_mode in ('constant', 'reflect')
        assert window == 'hann', 'only the window the reference uses is restated'
        self.n_fft = n_fft
        self.center = center
        self.pad_mode = pad_mode
        win_length = n_fft if win_length is None else win_length
        hop_length = win_length // 4 if hop_length is None else hop_length
        self.hop_length = hop_length
        n_out = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, n_out, kernel_size=n_fft, stride=hop_length, padding=0,
                                   dilation=1, groups=1, bias=False)
        self.conv_imag = nn.Conv1d(1, n_out, kernel_size=n_fft, stride=hop_length, padding=0,
                                   dilation=1, groups=1, bias=False)
        w_real, w_imag = hann_dft_conv_weights(n_fft, win_length)
        self.conv_real.weight.data = torch.from_numpy(w_real)
        self.conv_imag.weight.data = torch.from_numpy(w_imag)
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, input):
        x = input[:, None, :]
        if self.center:
            x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode=self.pad_mode)
        real = self.conv_real(x)[:, None, :, :].transpose(2, 3)
        imag = self.conv_imag(x)[:, None, :, :].transpose(2, 3)
        return real, imag


class Spectrogram(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window='hann',
                 center=True, pad_mode='reflect', power=2.0, freeze_parameters=True):
        super().__init__()
        self.power = power
        self.stft = STFT(n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                         window=window, center=center, pad_mode=pad_mode,
                         freeze_parameters=True)

    def forward(self, input):
        real, imag = self.stft(input)
        spec = real ** 2 + imag ** 2
        if self.power != 2.0:
            spec = spec ** (self.power / 2.0)
        return spec


class LogmelFilterBank(nn.Module):
    def __init__(self, sr=32000, n_fft=2048, n_mels=64, fmin=50, fmax=14000, is_log=True,
                 ref=1.0, amin=1e-10, top_db=80.0, freeze_parameters=True):
        super().__init__()
        self.is_log = is_log
        self.ref = ref
        self.amin = amin
        self.top_db = top_db
        bank = slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax).T      # (n_bins, n_mels)
        self.melW = nn.Parameter(torch.from_numpy(np.ascontiguousarray(bank)))
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, input):
        mel = torch.matmul(input, self.melW)
        return self.power_to_db(mel) if self.is_log else mel

    def power_to_db(self, x):
        db = 10.0 * torch.log10(torch.clamp(x, min=self.amin, max=np.inf))
        db = db - 10.0 * np.log10(np.maximum(self.amin, self.ref))
        if self.top_db is not None:
            if self.top_db < 0:
                raise ValueError('top_db must be non-negative')
            db = torch.clamp(db, min=db.max().item() - self.top_db, max=np.inf)
        return db


# ----------------------------------------------------------------------------- SpecAugment
def draw_stripes(count, total_width, drop_width, stripes_num):
    """Replay of the upstream RNG protocol on the torch CPU default generator:
    for each of ``count`` samples, ``stripes_num`` times: width = randint(0, drop_width),
    begin = randint(0, total_width - width).  Returns int64 (count, stripes_num, 2) =
    (begin, width).  Index-valued => bit-exact contract."""
    out = torch.zeros((count, stripes_num, 2), dtype=torch.int64)
    for n in range(count):
        for s in range(stripes_num):
            width = torch.randint(low=0, high=drop_width, size=(1,))[0]
            begin = torch.randint(low=0, high=total_width - width, size=(1,))[0]
            out[n, s, 0] = begin
            out[n, s, 1] = width
    return out


class DropStripes(nn.Module):
    def __init__(self, dim, drop_width, stripes_num):
        super().__init__()
        assert dim in (2, 3)
        self.dim = dim
        self.drop_width = drop_width
        self.stripes_num = stripes_num

    def forward(self, input):
        assert input.ndimension() == 4
        if not self.training:
            return input
        stripes = draw_stripes(input.shape[0], input.shape[self.dim], self.drop_width,
                               self.stripes_num)
        for n in range(input.shape[0]):
            for s in range(self.stripes_num):
                b, w = int(stripes[n, s, 0]), int(stripes[n, s, 1])
                if self.dim == 2:
                    input[n, :, b:b + w, :] = 0
                else:
                    input[n, :, :, b:b + w] = 0
        return input


class SpecAugmentation(nn.Module):
    def __init__(self, time_drop_width, time_stripes_num, freq_drop_width, freq_stripes_num):
        super().__init__()
        self.time_dropper = DropStripes(dim=2, drop_width=time_drop_width,
                                        stripes_num=time_stripes_num)
        self.freq_dropper = DropStripes(dim=3, drop_width=freq_drop_width,
                                        stripes_num=freq_stripes_num)

    def forward(self, input):
        return self.freq_dropper(self.time_dropper(input))


# ----------------------------------------------------------------------------- independent check
def logmel_float64_reference(wave, sr=32000, n_fft=1024, hop=320, n_mels=64, fmin=50, fmax=14000,
                             amin=1e-10):
    """Independent float64 path (torch.stft + the mel bank in float64) used by the tests to
    bound the fp32 oracle's own rounding noise.  wave: (B, L) tensor."""
    w64 = wave.double()
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
    spec = torch.stft(w64, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=win, center=True,
                      pad_mode='reflect', return_complex=True)
    power = (spec.real ** 2 + spec.imag ** 2).transpose(1, 2)            # (B, T, bins)
    bank = torch.from_numpy(slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)).double()
    mel = power @ bank.T
    return 10.0 * torch.log10(torch.clamp(mel, min=amin))
