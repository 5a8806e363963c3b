"""torchlibrosa.stft drop-in: Spectrogram / LogmelFilterBank (ctor sites
/root/reference/pytorch/models.py:166-173, calls :199-200).

The reference calls the two modules back to back; the 513-bin spectrogram between them is
2 MB per clip of pure intermediate traffic.  ``Spectrogram.forward`` therefore returns a
``LazyPowerSpectrogram`` -- a storage-less tensor subclass that remembers the waveform --
and ``LogmelFilterBank.forward`` recognises it and runs the fused waveform->log-mel kernel.
Any other use of the lazy tensor (any torch op at all) materialises the real power
spectrogram first, so the unfused seam stays correct.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.utils._pytree import tree_map

from sound_event_detection_dcase2017_task4_b200 import frontend as _fe


class LazyPowerSpectrogram(torch.Tensor):
    @staticmethod
    def __new__(cls, wave, hop):
        b, n = wave.shape
        shape = (b, 1, _fe.num_frames(n, hop), _fe.N_FFT // 2 + 1)
        r = torch.Tensor._make_wrapper_subclass(cls, shape, dtype=torch.float32, device=wave.device)
        r._wave, r._hop, r._dense = wave, hop, None
        return r

    def materialize(self):
        if self._dense is None:
            self._dense = _fe.stft_power(self._wave, self._hop)
        return self._dense

    def __repr__(self):
        return 'LazyPowerSpectrogram(shape=%s, device=%s)' % (tuple(self.shape), self.device)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        def unwrap(x):
            return x.materialize() if isinstance(x, LazyPowerSpectrogram) else x
        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))


class STFT(nn.Module):
    """Holder of the frozen DFT-matrix parameters (``conv_real.weight``, ``conv_imag.weight``):
    they exist for state_dict / parameters() / RNG-stream compatibility; the transform itself
    is the shared-memory real FFT in csrc/logmel.cu."""

    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
                 pad_mode='reflect', freeze_parameters=True):
        super().__init__()
        if n_fft != _fe.N_FFT or window != 'hann' or not center or pad_mode != 'reflect':
            raise NotImplementedError(
                'the B200 front-end implements the reference configuration only: n_fft=1024, '
                "window='hann', center=True, pad_mode='reflect' (utils/config.py:11-16)")
        win_length = n_fft if win_length is None else win_length
        if win_length != n_fft:
            raise NotImplementedError('win_length must equal n_fft (1024)')
        self.n_fft = n_fft
        self.hop_length = win_length // 4 if hop_length is None else hop_length
        self.center, self.pad_mode = center, pad_mode
        out_channels = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, out_channels, kernel_size=n_fft, stride=self.hop_length,
                                   padding=0, dilation=1, groups=1, bias=False)
        self.conv_imag = nn.Conv1d(1, out_channels, kernel_size=n_fft, stride=self.hop_length,
                                   padding=0, dilation=1, groups=1, bias=False)
        w_real, w_imag = _fe.dft_conv_weights(n_fft, win_length)
        self.conv_real.weight.data = torch.from_numpy(w_real)
        self.conv_imag.weight.data = torch.from_numpy(w_imag)
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False


class Spectrogram(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True,
                 pad_mode='reflect', power=2.0, freeze_parameters=True):
        super().__init__()
        if power != 2.0:
            raise NotImplementedError('only the power spectrogram (power=2.0) is implemented')
        self.power = power
        self.stft = STFT(n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                         center=center, pad_mode=pad_mode, freeze_parameters=True)

    def forward(self, input):
        """(B, L) -> (B, 1, T, 513) power spectrogram (lazy: see module docstring)."""
        if input.dtype not in (torch.float32, torch.int16):
            input = input.float()
        if not input.is_cuda:
            raise RuntimeError('Spectrogram: CUDA tensor required (no CPU path in this package)')
        return LazyPowerSpectrogram(input.detach().contiguous(), self.stft.hop_length)


class LogmelFilterBank(nn.Module):
    def __init__(self, sr=32000, n_fft=2048, n_mels=64, fmin=50, fmax=14000, is_log=True, ref=1.0,
                 amin=1e-10, top_db=80.0, freeze_parameters=True):
        super().__init__()
        self.is_log, self.ref, self.amin, self.top_db = is_log, ref, amin, top_db
        self.melW = nn.Parameter(torch.from_numpy(_fe.mel_weight_matrix(sr, n_fft, n_mels, fmin, fmax)))
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False
        self._banks = {}        # device -> MelBankCSR; the dict object is shared by DataParallel replicas

    def mel_bank(self):
        return _fe.mel_bank_for(self.melW, self._banks)

    def _apply(self, fn, *args, **kwargs):
        self._banks.clear()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._banks.clear()
        return super()._load_from_state_dict(*args, **kwargs)

    def forward(self, input):
        """(B, 1, T, 513) -> (B, 1, T, n_mels); fused with the STFT when fed by Spectrogram."""
        bank = self.mel_bank()
        if isinstance(input, LazyPowerSpectrogram) and input._dense is None and self.is_log:
            out = _fe.logmel(input._wave, input._hop, bank, amin=self.amin, ref=self.ref)
        else:
            if isinstance(input, LazyPowerSpectrogram):
                input = input.materialize()
            out = _fe.mel_db(input, bank, amin=self.amin, ref=self.ref, is_log=self.is_log)
        if self.is_log and self.top_db is not None:
            if self.top_db < 0:
                raise ValueError('top_db must be non-negative')
            out = torch.clamp(out, min=out.max().item() - self.top_db, max=np.inf)
        return out
