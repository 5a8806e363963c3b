"""Host side of the log-mel front-end: frozen-parameter construction (state-dict contract of
torchlibrosa 0.0.4) and the launch wrappers over the C ABI.

Reference interfaces served: ``torchlibrosa.stft.Spectrogram`` / ``LogmelFilterBank`` as
constructed at /root/reference/pytorch/models.py:166-173 and called at :199-200.
"""
import math
import threading

import numpy as np
import torch

from . import _lib

N_FFT = 1024            # the only transform size the kernels implement (utils/config.py:11)


# ------------------------------------------------------------------ frozen parameters (host)
def dft_conv_weights(n_fft, win_length):
    """Values of ``stft.conv_real.weight`` / ``stft.conv_imag.weight``: rows k of the Hann-
    windowed DFT matrix, (n_fft//2+1, 1, n_fft) float32 each.  Kept only because they are part
    of the reference's state_dict / parameters() contract -- the kernels never read them."""
    taps = np.arange(n_fft, dtype=np.float64)
    window = np.zeros(n_fft, dtype=np.float64)
    lpad = (n_fft - win_length) // 2
    m = np.arange(win_length, dtype=np.float64)
    window[lpad:lpad + win_length] = 0.5 - 0.5 * np.cos(2.0 * np.pi * m / win_length)
    bins = np.arange(n_fft // 2 + 1)
    # exp(-2 pi i / n) ** (x*y), evaluated like the published builder (complex128 integer power)
    base = np.exp(-2.0j * np.pi / n_fft)
    mat = np.power(base, bins[:, None] * taps[None, :].astype(np.int64)) * window[None, :]
    return (np.ascontiguousarray(mat.real[:, None, :], dtype=np.float32),
            np.ascontiguousarray(mat.imag[:, None, :], dtype=np.float32))


def mel_weight_matrix(sr, n_fft, n_mels, fmin, fmax):
    """Value of ``logmel_extractor.melW``: (n_fft//2+1, n_mels) float32, Slaney scale, area norm
    (librosa ``filters.mel(htk=False, norm=1)``).  float64 ramps -> float32 store -> in-place
    float32 scale by the float64 area factor."""
    lin, knee, logstep = 200.0 / 3.0, 1000.0, math.log(6.4) / 27.0

    def to_mel(f):
        return f / lin if f < knee else knee / lin + math.log(f / knee) / logstep

    mels = np.linspace(to_mel(float(fmin)), to_mel(float(fmax)), n_mels + 2)
    hz = np.where(mels >= knee / lin, knee * np.exp(logstep * (mels - knee / lin)), mels * lin)
    fft_hz = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    ramps = hz[:, None] - fft_hz[None, :]
    widths = np.diff(hz)
    lower = -ramps[:-2] / widths[:-1, None]
    upper = ramps[2:] / widths[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper)).astype(np.float32)
    w *= (2.0 / (hz[2:n_mels + 2] - hz[:n_mels]))[:, None]
    return np.ascontiguousarray(w.T)


# ------------------------------------------------------------------ mel bank in CSR-by-mel form
class MelBankCSR(object):
    """Device-resident sparse view of a (n_bins, n_mels) melW: for each mel the contiguous run of
    FFT bins [lo, lo+cnt) covering its non-zero taps (zeros inside the run are kept, so any melW
    is represented exactly; the reference bank has 866 taps of 32 832)."""

    def __init__(self, melW):
        w = melW.detach().to('cpu', torch.float32).numpy()
        n_bins, n_mels = w.shape
        lo = np.zeros(n_mels, dtype=np.int32)
        off = np.zeros(n_mels + 1, dtype=np.int32)
        taps = []
        for m in range(n_mels):
            nz = np.nonzero(w[:, m])[0]
            if len(nz) == 0:
                off[m + 1] = off[m]
                continue
            lo[m] = nz[0]
            run = w[nz[0]:nz[-1] + 1, m]
            taps.append(run)
            off[m + 1] = off[m] + len(run)
        flat = np.concatenate(taps) if taps else np.zeros(1, dtype=np.float32)
        # limits of the fused kernel's shared-memory projection schedule (csrc/logmel.cu: kMaxMels, kMaxPieces,
        # pieces of <= 16 consecutive taps); the reference bank needs 64 filters / 90 pieces
        pieces = int(sum((int(off[m + 1] - off[m]) + 15) // 16 for m in range(n_mels)))
        if n_mels > 128 or pieces > 128:
            raise NotImplementedError('mel bank too large for the fused log-mel kernel: %d filters / %d 16-tap pieces '
                                      '(limits 128 / 128)' % (n_mels, pieces))
        dev = melW.device
        self.n_bins, self.n_mels = n_bins, n_mels
        self.w = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float32)).to(dev)
        self.lo = torch.from_numpy(lo).to(dev)
        self.off = torch.from_numpy(off).to(dev)


_csr_lock = threading.Lock()


def mel_bank_for(melW, shared=None):
    """Cached CSR bank for a melW parameter.

    First level: an attribute on the tensor object itself, valid while (version, device) match --
    the steady state of a single-device model (same Parameter object every step).
    Second level (``shared``: a dict owned by the LogmelFilterBank module and therefore shared by
    DataParallel's shallow-copied replicas, whose parameters are fresh tensors every forward):
    one bank per device for FROZEN parameters; the owning module clears it whenever its
    parameters are reloaded or moved (load_state_dict / .to())."""
    hit = getattr(melW, '_sed_bank', None)
    if hit is not None and hit[0] == melW._version and hit[1] == melW.device:
        return hit[2]
    bank = None
    use_shared = shared is not None and not melW.requires_grad
    if use_shared:
        with _csr_lock:
            bank = shared.get(melW.device)
    if bank is None:
        bank = MelBankCSR(melW)
        if use_shared:
            with _csr_lock:
                shared[melW.device] = bank
    try:
        melW._sed_bank = (melW._version, melW.device, bank)
    except Exception:
        pass
    return bank


# ------------------------------------------------------------------ launch wrappers
def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError('%s: expected a CUDA tensor -- this path has no CPU implementation '
                           '(the CPU oracle lives in oracle/ and is test infrastructure)' % what)


def num_frames(n_samples, hop):
    return n_samples // hop + 1


def logmel(wave, hop, bank, amin=1e-10, ref=1.0, out=None):
    """Fused Spectrogram+LogmelFilterBank: (B, L) float32 or int16 PCM -> (B, 1, T, n_mels) fp32."""
    _require_cuda(wave, 'logmel')
    assert wave.dim() == 2
    wave = wave.contiguous()
    b, n = wave.shape
    t = num_frames(n, hop)
    if out is None:
        out = torch.empty((b, 1, t, bank.n_mels), dtype=torch.float32, device=wave.device)
    db_offset = 10.0 * math.log10(max(amin, ref))
    name = {torch.float32: 'sed_logmel_f32', torch.int16: 'sed_logmel_i16'}.get(wave.dtype)
    if name is None:
        raise TypeError('logmel: waveform must be float32 or int16, got %s' % wave.dtype)
    with torch.cuda.device(wave.device):
        _lib.call(name, wave.data_ptr(), b, n, hop, bank.w.data_ptr(), bank.lo.data_ptr(),
                  bank.off.data_ptr(), bank.n_mels, amin, db_offset, out.data_ptr(),
                  _lib.stream_of(wave))
    return out


def stft_power(wave, hop):
    """Spectrogram.forward alone: (B, L) float32 -> (B, 1, T, 513) power spectrogram."""
    _require_cuda(wave, 'stft_power')
    wave = wave.contiguous().float()
    b, n = wave.shape
    out = torch.empty((b, 1, num_frames(n, hop), N_FFT // 2 + 1), dtype=torch.float32,
                      device=wave.device)
    with torch.cuda.device(wave.device):
        _lib.call('sed_stft_power_f32', wave.data_ptr(), b, n, hop, out.data_ptr(),
                  _lib.stream_of(wave))
    return out


def mel_db(power, bank, amin=1e-10, ref=1.0, is_log=True):
    """LogmelFilterBank.forward alone: (..., n_bins) fp32 -> (..., n_mels)."""
    _require_cuda(power, 'mel_db')
    power = power.contiguous().float()
    assert power.shape[-1] == bank.n_bins
    rows = power.numel() // bank.n_bins
    out = torch.empty(power.shape[:-1] + (bank.n_mels,), dtype=torch.float32, device=power.device)
    db_offset = 10.0 * math.log10(max(amin, ref))
    with torch.cuda.device(power.device):
        _lib.call('sed_mel_db_f32', power.data_ptr(), rows, bank.n_bins, bank.w.data_ptr(),
                  bank.lo.data_ptr(), bank.off.data_ptr(), bank.n_mels, amin, db_offset,
                  1 if is_log else 0, out.data_ptr(), _lib.stream_of(power))
    return out
