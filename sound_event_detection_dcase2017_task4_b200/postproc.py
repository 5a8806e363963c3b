"""Frame-wise probabilities -> sound events on the device: the drop-in for

    utilities.frame_prediction_to_event_prediction      /root/reference/utils/utilities.py:70-123
    vad.activity_detection                              /root/reference/utils/vad.py:11-41

Same arguments, same event list (filenames, onset / offset seconds = frame index / frames_per_second, labels, in
the reference's (audio, class, pair) order); the per-(clip, class) Python loops are one CUDA thread each
(csrc/vad.cu), bit-exact frame indices.  Thresholds are compared in fp32, which is what numpy >= 2 does for the
reference's ``float32 array > python float`` expressions.

There is no CPU path here: inputs given as numpy arrays are uploaded, a missing CUDA device is an error.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_of

FRAMES_PER_SECOND = 32000 // 320          # utils/config.py:13
LABELS = ['Train horn', 'Air horn, truck horn', 'Car alarm', 'Reversing beeps', 'Bicycle', 'Skateboard',
          'Ambulance (siren)', 'Fire engine, fire truck (siren)', 'Civil defense siren', 'Police car (siren)',
          'Screaming', 'Car', 'Car passing by', 'Bus', 'Truck', 'Motorcycle', 'Train']      # utils/config.py:29-34


def _device_tensor(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=device, dtype=dtype).contiguous()


def _per_class(v, k, dtype, device):
    if isinstance(v, (list, tuple, np.ndarray, torch.Tensor)):
        t = torch.as_tensor(np.asarray(v, dtype=np.float64))
        if t.numel() != k:
            raise ValueError('per-class parameter has %d entries, expected %d' % (t.numel(), k))
    else:
        t = torch.full((k,), float(v), dtype=torch.float64)
    return t.to(dtype).to(device)


def activity_detection_batch(framewise, clipwise, audio_tagging_threshold, sed_high_threshold, sed_low_threshold,
                             n_smooth, n_salt, device=None):
    """framewise (N, T, K), clipwise (N, K) or None -> (counts (N, K) int32 numpy, pairs (P, 2) int32 numpy).
    ``pairs`` holds every series' [bgn, fin] frame indices in (n, k, pair) order; ``counts`` says how many belong
    to each series.  Raises IndexError where the reference does (vad.py:78 reading x[T])."""
    if device is None:
        device = framewise.device if isinstance(framewise, torch.Tensor) and framewise.is_cuda else torch.device('cuda')
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('postproc: a CUDA device is required (no CPU path)')
    frame = _device_tensor(framewise, torch.float32, device)
    n, t, k = frame.shape
    clip = None if clipwise is None else _device_tensor(clipwise, torch.float32, device)
    at = None if clip is None else _per_class(audio_tagging_threshold, k, torch.float32, device)
    hi = _per_class(sed_high_threshold, k, torch.float32, device)
    lo = None if sed_low_threshold is None else _per_class(sed_low_threshold, k, torch.float32, device)
    ns = _per_class(n_smooth, k, torch.int32, device)
    nt = _per_class(n_salt, k, torch.int32, device)
    counts = torch.zeros((n, k), dtype=torch.int32, device=device)
    flags = torch.zeros((n, k), dtype=torch.int32, device=device)
    if n == 0:
        return counts.cpu().numpy(), np.zeros((0, 2), dtype=np.int32)
    with torch.cuda.device(device):
        s = stream_of(frame)
        call('sed_vad_count', frame.data_ptr(), ptr(clip), n, t, k, ptr(at), hi.data_ptr(), ptr(lo), ns.data_ptr(),
             nt.data_ptr(), counts.data_ptr(), flags.data_ptr(), s)
        csum = torch.cumsum(counts.view(-1).to(torch.int64), 0)
        offsets = (csum - counts.view(-1)).contiguous()
        total = int(csum[-1].item())
        if int(flags.max().item()) != 0:
            raise IndexError('index %d is out of bounds for axis 0 with size %d (a second active run begins on the '
                             'last frame: utils/vad.py:78 reads x[len(x)])' % (t, t))
        pairs = torch.empty((max(total, 1), 2), dtype=torch.int32, device=device)
        if total > 0:
            call('sed_vad_fill', frame.data_ptr(), ptr(clip), n, t, k, ptr(at), hi.data_ptr(), ptr(lo),
                 ns.data_ptr(), nt.data_ptr(), offsets.data_ptr(), pairs.data_ptr(), s)
    return counts.cpu().numpy(), pairs[:total].cpu().numpy()


def activity_detection(x, thres, low_thres=None, n_smooth=1, n_salt=0):
    """vad.activity_detection (vad.py:11-41) for one series: list of [bgn, fin]."""
    x = np.asarray(x, dtype=np.float32).reshape(1, -1, 1)
    _, pairs = activity_detection_batch(x, None, None, thres, low_thres, n_smooth, n_salt)
    return [[int(a), int(b)] for a, b in pairs]


def frame_prediction_to_event_prediction(output_dict, sed_params_dict, frames_per_second=FRAMES_PER_SECOND,
                                         labels=LABELS):
    """Same contract as utilities.frame_prediction_to_event_prediction (utilities.py:70-123)."""
    frame = output_dict['framewise_output']
    counts, pairs = activity_detection_batch(
        frame, output_dict['clipwise_output'], sed_params_dict['audio_tagging_threshold'],
        sed_params_dict['sed_high_threshold'], sed_params_dict['sed_low_threshold'], sed_params_dict['n_smooth'],
        sed_params_dict['n_salt'])
    n, k = counts.shape
    series = np.repeat(np.arange(n * k), counts.reshape(-1))
    names = output_dict['audio_name']
    fps = float(frames_per_second)
    return [{'filename': names[s // k], 'onset': int(b) / fps, 'offset': int(f) / fps, 'event_label': labels[s % k]}
            for s, (b, f) in zip(series, pairs)]
