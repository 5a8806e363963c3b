"""pytorch_utils drop-in (/root/reference/pytorch/pytorch_utils.py): move_data_to_device,
append_to_dict, forward (inference loop), do_mixup."""
import numpy as np
import torch


def move_data_to_device(x, device):
    if 'float' in str(x.dtype):
        x = torch.Tensor(x)
    elif 'int' in str(x.dtype):
        x = torch.LongTensor(x)
    else:
        return x
    return x.to(device)


def append_to_dict(dict, key, value):
    dict.setdefault(key, []).append(value)


def do_mixup(x, mixup_lambda):
    """out[i] = x[2i] * lam[2i] + x[2i+1] * lam[2i+1] (pytorch_utils.py:80-93).  Inside the models the
    feature mixup is fused into the bn0 kernel; this host-visible version serves the *targets*
    (main.py:246): a (2B, 17) tensor, i.e. plumbing-sized."""
    lam = mixup_lambda.to(x.dtype)
    shape = (-1,) + (1,) * (x.dim() - 1)
    return x[0::2] * lam[0::2].reshape(shape) + x[1::2] * lam[1::2].reshape(shape)


class _OutputDrain(object):
    """Device -> host drain of the per-batch outputs of ``forward``: two sets of pinned host buffers and a copy
    stream, so the (B, 1000, 17) framewise block of batch n leaves the device while batch n+1 computes (the reference
    blocks the compute stream twice per batch with ``.data.cpu().numpy()``, pytorch_utils.py:56-61)."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]          # {'key': pinned tensor}
        self.pending = [None, None]        # (event, {key: (pinned tensor, rows)})

    def submit(self, slot, outputs):
        """Start copying a dict of device tensors into pinned set ``slot`` (after the kernels that produce them)."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        bufs = self.slots[slot]
        if bufs is None or any(k not in bufs or bufs[k].shape[1:] != v.shape[1:] or bufs[k].shape[0] < v.shape[0]
                               or bufs[k].dtype != v.dtype for k, v in outputs.items()):
            bufs = self.slots[slot] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in outputs.items()}
        taken = {}
        with torch.cuda.stream(self.stream):
            for k, v in outputs.items():
                v = v.detach()
                bufs[k][:v.shape[0]].copy_(v, non_blocking=True)
                taken[k] = (bufs[k], v.shape[0], v)         # v stays referenced until its copy has finished
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending[slot] = (ev, taken)

    def collect(self, slot):
        """Host arrays of the batch last submitted into ``slot`` (waits for its copy only)."""
        if self.pending[slot] is None:
            return None
        ev, taken = self.pending[slot]
        ev.synchronize()
        self.pending[slot] = None
        return {k: buf.numpy()[:rows].copy() for k, (buf, rows, _) in taken.items()}


class _InputFeed(object):
    """Host -> device feed of the waveform batches of ``forward``: two pinned staging buffers and a copy stream.  The
    reference hands pageable numpy batches to ``move_data_to_device`` -- a synchronous copy on the compute stream that
    waits for the previous batch's kernels and blocks the host (measured: the loop ran at 8.5 k clips/s whatever the
    batch size).  Here batch n+1 is staged and copied while batch n computes."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.pinned = [None, None]
        self.dev = [None, None]
        self.copied = [None, None]         # H2D of the slot finished (recorded on the copy stream)
        self.consumed = [None, None]       # kernels that read the slot finished (recorded on the compute stream)

    def push(self, slot, wave):
        if isinstance(wave, np.ndarray):
            if wave.dtype not in (np.float32, np.int16):
                wave = wave.astype(np.float32) if 'float' in str(wave.dtype) else wave.astype(np.int16)
            wave = torch.from_numpy(np.ascontiguousarray(wave))
        elif wave.dtype not in (torch.float32, torch.int16):
            wave = wave.float()
        if wave.is_cuda:
            return wave.to(self.device)
        n = wave.shape[0]
        src = wave
        if not wave.is_pinned():
            buf = self.pinned[slot]
            if buf is None or buf.dtype != wave.dtype or buf.shape[1:] != wave.shape[1:] or buf.shape[0] < n:
                buf = self.pinned[slot] = torch.empty(wave.shape, dtype=wave.dtype).pin_memory()
            if self.copied[slot] is not None:
                self.copied[slot].synchronize()            # the previous copy out of this staging buffer is done
            buf[:n].copy_(wave)
            src = buf[:n]
        dst = self.dev[slot]
        if dst is None or dst.dtype != wave.dtype or dst.shape[1:] != wave.shape[1:] or dst.shape[0] < n:
            dst = self.dev[slot] = torch.empty(wave.shape, dtype=wave.dtype, device=self.device)
        if self.consumed[slot] is not None:
            self.stream.wait_event(self.consumed[slot])
        with torch.cuda.stream(self.stream):
            dst[:n].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.copied[slot] = ev
        torch.cuda.current_stream(self.device).wait_event(ev)
        return dst[:n]

    def done(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.consumed[slot] = ev


_PIPES = {}


def _pipes(device):
    """The (drain, feed) pair of ``device``, kept across ``forward`` calls: page-locking the 164 MB staging buffers of
    a 256-clip batch costs more than the batch's forward pass, and main.py evaluates every 2000 iterations."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _PIPES:
        dev = torch.device('cuda', key)
        _PIPES[key] = (_OutputDrain(dev), _InputFeed(dev))
    return _PIPES[key]


def pad_framewise_output(framewise_output, frames_num):
    """(N, T, K) numpy -> (N, frames_num, K): crop, or repeat the last frame.  ``interpolate`` (models.py:58-69)
    yields 8 * 125 = 1000 frames while the packed ``strong_target`` has 1001 (utils/config.py:14), which trips the
    shape assertion at pytorch/evaluate.py:19."""
    n, t, k = framewise_output.shape
    if frames_num == t:
        return framewise_output
    if frames_num < t:
        return np.ascontiguousarray(framewise_output[:, :frames_num])
    out = np.empty((n, frames_num, k), dtype=framewise_output.dtype)
    out[:, :t] = framewise_output
    out[:, t:] = framewise_output[:, t - 1:t]
    return out


def forward(model, data_loader, return_input=False, return_target=False, frames_num=None):
    """Inference loop of pytorch_utils.py:25-77: eval mode, no grad, outputs gathered on the host as numpy arrays
    with the reference's keys and shapes.  int16 PCM waveforms are accepted as they are (x/32767 happens on the
    device).  ``frames_num``: None keeps the model's own frame count (reference behaviour); an int, or 'target' for
    the frame count of the batches' ``strong_target``, pads/crops ``framewise_output`` (pad_framewise_output).

    Pipelining: batch n+1 is staged in pinned memory and copied on a copy stream while batch n computes (_InputFeed);
    batch n's outputs are drained to pinned host memory on another copy stream while batch n+1 runs; the host only
    waits for a batch's output copy when the NEXT batch has been enqueued."""
    device = next(model.parameters()).device
    if device.type != 'cuda':
        raise RuntimeError('forward: the model must live on a CUDA device (no CPU path in this package)')
    output_dict = {}
    drain, feed = _pipes(device)
    drain.pending = [None, None]                            # a call that raised mid-loop leaves nothing behind

    def harvest(slot):
        got = drain.collect(slot)
        if got is None:
            return
        append_to_dict(output_dict, 'clipwise_output', got['clipwise_output'])
        if 'framewise_output' in got:
            frame = got['framewise_output']
            want = pending_frames[slot]
            if want is not None:
                frame = pad_framewise_output(frame, want)
            append_to_dict(output_dict, 'framewise_output', frame)

    pending_frames = [None, None]
    n = -1
    for n, batch_data_dict in enumerate(data_loader):
        slot = n & 1
        batch_waveform = feed.push(slot, batch_data_dict['waveform'])
        with torch.no_grad():
            model.eval()
            batch_output = model(batch_waveform)
        feed.done(slot)
        outs = {'clipwise_output': batch_output['clipwise_output']}
        if 'framewise_output' in batch_output.keys():
            outs['framewise_output'] = batch_output['framewise_output']
        drain.submit(slot, outs)
        want = frames_num
        if want == 'target':
            want = batch_data_dict['strong_target'].shape[1] if 'strong_target' in batch_data_dict.keys() else None
        pending_frames[slot] = want
        harvest(slot ^ 1)                                   # the previous batch: its copy overlapped this launch
        append_to_dict(output_dict, 'audio_name', batch_data_dict['audio_name'])
        if return_input:
            append_to_dict(output_dict, 'waveform', batch_data_dict['waveform'])
        if return_target:
            for key in ('target', 'strong_target'):
                if key in batch_data_dict.keys():
                    append_to_dict(output_dict, key, batch_data_dict[key])
    if n >= 0:
        harvest(n & 1)
    for key in output_dict.keys():
        output_dict[key] = np.concatenate(output_dict[key], axis=0)
    return output_dict
