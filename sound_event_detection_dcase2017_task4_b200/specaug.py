"""Host-side replay of torchlibrosa's SpecAugmentation RNG protocol.

torchlibrosa 0.0.4 ``DropStripes`` (ctor /root/reference/pytorch/models.py:176-177, call
:206-207) draws, per sample and stripe, ``width = torch.randint(0, drop_width, (1,))`` then
``begin = torch.randint(0, total_width - width, (1,))`` from the torch *CPU default generator*:
8*B2 tiny host calls + 4*B2 slice-fill launches per step.  Each such randint consumes exactly
one 32-bit MT19937 output ``r`` and returns ``r % range``.  We therefore read the generator
state, produce the same outputs in bulk with numpy's MT19937 (identical engine), compute every
(begin, width) pair vectorised, and write the advanced state back -- the global torch RNG stream
stays bit-identical to the reference's, at ~100x less host time.  The index table is then applied
on the device inside the fused bn0 kernel.

A one-time self-check compares against real ``torch.randint`` calls; on mismatch (unknown
generator layout in another torch build) we fall back to calling ``torch.randint`` per draw --
still exact, only slower.
"""
import threading

import numpy as np
import torch

_MT_N = 624
_OFF_LEFT, _OFF_NEXT, _OFF_KEY = 8, 16, 24
_STATE_BYTES = 5056

_lock = threading.Lock()
_fast_ok = None


def _raw_draws(n):
    """n consecutive uint32 outputs of the torch default CPU generator (advances it)."""
    st = torch.get_rng_state()
    buf = st.numpy().copy()
    if buf.shape[0] != _STATE_BYTES:
        raise RuntimeError('unexpected CPU generator state size %d' % buf.shape[0])
    left = int(buf[_OFF_LEFT:_OFF_LEFT + 4].view(np.int32)[0])
    nxt = int(buf[_OFF_NEXT:_OFF_NEXT + 8].view(np.uint64)[0])
    key = buf[_OFF_KEY:_OFF_KEY + _MT_N * 8].view(np.uint64).astype(np.uint32)
    bg = np.random.MT19937()
    bg.state = {'bit_generator': 'MT19937', 'state': {'key': key, 'pos': _MT_N if left == 1 else nxt}}
    raw = bg.random_raw(n)
    new = bg.state['state']
    pos = int(new['pos'])
    buf[_OFF_KEY:_OFF_KEY + _MT_N * 8] = new['key'].astype(np.uint64).view(np.uint8)
    buf[_OFF_NEXT:_OFF_NEXT + 8] = np.array([pos], dtype=np.uint64).view(np.uint8)
    buf[_OFF_LEFT:_OFF_LEFT + 4] = np.array([_MT_N + 1 - pos], dtype=np.int32).view(np.uint8)
    torch.set_rng_state(torch.from_numpy(buf))
    return raw.astype(np.int64)


def _slow(count, total_width, drop_width, stripes_num):
    out = np.zeros((count, stripes_num, 2), dtype=np.int32)
    for n in range(count):
        for s in range(stripes_num):
            width = int(torch.randint(low=0, high=drop_width, size=(1,))[0])
            begin = int(torch.randint(low=0, high=total_width - width, size=(1,))[0])
            out[n, s] = (begin, width)
    return out


def _fast(count, total_width, drop_width, stripes_num):
    raw = _raw_draws(2 * count * stripes_num).reshape(count, stripes_num, 2)
    width = raw[:, :, 0] % drop_width
    begin = raw[:, :, 1] % (total_width - width)
    return np.stack([begin, width], axis=2).astype(np.int32)


def _self_check():
    saved = torch.get_rng_state()
    try:
        ok = True
        for seed in (1, 12345):
            torch.manual_seed(seed)
            torch.randint(0, 10, (3,))                   # move off the freshly-seeded state
            mid = torch.get_rng_state()
            a = _slow(5, 1001, 64, 2)
            after_slow = torch.get_rng_state()
            torch.set_rng_state(mid)
            b = _fast(5, 1001, 64, 2)
            ok = ok and np.array_equal(a, b) and torch.equal(after_slow, torch.get_rng_state())
        return ok
    except Exception:
        return False
    finally:
        torch.set_rng_state(saved)


def draw_stripes(count, total_width, drop_width, stripes_num):
    """int32 (count, stripes_num, 2) = (begin, width), consuming the torch CPU generator exactly
    like ``for n in range(count): for s in range(stripes_num): randint; randint``."""
    global _fast_ok
    if drop_width <= 0 or total_width - (drop_width - 1) <= 0:
        raise ValueError('invalid stripe configuration')
    with _lock:
        if _fast_ok is None:
            _fast_ok = _self_check()
        if _fast_ok:
            return _fast(count, total_width, drop_width, stripes_num)
        return _slow(count, total_width, drop_width, stripes_num)


def draw_spec_augment(count, n_frames, n_mels, time_drop_width=64, time_stripes_num=2, freq_drop_width=8,
                      freq_stripes_num=2):
    """Both tables in the reference's order: all time stripes (sample by sample), then all
    frequency stripes."""
    t = draw_stripes(count, n_frames, time_drop_width, time_stripes_num)
    f = draw_stripes(count, n_mels, freq_drop_width, freq_stripes_num)
    return t, f
