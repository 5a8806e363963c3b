"""-m gpu: csrc/vad.cu through the C ABI against the CPU oracle (oracle/vad.py, pinned to the unmodified
reference by tests/golden/vad_golden.*) and against the golden vectors themselves.  Bit-exact frame indices."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    with open(os.path.join(ROOT, 'tests', 'golden', 'vad_golden.json')) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(ROOT, 'tests', 'golden', 'vad_golden.npz'))


def test_series_golden_vectors():
    from sound_event_detection_dcase2017_task4_b200 import postproc
    meta, arrays = _golden()
    for i, case in enumerate(meta['series']):
        got = postproc.activity_detection(arrays['x_%d' % i], case['thres'], case['low'], case['n_smooth'],
                                          case['n_salt'])
        assert got == case['pairs'], (i, got, case['pairs'])


def test_event_list_golden_vectors():
    from sound_event_detection_dcase2017_task4_b200 import postproc
    meta, arrays = _golden()
    for t, case in enumerate(meta['events']):
        frame = np.repeat(arrays['frame8_%d' % t], 8, axis=1)
        frame[:, -1, :] = 0.0
        out = {'audio_name': case['names'], 'clipwise_output': arrays['clip_%d' % t], 'framewise_output': frame}
        ev = postproc.frame_prediction_to_event_prediction(out, dict(case['params']), meta['frames_per_second'],
                                                           meta['labels'])
        assert [[e['filename'], e['onset'], e['offset'], e['event_label']] for e in ev] == case['events']


@pytest.mark.parametrize('n,t,k', [(64, 1000, 17), (3, 1, 17), (5, 7, 3), (0, 1000, 17)])
def test_batch_against_oracle(n, t, k):
    """Seeded random tracks (smooth, spiky, x8-interpolated), per-class parameters, single and double threshold."""
    from oracle import vad as ovad
    from sound_event_detection_dcase2017_task4_b200 import postproc
    rs = np.random.RandomState(n * 1000 + t)
    frame = rs.rand(n, t, k).astype(np.float32)
    if n:
        frame[n // 2:] = np.repeat(rs.rand(n - n // 2, (t + 7) // 8, k), 8, axis=1)[:, :t].astype(np.float32)
        frame[:, -1, :] = 0.0                       # keep clear of the reference's IndexError shape
    clip = rs.rand(n, k).astype(np.float32)
    at, hi = rs.rand(k) * 0.5, 0.5 + rs.rand(k) * 0.4
    lo, ns, nt = rs.rand(k) * 0.4, rs.randint(0, 12, k), rs.randint(0, 12, k)
    for low in (lo, None):
        counts, pairs = postproc.activity_detection_batch(frame, clip, at, hi, low, ns, nt)
        want, want_counts = [], np.zeros((n, k), dtype=np.int32)
        for i in range(n):
            for c in range(k):
                if clip[i, c] > np.float32(at[c]):
                    p = ovad.activity_detection(frame[i, :, c], hi[c], None if low is None else low[c], int(ns[c]),
                                                int(nt[c]))
                    want_counts[i, c] = len(p)
                    want += p
        assert np.array_equal(counts, want_counts)
        assert np.array_equal(pairs, np.asarray(want, dtype=np.int32).reshape(-1, 2))


def test_reference_crash_shape_raises():
    from sound_event_detection_dcase2017_task4_b200 import postproc
    with pytest.raises(IndexError):
        postproc.activity_detection(np.array([0.9, 0.0, 0.9], dtype=np.float32), 0.5, 0.2, 1, 0)
    assert postproc.activity_detection(np.array([0.9, 0.0, 0.9], dtype=np.float32), 0.5, None, 0, 0) == [[0, 1]]


def test_full_size_properties():
    """10 000 clips x 1000 frames x 17 classes: pairs are ordered per series, inside [0, T], and re-running is
    idempotent (same bytes) -- size-independent checks at evaluation-set scale."""
    from sound_event_detection_dcase2017_task4_b200 import postproc
    g = torch.Generator(device='cuda').manual_seed(5)
    n, t, k = 10000, 1000, 17
    frame = torch.rand((n, t // 8, k), generator=g, device='cuda').repeat_interleave(8, dim=1).contiguous()
    frame[:, -1, :] = 0
    clip = frame.amax(dim=1)
    c1, p1 = postproc.activity_detection_batch(frame, clip, 0.3, 0.7, 0.2, 10, 10)
    c2, p2 = postproc.activity_detection_batch(frame, clip, 0.3, 0.7, 0.2, 10, 10)
    assert np.array_equal(c1, c2) and np.array_equal(p1, p2)
    assert p1.shape[0] == int(c1.sum()) and p1.shape[0] > 1000
    assert (p1[:, 0] >= 0).all() and (p1[:, 1] <= t).all() and (p1[:, 1] - p1[:, 0] > 10).all()
    series = np.repeat(np.arange(n * k), c1.reshape(-1))
    same = series[1:] == series[:-1]
    assert (p1[1:, 0][same] - p1[:-1, 1][same] > 10).all()          # merged gaps are > n_smooth apart
