"""CPU: oracle/vad.py against the golden vectors minted from the unmodified reference
(tests/golden/make_golden_vad.py ran /root/reference/utils/vad.py and
utilities.frame_prediction_to_event_prediction).  Frame indices must be bit-exact."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def vad_golden():
    with open(os.path.join(ROOT, 'tests', 'golden', 'vad_golden.json')) as f:
        meta = json.load(f)
    arrays = np.load(os.path.join(ROOT, 'tests', 'golden', 'vad_golden.npz'))
    return meta, arrays


def test_activity_detection_matches_reference(vad_golden):
    from oracle import vad
    meta, arrays = vad_golden
    assert len(meta['series']) == 60
    nonempty = 0
    for i, case in enumerate(meta['series']):
        got = vad.activity_detection(arrays['x_%d' % i], case['thres'], case['low'], case['n_smooth'], case['n_salt'])
        assert got == case['pairs'], (i, got, case['pairs'])
        nonempty += bool(got)
    assert nonempty >= 20


def test_reference_crash_shape_is_reproduced(vad_golden):
    from oracle import vad
    meta, _ = vad_golden
    assert meta['last_frame_run_raises_index_error'] is True
    with pytest.raises(IndexError):
        vad.activity_detection(np.array([0.9, 0.0, 0.9], dtype=np.float32), 0.5, 0.2, 1, 0)


def test_event_prediction_matches_reference(vad_golden):
    from oracle import vad
    meta, arrays = vad_golden
    for t, case in enumerate(meta['events']):
        frame = np.repeat(arrays['frame8_%d' % t], 8, axis=1)
        frame[:, -1, :] = 0.0
        out = {'audio_name': case['names'], 'clipwise_output': arrays['clip_%d' % t], 'framewise_output': frame}
        ev = vad.frame_prediction_to_event_prediction(out, dict(case['params']), meta['frames_per_second'],
                                                      meta['labels'])
        got = [[e['filename'], e['onset'], e['offset'], e['event_label']] for e in ev]
        assert got == case['events']
        assert len(got) > 100
