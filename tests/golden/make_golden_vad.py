"""Mint golden vectors for the frame -> event post-processing from the UNMODIFIED reference
(/root/reference/utils/vad.py, utilities.frame_prediction_to_event_prediction).  Run in the authoring
container only; only tests/golden/vad_golden.{json,npz} travel.

    python tests/golden/make_golden_vad.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def series_cases(rs):
    """(x, thres, low, n_smooth, n_salt) cases: smooth probability tracks, spiky tracks, constant tracks, edges."""
    cases = []
    for i in range(60):
        T = int(rs.choice([1, 2, 3, 7, 50, 125, 1000]))
        kind = i % 5
        if kind == 0:       # smooth random walk through a sigmoid (what a trained model emits)
            x = 1.0 / (1.0 + np.exp(-np.cumsum(rs.randn(T)) * 0.7))
        elif kind == 1:     # independent uniform frames (maximally spiky)
            x = rs.rand(T)
        elif kind == 2:     # piecewise-constant blocks of 8 frames (the x8 interpolate output)
            x = np.repeat(rs.rand((T + 7) // 8), 8)[:T]
        elif kind == 3:     # mostly silent with a few bursts
            x = rs.rand(T) * 0.2
            for _ in range(3):
                a = rs.randint(0, T)
                x[a:a + rs.randint(1, 20)] = 0.9
        else:               # everything above / below
            x = np.full(T, 0.95 if i % 2 else 0.01)
        x = x.astype(np.float32)
        if T >= 2 and x[-1] > 0.5 and x[-2] <= 0.5 and (x[:-2] > 0.5).any():
            x[-1] = 0.0      # the reference raises IndexError on this shape (vad.py:78); covered separately
        thres = float(rs.choice([0.5, 0.3, 0.7, 0.9]))
        low = float(rs.choice([0.1, 0.2, 0.3, thres]))
        cases.append((x, thres, low, int(rs.choice([0, 1, 2, 10])), int(rs.choice([0, 1, 4, 10]))))
    return cases


def main():
    sys.path.insert(0, os.path.join(REF, 'utils'))
    for name in ('librosa', 'h5py', 'sed_eval', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    import vad                                     # the reference's own file
    import config
    import utilities
    rs = np.random.RandomState(20171004)
    gold = {'series': [], 'events': []}
    arrays = {}
    for x, thres, low, n_smooth, n_salt in series_cases(rs):
        try:
            pairs = vad.activity_detection(x, thres, low, n_smooth, n_salt)
            pairs = [[int(a), int(b)] for a, b in pairs]
        except IndexError:
            pairs = 'IndexError'
        arrays['x_%d' % len(gold['series'])] = x
        gold['series'].append({'thres': thres, 'low': low, 'n_smooth': n_smooth,
                               'n_salt': n_salt, 'pairs': pairs})
    # the known crash shape of the reference: a second run that begins on the last frame
    x = np.array([0.9, 0.0, 0.9], dtype=np.float32)
    try:
        vad.activity_detection(x, 0.5, 0.2, 1, 0)
        crash = False
    except IndexError:
        crash = True
    gold['last_frame_run_raises_index_error'] = crash
    # whole-function cases: (audios, 1000 frames, 17 classes), scalar and per-class parameters
    for trial in range(3):
        N, T, K = 6, 1000, config.classes_num
        frame = np.repeat(rs.rand(N, T // 8, K), 8, axis=1).astype(np.float32)      # interpolate-x8 structure
        frame[:, -1, :] = 0.0
        clip = frame.max(axis=1) * rs.rand(N, K).astype(np.float32)
        names = ['Y%05d.wav' % i for i in range(N)]
        if trial == 0:
            params = {'audio_tagging_threshold': 0.5, 'sed_high_threshold': 0.75, 'sed_low_threshold': 0.25,
                      'n_smooth': 10, 'n_salt': 10}
        else:
            params = {'audio_tagging_threshold': [float(v) for v in rs.rand(K) * 0.6],
                      'sed_high_threshold': [float(v) for v in 0.4 + rs.rand(K) * 0.5],
                      'sed_low_threshold': [float(v) for v in 0.05 + rs.rand(K) * 0.3],
                      'n_smooth': [int(v) for v in rs.randint(0, 20, K)],
                      'n_salt': [int(v) for v in rs.randint(0, 20, K)]}
        out = {'audio_name': names, 'clipwise_output': clip, 'framewise_output': frame}
        ev = utilities.frame_prediction_to_event_prediction(out, {k: (list(v) if isinstance(v, list) else v)
                                                                  for k, v in params.items()})
        arrays['frame8_%d' % trial] = frame[:, ::8, :]
        arrays['clip_%d' % trial] = clip
        gold['events'].append({'params': params, 'names': names,
                               'events': [[e['filename'], e['onset'], e['offset'], e['event_label']] for e in ev]})
    gold['frames_per_second'] = config.frames_per_second
    gold['labels'] = config.labels
    with open(os.path.join(HERE, 'vad_golden.json'), 'w') as f:
        json.dump(gold, f)
    np.savez_compressed(os.path.join(HERE, 'vad_golden.npz'), **arrays)
    print('series cases: %d (IndexError in %d), event cases: %s' % (
        len(gold['series']), sum(1 for s in gold['series'] if s['pairs'] == 'IndexError'),
        [len(e['events']) for e in gold['events']]))


if __name__ == '__main__':
    main()
