"""CPU: the clip store / sampler drop-in against index streams minted from the unmodified reference samplers
(tests/golden/make_golden_sampler.py) and against the reference's conversion rules."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def gold():
    with open(os.path.join(ROOT, 'tests', 'golden', 'sampler_golden.json')) as f:
        return json.load(f)


def _store(tmp_path, n, samples=64, classes=17, frames=None, seed=0):
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    rs = np.random.RandomState(seed)
    wave = rs.randint(-32768, 32768, size=(n, samples)).astype(np.int16)
    target = (rs.rand(n, classes) < 0.2)
    strong = (rs.rand(n, frames, classes) < 0.1) if frames else None
    names = ['Y%06d.wav' % i for i in range(n)]
    path = str(tmp_path / ('store_%d' % n))
    dg.ClipStore.write(path, names, wave, target, strong)
    return path, names, wave, target, strong


def test_train_sampler_index_stream_is_the_reference_stream(gold, tmp_path):
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    for case in gold['train']:
        if case['audios_num'] > 1000:
            path, _, _, _, _ = _store(tmp_path, case['audios_num'], samples=1)
        else:
            path, _, _, _, _ = _store(tmp_path, case['audios_num'])
        it = iter(dg.TrainSampler(path, case['batch_size']))
        for want in case['batches']:
            metas = next(it)
            assert [int(m['index_in_hdf5']) for m in metas] == want
            assert all(m['hdf5_path'] == path for m in metas)


def test_test_sampler_is_sequential_with_ragged_tail(gold, tmp_path):
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    for case in gold['test']:
        path, _, _, _, _ = _store(tmp_path, case['audios_num'], samples=2)
        got = [[int(m['index_in_hdf5']) for m in b] for b in dg.TestSampler(path, case['batch_size'])]
        assert got == case['batches']


def test_dataset_item_and_collate_follow_the_reference(tmp_path):
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    path, names, wave, target, strong = _store(tmp_path, 9, frames=11)
    ds = dg.DCASE2017Task4Dataset()
    items = [ds[{'hdf5_path': path, 'index_in_hdf5': i}] for i in (3, 0, 8)]
    for i, d in zip((3, 0, 8), items):
        assert d['audio_name'] == names[i]
        assert d['waveform'].dtype == np.float32 and np.array_equal(d['waveform'], (wave[i] / 32767.).astype(np.float32))
        assert d['target'].dtype == np.float32 and np.array_equal(d['target'], target[i].astype(np.float32))
        assert np.array_equal(d['strong_target'], strong[i].astype(np.float32))
    batch = dg.collate_fn(items)
    assert batch['waveform'].shape == (3, 64) and batch['target'].shape == (3, 17)
    assert batch['strong_target'].shape == (3, 11, 17) and list(batch['audio_name']) == [names[i] for i in (3, 0, 8)]


def test_batcher_matches_sampler_plus_dataset(tmp_path):
    """The int16 fast path delivers the same clips, in the same order, as DataLoader(dataset, TrainSampler, collate_fn)."""
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    path, names, wave, target, _ = _store(tmp_path, 23)
    ds = dg.DCASE2017Task4Dataset()
    slow = iter(dg.TrainSampler(path, 8))
    fast = iter(dg.TrainBatcher(path, 8))
    for _ in range(9):                                   # crosses two re-shuffles
        want = dg.collate_fn([ds[m] for m in next(slow)])
        got = next(fast)
        assert list(got['audio_name']) == list(want['audio_name'])
        assert got['waveform'].dtype == np.int16
        assert np.array_equal((got['waveform'] / 32767.).astype(np.float32), want['waveform'])
        assert np.array_equal(got['target'], want['target'])


def test_clip_store_gather_threaded_equals_rows(tmp_path):
    """ClipStore.gather copies rows with positional reads on a thread pool, in request order, repeats included, into
    caller-provided buffers or fresh arrays; targets come back as fp32."""
    import numpy as np
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    rs = np.random.RandomState(3)
    n, samples = 41, 120000                                   # 41 x 120000 int16 > 4 M elements: the threaded path
    wave = rs.randint(-32768, 32767, size=(n, samples)).astype(np.int16)
    target = (rs.rand(n, 17) > 0.7).astype(np.uint8)
    strong = (rs.rand(n, 11, 17) > 0.7).astype(np.uint8)
    store = dg.ClipStore.write(str(tmp_path / 'store'), ['a%d' % i for i in range(n)], wave, target, strong)
    idx = rs.randint(0, n, size=37)
    idx[5] = idx[6]                                           # a repeated clip
    for threads in (1, 4):
        got = store.gather(idx, threads=threads)
        assert got['waveform'].dtype == np.int16 and np.array_equal(got['waveform'], wave[idx])
        assert got['target'].dtype == np.float32 and np.array_equal(got['target'], target[idx].astype(np.float32))
        assert np.array_equal(got['strong_target'], strong[idx].astype(np.float32))
        assert list(got['audio_name']) == ['a%d' % i for i in idx]
        out = {'waveform': np.full((37, samples), 7, dtype=np.int16), 'target': np.zeros((37, 17), dtype=np.float32)}
        got = store.gather(idx, out=out, threads=threads)
        assert got['waveform'] is out['waveform'] and np.array_equal(out['waveform'], wave[idx])
        assert np.array_equal(out['target'], target[idx].astype(np.float32))
