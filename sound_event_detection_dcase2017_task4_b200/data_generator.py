"""data_generator drop-in (/root/reference/utils/data_generator.py) over a memory-mapped int16 clip store.

The reference opens the HDF5 file once per ITEM in 8 DataLoader workers (data_generator.py:37), converts every clip
to fp32 on the host and collates a Python list of dicts.  At the B200 step rate the loader would have to deliver
~13 k clips/s x 640 KB; here the packed arrays are plain ``np.memmap`` files (the same rows the HDF5 datasets hold:
``audio_name``, int16 ``waveform``, ``target``, optional ``strong_target``; schema utils/features.py:232-260), a
whole batch is gathered with one fancy-index per key straight into a pinned buffer, and the int16 PCM goes to the
device as is (x/32767 is fused into the log-mel kernel).

Same names and semantics as the reference: ``DCASE2017Task4Dataset.__getitem__(meta)``, ``TrainSampler`` (infinite,
``RandomState(1234)`` shuffles, including its double indirection ``audio_indexes[audio_indexes[pointer]]`` and the
re-shuffle that happens BEFORE the wrapped item is appended), ``TestSampler`` (sequential, ragged last batch),
``collate_fn``.  ``hdf5_path`` may be a clip-store directory (this module) -- h5py is not needed then.
"""
import json
import os

import numpy as np

STORE_META = 'store.json'


def int16_to_float32(x):
    """utils/utilities.py:66-67."""
    return (x / 32767.).astype(np.float32)


class ClipStore(object):
    """Directory of flat arrays: waveform.i16 (N, L), target.u8 (N, K), [strong_target.u8 (N, T, K)], names."""

    def __init__(self, path):
        self.path = path
        with open(os.path.join(path, STORE_META)) as f:
            m = json.load(f)
        self.audios_num, self.samples, self.classes_num = m['audios_num'], m['samples'], m['classes_num']
        self.frames_num = m.get('frames_num')
        self.audio_name = m['audio_name']
        self.waveform = np.memmap(os.path.join(path, 'waveform.i16'), dtype=np.int16, mode='r',
                                  shape=(self.audios_num, self.samples))
        self.target = np.memmap(os.path.join(path, 'target.u8'), dtype=np.uint8, mode='r',
                                shape=(self.audios_num, self.classes_num))
        self.strong_target = None
        if self.frames_num:
            self.strong_target = np.memmap(os.path.join(path, 'strong_target.u8'), dtype=np.uint8, mode='r',
                                           shape=(self.audios_num, self.frames_num, self.classes_num))

    def _wave_fd(self):
        fd = getattr(self, '_fd', None)
        if fd is None:
            fd = self._fd = os.open(os.path.join(self.path, 'waveform.i16'), os.O_RDONLY)
        return fd

    def __del__(self):
        fd = getattr(self, '_fd', None)
        if fd is not None:
            try:
                os.close(fd)
            except OSError:
                pass

    @staticmethod
    def write(path, audio_name, waveform_i16, target, strong_target=None):
        """Pack arrays (what utils/features.py:232-260 writes into HDF5) into a store directory."""
        os.makedirs(path, exist_ok=True)
        waveform_i16 = np.ascontiguousarray(waveform_i16, dtype=np.int16)
        target = np.ascontiguousarray(target).astype(np.uint8)
        n, samples = waveform_i16.shape
        waveform_i16.tofile(os.path.join(path, 'waveform.i16'))
        target.tofile(os.path.join(path, 'target.u8'))
        meta = {'audios_num': int(n), 'samples': int(samples), 'classes_num': int(target.shape[1]),
                'audio_name': [str(a) for a in audio_name]}
        if strong_target is not None:
            st = np.ascontiguousarray(strong_target).astype(np.uint8)
            st.tofile(os.path.join(path, 'strong_target.u8'))
            meta['frames_num'] = int(st.shape[1])
        with open(os.path.join(path, STORE_META), 'w') as f:
            json.dump(meta, f)
        return ClipStore(path)

    @staticmethod
    def from_hdf5(hdf5_path, path):
        """One-off conversion of a reference HDF5 pack (needs h5py, which the reference itself requires)."""
        import h5py
        with h5py.File(hdf5_path, 'r') as hf:
            names = [a.decode() for a in hf['audio_name'][:]]
            strong = hf['strong_target'][:] if 'strong_target' in hf.keys() else None
            return ClipStore.write(path, names, hf['waveform'][:], hf['target'][:], strong)

    def gather(self, indexes, out=None, threads=None):
        """Batch rows as numpy arrays: int16 waveforms (no host conversion), fp32 targets.  ``out``: optional dict of
        preallocated (e.g. pinned) arrays to fill.

        The waveform rows (640 KB each for 10 s clips) are copied straight from the memmap into their final place --
        no fancy-index temporary -- by ``threads`` workers over contiguous blocks of output rows (numpy releases the GIL
        inside the row copies); default: min(8, cores).  tools/bench_feed.py measures the rate."""
        idx = np.asarray(indexes, dtype=np.int64)
        n = len(idx)
        batch = {'audio_name': np.array([self.audio_name[i] for i in idx])}
        wave = out['waveform'] if out is not None and 'waveform' in out else np.empty((n, self.samples), dtype=np.int16)
        fd, row_bytes = self._wave_fd(), self.samples * 2

        def copy_rows(lo, hi):
            # positional reads from the page cache straight into the destination rows: no mapping faults (a first touch
            # of memmap pages costs a minor fault per 4 KB: 1.8 GB/s per thread here against 5+ with pread), GIL released
            for j in range(lo, hi):
                row = memoryview(wave[j]).cast('B')
                got = os.preadv(fd, [row], int(idx[j]) * row_bytes)
                if got != row_bytes:
                    raise IOError('short read of clip %d from %s' % (int(idx[j]), self.path))

        workers = threads if threads is not None else min(8, os.cpu_count() or 1)
        workers = max(1, min(workers, n))
        if workers == 1 or n * self.samples < (1 << 22):
            copy_rows(0, n)
        else:
            pool = _gather_pool(workers)
            step = (n + workers - 1) // workers
            futures = [pool.submit(copy_rows, lo, min(n, lo + step)) for lo in range(0, n, step)]
            for f in futures:
                f.result()
        batch['waveform'] = wave[:n] if wave.shape[0] != n else wave
        for key, arr in (('target', self.target), ('strong_target', self.strong_target)):
            if arr is None:
                continue
            if out is not None and key in out:
                np.take(arr, idx, axis=0, out=out[key], mode='clip') if out[key].dtype == arr.dtype else \
                    out[key].__setitem__(Ellipsis, arr[idx])
                batch[key] = out[key]
            else:
                batch[key] = arr[idx].astype(np.float32)
        return batch


_POOLS = {}


def _gather_pool(workers):
    pool = _POOLS.get(workers)
    if pool is None:
        from concurrent.futures import ThreadPoolExecutor
        pool = _POOLS[workers] = ThreadPoolExecutor(max_workers=workers, thread_name_prefix='sed-gather')
    return pool


_STORES = {}


def _store(path):
    s = _STORES.get(path)
    if s is None:
        s = _STORES[path] = ClipStore(path)
    return s


def _audios_num(path):
    if os.path.isdir(path):
        return _store(path).audios_num
    import h5py                                             # a genuine HDF5 pack: same call as the reference
    with h5py.File(path, 'r') as hf:
        return len(hf['audio_name'])


def _hdf5_item(hdf5_path, index_in_hdf5):
    """A genuine HDF5 pack: the reference's own per-item read (data_generator.py:37-47; needs h5py)."""
    import h5py
    with h5py.File(hdf5_path, 'r') as hf:
        d = {'audio_name': hf['audio_name'][index_in_hdf5].decode(),
             'waveform': int16_to_float32(hf['waveform'][index_in_hdf5]),
             'target': hf['target'][index_in_hdf5].astype(np.float32)}
        if 'strong_target' in hf.keys():
            d['strong_target'] = hf['strong_target'][index_in_hdf5].astype(np.float32)
    return d


class DCASE2017Task4Dataset(object):
    def __init__(self):
        pass

    def __getitem__(self, meta):
        """data_generator.py:20-49: {'audio_name', 'waveform' fp32 (x/32767), 'target' fp32[, 'strong_target']}."""
        if not os.path.isdir(meta['hdf5_path']):
            return _hdf5_item(meta['hdf5_path'], meta['index_in_hdf5'])
        s = _store(meta['hdf5_path'])
        i = int(meta['index_in_hdf5'])
        d = {'audio_name': s.audio_name[i], 'waveform': int16_to_float32(np.asarray(s.waveform[i])),
             'target': np.asarray(s.target[i]).astype(np.float32)}
        if s.strong_target is not None:
            d['strong_target'] = np.asarray(s.strong_target[i]).astype(np.float32)
        return d


class TrainSampler(object):
    def __init__(self, hdf5_path, batch_size, random_seed=1234):
        """data_generator.py:52-76."""
        self.hdf5_path = hdf5_path
        self.batch_size = batch_size
        self.random_state = np.random.RandomState(random_seed)
        self.audios_num = _audios_num(hdf5_path)
        self.audio_indexes = np.arange(self.audios_num)
        self.random_state.shuffle(self.audio_indexes)
        self.pointer = 0

    def next_indexes(self):
        """The ``index_in_hdf5`` values of the next batch (data_generator.py:87-101, one pass of the inner loop)."""
        out = np.empty(self.batch_size, dtype=np.int64)
        for i in range(self.batch_size):
            index = self.audio_indexes[self.pointer]
            self.pointer += 1
            if self.pointer >= self.audios_num:             # re-shuffle BEFORE the item is appended (reference order)
                self.pointer = 0
                self.random_state.shuffle(self.audio_indexes)
            out[i] = self.audio_indexes[index]              # the reference's double indirection
        return out

    def __iter__(self):
        while True:
            yield [{'hdf5_path': self.hdf5_path, 'index_in_hdf5': i} for i in self.next_indexes()]


class TestSampler(object):
    def __init__(self, hdf5_path, batch_size):
        """data_generator.py:104-120."""
        self.hdf5_path = hdf5_path
        self.batch_size = batch_size
        self.audios_num = _audios_num(hdf5_path)
        self.audio_indexes = np.arange(self.audios_num)

    def __iter__(self):
        pointer = 0
        while pointer < self.audios_num:
            batch_indexes = np.arange(pointer, min(pointer + self.batch_size, self.audios_num))
            yield [{'hdf5_path': self.hdf5_path, 'index_in_hdf5': self.audio_indexes[i]} for i in batch_indexes]
            pointer += self.batch_size


def collate_fn(list_data_dict):
    """data_generator.py:148-164."""
    return {key: np.array([d[key] for d in list_data_dict]) for key in list_data_dict[0].keys()}


class TrainBatcher(object):
    """The fast path the fused trainer uses instead of DataLoader(dataset, batch_sampler, collate_fn): the same index
    stream as ``TrainSampler``, batches gathered from the memmap as int16 into (optionally pinned) buffers.

    ``pinned=True`` keeps ``slots`` independent sets of pinned buffers and fills them round-robin: batch i lives in
    set i % slots (``batch['slot']``), so an asynchronous host->device copy of batch i (feed.DeviceFeed.submit) is not
    overwritten by the gather of batch i+1.  Before set k is refilled, ``before_refill(k)`` is called if given -- pass
    ``DeviceFeed.wait_copied`` so that the refill waits for the copy that last read that set."""

    def __init__(self, store_path, batch_size, random_seed=1234, pinned=False, slots=2, before_refill=None):
        self.sampler = TrainSampler(store_path, batch_size, random_seed)
        self.store = _store(store_path)
        self.buffers = None
        self.before_refill = before_refill
        if pinned:
            import torch
            s = self.store
            self._pinned = [{'waveform': torch.empty((batch_size, s.samples), dtype=torch.int16).pin_memory(),
                             'target': torch.empty((batch_size, s.classes_num), dtype=torch.float32).pin_memory()}
                            for _ in range(max(1, slots))]
            self.buffers = [{k: v.numpy() for k, v in d.items()} for d in self._pinned]

    def __iter__(self):
        i = 0
        while True:
            if self.buffers is None:
                yield self.store.gather(self.sampler.next_indexes())
            else:
                slot = i % len(self.buffers)
                if self.before_refill is not None:
                    self.before_refill(slot)
                batch = self.store.gather(self.sampler.next_indexes(), out=self.buffers[slot])
                batch['slot'] = slot
                batch['pinned'] = self._pinned[slot]
                yield batch
            i += 1
