"""Mint golden index streams from the UNMODIFIED reference samplers (/root/reference/utils/data_generator.py).
h5py is absent here, so a stub ``h5py.File`` that only answers ``len(hf['audio_name'])`` stands in for the pack;
the sampler logic that runs is the reference's own.  Run in the authoring container only.

    python tests/golden/make_golden_sampler.py
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def main():
    sys.path.insert(0, os.path.join(REF, 'utils'))
    for name in ('librosa', 'sed_eval', 'matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    h5 = types.ModuleType('h5py')

    class File(object):
        def __init__(self, path, mode):
            self.n = int(os.path.basename(path).split('.')[0])

        def __enter__(self):
            return {'audio_name': [b''] * self.n}

        def __exit__(self, *a):
            return False
    h5.File = File
    sys.modules['h5py'] = h5
    import data_generator as ref
    gold = {'train': [], 'test': []}
    for audios_num, batch_size, batches in ((10, 4, 12), (7, 7, 5), (51172, 64, 6), (100, 32, 16), (3, 8, 4)):
        it = iter(ref.TrainSampler('%d.h5' % audios_num, batch_size))
        seq = [[int(m['index_in_hdf5']) for m in next(it)] for _ in range(batches)]
        gold['train'].append({'audios_num': audios_num, 'batch_size': batch_size, 'batches': seq})
    for audios_num, batch_size in ((10, 4), (8, 4), (1, 16), (488, 64)):
        seq = [[int(m['index_in_hdf5']) for m in b] for b in ref.TestSampler('%d.h5' % audios_num, batch_size)]
        gold['test'].append({'audios_num': audios_num, 'batch_size': batch_size, 'batches': seq})
    with open(os.path.join(HERE, 'sampler_golden.json'), 'w') as f:
        json.dump(gold, f)
    print('train streams: %d, test streams: %d' % (len(gold['train']), len(gold['test'])))


if __name__ == '__main__':
    main()
