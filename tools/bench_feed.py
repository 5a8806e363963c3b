"""Host feed rate of the clip store (row f2): TrainBatcher gathers batches of raw int16 clips from the memory-mapped
store into pinned buffers (and, with --h2d, DeviceFeed copies them to the device).  The B200 training step consumes
512 clips (328 MB of PCM) per 28 ms = 11.7 GB/s; the reference's loader opens an HDF5 file per item in 8 workers and
converts every clip to fp32 on the host.

    python tools/bench_feed.py [--clips 4096] [--batch 512] [--batches 20] [--threads 1,4,8,16] [--h2d]
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import data_generator as dg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--clips', type=int, default=4096)
ap.add_argument('--batch', type=int, default=512)
ap.add_argument('--batches', type=int, default=20)
ap.add_argument('--threads', default='1,4,8,16')
ap.add_argument('--h2d', action='store_true')
args = ap.parse_args()

samples = 320000
root = tempfile.mkdtemp(prefix='sed_store_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
try:
    rs = np.random.RandomState(0)
    wave = rs.randint(-3000, 3000, size=(args.clips, samples)).astype(np.int16)
    target = (rs.rand(args.clips, 17) > 0.8).astype(np.uint8)
    dg.ClipStore.write(root, ['clip%05d.wav' % i for i in range(args.clips)], wave, target)
    del wave
    import torch
    pinned = torch.cuda.is_available()
    for th in [int(t) for t in args.threads.split(',')]:
        batcher = dg.TrainBatcher(root, args.batch, pinned=pinned, slots=2)
        store = batcher.store
        real_gather = store.gather
        store.gather = lambda idx, out=None, _t=th: real_gather(idx, out=out, threads=_t)
        feed = args.h2d and pinned
        it = iter(batcher)
        for _ in range(2):
            next(it)
        t0 = time.perf_counter()
        for i in range(args.batches):
            b = next(it)
            if feed:
                dev = torch.empty(b['pinned']['waveform'].shape, dtype=torch.int16, device='cuda')
                dev.copy_(b['pinned']['waveform'], non_blocking=True)
                torch.cuda.current_stream().synchronize()
        dt = time.perf_counter() - t0
        nbytes = args.batches * args.batch * samples * 2
        print('threads %2d: %.1f ms per batch of %d clips  = %6.0f clips/s  %.2f GB/s%s'
              % (th, dt / args.batches * 1e3, args.batch, args.batches * args.batch / dt, nbytes / dt / 1e9,
                 '  (incl. synchronous H2D)' if feed else ''), flush=True)
        store.gather = real_gather
finally:
    shutil.rmtree(root, ignore_errors=True)
