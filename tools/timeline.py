"""Device timeline of ONE training step: start / end of every libsedb200 launch relative to the step's first
launch, per stream (CUDA events on the launching stream; development tool, run under gpurun).

    python tools/timeline.py [--batch 256] [--model Cnn_9layers_Gru_FrameAtt] [--from sed_bce] > timeline.txt

Used to check that the side-stream weight gradients really run concurrently with the BatchNorm backward
(engine.trunk_backward): overlapping intervals on different streams show up in the `||` column.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import _lib, models  # noqa: E402
from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--model', default=bench.MODEL)
ap.add_argument('--set', action='append', default=[], help='module.attr=int switches, e.g. engine.OVERLAP_WGRAD=0')
args = ap.parse_args()
import importlib
for spec in args.set:
    target, v = spec.split('=')
    modname, attr = target.rsplit('.', 1)
    setattr(importlib.import_module('sound_event_detection_dcase2017_task4_b200.' + modname), attr, bool(int(v)))

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = getattr(models, args.model)(*bench.CTOR).to(dev)
model.train()
trainer = FusedTrainer(model, lr=1e-3)
b2 = 2 * args.batch
pcm, target_np = bench.synthetic_rank_batch(b2, 0)
wave = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).to(dev)
tgt = torch.from_numpy(target_np).to(dev)
lam = torch.rand(b2, device=dev)
for _ in range(4):
    trainer.step(wave, tgt, lam)
torch.cuda.synchronize()

# stream id of every call: patch _lib.call to remember the stream the events were recorded on
records = []
orig_call = _lib.call


def call(name, *a):
    st = torch.cuda.current_stream().cuda_stream
    n0 = len(_lib.PROFILE)
    orig_call(name, *a)
    for rec in _lib.PROFILE[n0:]:
        records.append((rec[0], st, rec[2], rec[3]))


_lib.PROFILE = []
for m in list(sys.modules.values()):
    if getattr(m, 'call', None) is orig_call:
        m.call = call
_lib.call = call
base = torch.cuda.Event(enable_timing=True)
base.record()
e_end = torch.cuda.Event(enable_timing=True)
trainer.step(wave, tgt, lam)
e_end.record()
torch.cuda.synchronize()
streams = {}
rows = []
for name, st, e0, e1 in records:
    sid = streams.setdefault(st, len(streams))
    rows.append((base.elapsed_time(e0), base.elapsed_time(e1), sid, name))
rows.sort()
print('# step total %.3f ms (events add ~3-5 us per launch); streams: %d' % (base.elapsed_time(e_end), len(streams)))
print('# %9s %9s %8s  s  ||  kernel' % ('start_ms', 'end_ms', 'dur_ms'))
for i, (t0, t1, sid, name) in enumerate(rows):
    par = [r[3] for r in rows if r[2] != sid and r[0] < t1 and r[1] > t0]
    print('%10.3f %9.3f %8.3f  %d  %-2s  %s%s' % (t0, t1, t1 - t0, sid, '||' if par else '', name,
                                                ('   <- with ' + ','.join(sorted(set(par)))) if par else ''))
busy = {}
for t0, t1, sid, name in rows:
    busy[sid] = busy.get(sid, 0.0) + (t1 - t0)
print('# busy per stream (ms):', {k: round(v, 3) for k, v in busy.items()})
