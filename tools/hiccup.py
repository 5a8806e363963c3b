"""Per-step device times of a long run of training steps + what the host did meanwhile (caching-allocator segment
allocations, Python GC pauses): finds the cause of sporadic slow steps (development tool, run under gpurun).
    python tools/hiccup.py [--steps 100] [--set engine.OVERLAP_WGRAD=0]"""
import argparse
import gc
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import models  # noqa: E402
from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=100)
ap.add_argument('--set', action='append', default=[])
ap.add_argument('--nogc', action='store_true')
args = ap.parse_args()
for spec in args.set:
    target, v = spec.split('=')
    modname, attr = target.rsplit('.', 1)
    setattr(importlib.import_module('sound_event_detection_dcase2017_task4_b200.' + modname), attr, bool(int(v)))

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = getattr(models, bench.MODEL)(*bench.CTOR).to(dev)
model.train()
trainer = FusedTrainer(model, lr=1e-3)
pcm, target_np = bench.synthetic_rank_batch(512, 0)
wave = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).to(dev)
tgt = torch.from_numpy(target_np).to(dev)
lam = torch.rand(512, device=dev)
for _ in range(5):
    trainer.step(wave, tgt, lam)
torch.cuda.synchronize()
gc_log = []
t_gc = [0.0]


def on_gc(phase, info):
    if phase == 'start':
        t_gc[0] = time.perf_counter()
    else:
        gc_log.append((info['generation'], (time.perf_counter() - t_gc[0]) * 1e3))


gc.callbacks.append(on_gc)
if args.nogc:
    gc.disable()
marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
host = []
segs = []
marks[0].record()
for i in range(args.steps):
    t0 = time.perf_counter()
    trainer.step(wave, tgt, lam)
    marks[i + 1].record()
    host.append((time.perf_counter() - t0) * 1e3)
    st = torch.cuda.memory_stats()
    segs.append((st['segment.all.allocated'], st['num_alloc_retries'], st['reserved_bytes.all.current'] >> 20))
torch.cuda.synchronize()
dur = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
print('device ms/step: median %.3f min %.3f max %.3f mean %.3f' % (float(np.median(dur)), min(dur), max(dur), float(np.mean(dur))))
print('host enqueue ms/step: median %.3f max %.3f' % (float(np.median(host)), max(host)))
print('segments allocated first/last: %s / %s' % (segs[0], segs[-1]))
for i, d in enumerate(dur):
    if d > 1.15 * np.median(dur) or host[i] > 3 * np.median(host):
        print('step %3d: device %.2f ms, host enqueue %.2f ms, segs %s' % (i, d, host[i], segs[i]))
print('gc events (generation, ms):', [(g, round(ms, 2)) for g, ms in gc_log if ms > 0.5][:40], 'total', len(gc_log))
