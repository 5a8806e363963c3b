"""A/B timing of the fused training step under module-level switches (development tool, run under gpurun).
    python tools/ab_step.py engine.OVERLAP_WGRAD=0,1 [--steps 20] [--rounds 3]"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import models  # noqa: E402
from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer  # noqa: E402

spec = sys.argv[1]
steps = int(sys.argv[sys.argv.index('--steps') + 1]) if '--steps' in sys.argv else 20
rounds = int(sys.argv[sys.argv.index('--rounds') + 1]) if '--rounds' in sys.argv else 3
target, values = spec.split('=')
modname, attr = target.rsplit('.', 1)
mod = importlib.import_module('sound_event_detection_dcase2017_task4_b200.' + modname)
import ast
values = [ast.literal_eval(v) for v in values.split(';')] if ';' in values or '(' in values else [int(v) for v in values.split(',')]

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = getattr(models, bench.MODEL)(*bench.CTOR).to(dev)
model.train()
trainer = FusedTrainer(model, lr=1e-3)
pcm, target_np = bench.synthetic_rank_batch(512, 0)
wave = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).to(dev)
tgt = torch.from_numpy(target_np).to(dev)
lam = torch.rand(512, device=dev)
for _ in range(3):
    trainer.step(wave, tgt, lam)
for r in range(rounds):
    for v in values:
        setattr(mod, attr, bool(v) if isinstance(v, int) else frozenset(v))
        trainer.step(wave, tgt, lam)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            trainer.step(wave, tgt, lam)
        e1.record()
        torch.cuda.synchronize()
        print('%s=%s round %d: %.3f ms/step' % (target, v, r, e0.elapsed_time(e1) / steps), flush=True)
