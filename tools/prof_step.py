"""A few fused training steps at the bench configuration -- target of ncu (development tool, run under gpurun).

    python tools/prof_step.py [--steps 3] [--model NAME] [--batch 256] [--range]

--range: bracket only the LAST step with cudaProfilerStart/Stop (use with `ncu --profile-from-start off`), so that one
whole step is captured after the warm-up steps."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import models  # noqa: E402
from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--model', default=bench.MODEL)
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--range', action='store_true')
args = ap.parse_args()
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
for i in range(args.steps):
    last = i == args.steps - 1
    if args.range and last:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    trainer.step(wave, tgt, lam)
    if args.range and last:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
torch.cuda.synchronize()
print('done')
