"""A few fused training steps at the bench configuration -- target of `ncu -k regex:... --launch-skip ...`
(development tool, run under gpurun).   python tools/prof_step.py [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import models  # noqa: E402
from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = getattr(models, bench.MODEL)(*bench.CTOR).to(dev)
model.train()
trainer = FusedTrainer(model, lr=1e-3)
pcm, target_np = bench.synthetic_rank_batch(512, 0)
wave = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).to(dev)
tgt = torch.from_numpy(target_np).to(dev)
lam = torch.rand(512, device=dev)
for _ in range(steps):
    trainer.step(wave, tgt, lam)
torch.cuda.synchronize()
print('done')
