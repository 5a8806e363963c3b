"""Per-entry-point device time of one fused training step of any of the seven models (development tool).
    python tools/step_shares.py Cnn_9layers_Transformer_FrameAtt 128"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import _lib, models  # noqa: E402
from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else bench.MODEL
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = getattr(models, name)(*bench.CTOR).to(dev)
model.train()
trainer = FusedTrainer(model, lr=1e-3)
pcm, target_np = bench.synthetic_rank_batch(2 * bs, 0)
wave = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).to(dev)
tgt = torch.from_numpy(target_np).to(dev)
lam = torch.rand(2 * bs, device=dev)
for _ in range(3):
    trainer.step(wave, tgt, lam)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    trainer.step(wave, tgt, lam)
e1.record()
torch.cuda.synchronize()
print('%s batch_size %d + mixup: %.3f ms/step, %.0f raw clips/s' % (name, bs, e0.elapsed_time(e1) / 10,
                                                                     2 * bs * 10 / e0.elapsed_time(e1) * 1e3))
agg = {}
_lib.PROFILE = []
trainer.step(wave, tgt, lam)
torch.cuda.synchronize()
for n, tag, a, b in _lib.PROFILE:
    v = agg.setdefault(n, [0, 0.0])
    v[0] += 1
    v[1] += a.elapsed_time(b)
_lib.PROFILE = None
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print('  %-34s %3d %8.3f ms' % (n, v[0], v[1]))
