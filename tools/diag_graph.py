"""Which tensors differ between an eager step and its CUDA-graph replay (development diagnostic, run under gpurun)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import models  # noqa: E402
from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer  # noqa: E402

dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
b2 = 2 * B
pcm, target_np = bench.synthetic_rank_batch(b2, 0)
wave = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).to(dev)
tgt = torch.from_numpy(target_np).to(dev)
lam = torch.rand(b2, device=dev, generator=torch.Generator(device=dev).manual_seed(3))


def run(use_graph, n):
    torch.manual_seed(0)
    model = getattr(models, bench.MODEL)(*bench.CTOR).to(dev)
    model.train()
    tr = FusedTrainer(model, lr=1e-3, use_graph=use_graph)
    torch.manual_seed(5)
    for _ in range(n):
        loss = tr.step(wave, tgt, lam)
    torch.cuda.synchronize()
    return tr, float(loss)


for n in (3, 4):
    te, le = run(False, n)
    tg, lg = run(True, n)
    print('after %d steps: loss eager %.9g graph %.9g' % (n, le, lg))
    off = 0
    for (k, p) in te.model.named_parameters():
        if not p.requires_grad:
            continue
        cnt = p.numel()
        ge, gg = te.flat_grad[off:off + cnt], tg.flat_grad[off:off + cnt]
        pe, pg = te.flat_param[off:off + cnt], tg.flat_param[off:off + cnt]
        off += cnt
        if not torch.equal(ge, gg) or not torch.equal(pe, pg):
            print('  %-40s grad max|d| %.3e (|g| %.3e)  param max|d| %.3e' % (
                k, (ge - gg).abs().max().item(), ge.abs().max().item(), (pe - pg).abs().max().item()))
    for k in ('exp_avg', 'exp_avg_sq', 'max_exp_avg_sq'):
        print('  %s equal: %s' % (k, torch.equal(getattr(te, k), getattr(tg, k))))
