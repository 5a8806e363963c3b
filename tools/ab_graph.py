"""Eager launches vs CUDA-graph replay of the fused training step: bit-equality of the loss trajectories and ms/step
(development tool, run under gpurun).   python tools/ab_graph.py [--steps 30] [--batch 256]"""
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
ap.add_argument('--steps', type=int, default=30)
ap.add_argument('--batch', type=int, default=256)
args = ap.parse_args()
dev = torch.device('cuda', 0)
b2 = 2 * args.batch
pcm, target_np = bench.synthetic_rank_batch(b2, 0)
wave = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).to(dev)
tgt = torch.from_numpy(target_np).to(dev)
lam = torch.rand(b2, device=dev, generator=torch.Generator(device=dev).manual_seed(3))


def run(use_graph):
    torch.manual_seed(0)
    model = getattr(models, bench.MODEL)(*bench.CTOR).to(dev)
    model.train()
    tr = FusedTrainer(model, lr=1e-3, use_graph=use_graph)
    torch.manual_seed(5)                                  # SpecAugment stream
    losses = [float(tr.step(wave, tgt, lam)) for _ in range(6)]
    for _ in range(3):
        tr.step(wave, tgt, lam)
    torch.cuda.synchronize()
    out = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            tr.step(wave, tgt, lam)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / args.steps)
    return losses, out, tr


l_e, t_e, _ = run(False)
l_g, t_g, tr = run(True)
print('eager losses', l_e)
print('graph losses', l_g)
print('bit-equal trajectories:', l_e == l_g, '| launches captured per step:', tr.graph_launches)
print('eager ms/step', [round(x, 3) for x in t_e])
print('graph ms/step', [round(x, 3) for x in t_g])
