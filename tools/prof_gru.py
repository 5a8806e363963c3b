"""Profiling driver: one forward + backward of the bidirectional GRU at the bench shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import temporal

B, T = int(os.environ.get('B', 256)), 125
torch.manual_seed(0)
gru = torch.nn.GRU(512, 256, num_layers=1, bias=True, batch_first=True, bidirectional=True).cuda()
x = torch.randn(B, T, 512, device='cuda') * 0.5
for it in range(3):
    out, ctx = temporal.gru_forward(gru, x, keep=True)
    grads = {}
    dx = temporal.gru_backward(gru, ctx, torch.randn_like(out), lambda p: grads.setdefault(p, torch.empty_like(p)))
torch.cuda.synchronize()
best = [1e9, 1e9]
dout = torch.randn_like(out)
for it in range(int(os.environ.get('REPS', 10))):
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record()
    out, ctx = temporal.gru_forward(gru, x, keep=True)
    e1.record()
    dx = temporal.gru_backward(gru, ctx, dout, lambda p: grads.setdefault(p, torch.empty_like(p)))
    e2.record()
    torch.cuda.synchronize()
    best = [min(best[0], e0.elapsed_time(e1)), min(best[1], e1.elapsed_time(e2))]
print('gru fwd %.3f ms  bwd %.3f ms (best of reps, incl. projections)' % tuple(best))
