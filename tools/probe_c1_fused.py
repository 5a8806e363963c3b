"""Kernel-alone timing at the bench shape: BatchNorm-backward apply + Cin = 1 weight gradient as two passes vs the fused
kernel (sed_bn_apply_conv_c1_wgrad).  Run under gpurun."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import ops

B, H, W = int(os.environ.get('B', 256)), 1000, 64
x = torch.randn(B, H, W, device='cuda')
y = torch.randn(B, H, W, 64, device='cuda').bfloat16()
dA = (torch.randn(B, H, W, 64, device='cuda') * 0.1).bfloat16()
bn = torch.nn.BatchNorm2d(64).cuda()
stats = torch.stack([y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))])[None].contiguous()
st = ops.bn_finalize(stats, B * H * W, bn)
coef = ops.bn_bwd_coef(y, dA, st, bn, 1, 1, None, None)
gw = torch.empty(64, 1, 3, 3, device='cuda')


def t(fn, reps=5):
    for _ in range(2):
        fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


dy = ops.bn_bwd_apply(y, dA, st, coef, 1, 1)
ta = t(lambda: ops.bn_bwd_apply(y, dA, st, coef, 1, 1))
tw = t(lambda: ops.conv_c1_wgrad(x, dy, gw))
tf = t(lambda: ops.bn_apply_conv_c1_wgrad(x, y, dA, st, coef, gw))
gb = B * H * W * 64 * 2 / 1e9
print('apply %.3f ms (%.0f GB/s)  wgrad %.3f ms (%.0f GB/s)  fused %.3f ms (%.0f GB/s of y + dA + dY)'
      % (ta, 3 * gb / ta * 1e3, tw, gb / tw * 1e3, tf, 3 * gb / tf * 1e3))
