"""Kernel-alone timing of the Cout = 64 convolution: unstacked CTA-pair kernel vs the kw-stacked one (run under gpurun)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import conv

B = int(os.environ.get('B', 256))
conv.KWSTACK_MIN_CIN = 64
for (h, w, cin, stats) in [(1000, 64, 64, True), (1000, 64, 64, False), (500, 32, 128, False)]:
    x = torch.randn(B, h, w, cin, device='cuda').bfloat16()
    wt = torch.randn(64, cin, 3, 3, device='cuda') * 0.05
    wf, _ = conv.pack_weights(wt)
    for flag in (False, True):
        conv.USE_KWSTACK = flag
        for _ in range(3):
            conv.conv3x3(x, wf, 64, want_stats=stats)
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            conv.conv3x3(x, wf, 64, want_stats=stats)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        flops = 2.0 * B * h * w * 64 * cin * 9
        print('H=%d W=%d Cin=%d stats=%d kwstack=%d: %.3f ms  %.0f TFLOP/s' % (h, w, cin, stats, flag, best, flops / best / 1e9), flush=True)
