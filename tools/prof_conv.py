"""One launch of every tensor-core conv kernel of a training step (fwd, dgrad, wgrad per layer) at the
bench batch size -- target of `ncu --set full -k regex:conv3x3` (development tool, run under gpurun)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import conv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
layers = [(1001, 64, 64, 64), (500, 32, 64, 128), (500, 32, 128, 128), (250, 16, 128, 256),
          (250, 16, 256, 256), (125, 8, 256, 512), (125, 8, 512, 512)]
for (H, W, Cin, Cout) in layers:
    x = torch.randn(B, H, W, Cin, device='cuda').to(torch.bfloat16)
    dy = torch.randn(B, H, W, Cout, device='cuda').to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda') * 0.05
    wf, wd = conv.pack_weights(w)
    conv.conv3x3(x, wf, Cout, want_stats=True)
    conv.conv3x3(dy, wd, Cin)
    conv.conv3x3_wgrad(dy, x)
    torch.cuda.synchronize()
    del x, dy
print('done')
