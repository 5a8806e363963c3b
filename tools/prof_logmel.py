"""A few launches of the fused log-mel kernel at the bench shape (512 raw 10 s clips, fp32) -- target of
`ncu --set full -k regex:logmel` (development tool, run under gpurun)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import frontend as fe  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bank = fe.MelBankCSR(torch.from_numpy(fe.mel_weight_matrix(32000, 1024, 64, 50, 14000)).cuda())
wave = (torch.rand(B, 320000, device='cuda') - 0.5) * 0.5
out = torch.empty(B, 1, 1001, 64, device='cuda')
for _ in range(4):
    fe.logmel(wave, 320, bank, out=out)
torch.cuda.synchronize()
print('done')
