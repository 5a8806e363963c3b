"""BASELINE.json config 5: log-mel-only throughput sweep -- clip length 1-30 s x batch 1-4096 (one GPU per process;
for N GPUs the path is pure replicas, launch N copies).  One JSON object per line:
Mframes/s, achieved GB/s (fp32 in + fp32 log-mel out) and the fraction of the measured HBM peak.
Development / evidence tool, run under gpurun:   python tools/logmel_sweep.py [--int16] > gpurun_out/logmel_sweep.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import frontend as fe  # noqa: E402


def main():
    int16 = '--int16' in sys.argv
    peak = bench.measured_peaks()['hbm_gbs']
    bank = fe.MelBankCSR(torch.from_numpy(fe.mel_weight_matrix(32000, 1024, 64, 50, 14000)).cuda())
    for seconds in (1, 2, 5, 10, 20, 30):
        n = 32000 * seconds
        t = n // 320 + 1
        for batch in (1, 4, 16, 64, 256, 1024, 4096):
            in_bytes = batch * n * (2 if int16 else 4)
            if in_bytes + batch * t * 256 > 40e9:
                continue
            if int16:
                wave = torch.randint(-8192, 8192, (batch, n), dtype=torch.int16, device='cuda')
            else:
                wave = (torch.rand(batch, n, device='cuda') - 0.5) * 0.5
            out = torch.empty((batch, 1, t, 64), device='cuda')
            for _ in range(3):
                fe.logmel(wave, 320, bank, out=out)
            reps = 20 if batch * seconds <= 2560 else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fe.logmel(wave, 320, bank, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nbytes = in_bytes + batch * t * 64 * 4
            print(json.dumps({'seconds': seconds, 'batch': batch, 'input': 'int16' if int16 else 'fp32', 'ms': round(ms, 4),
                              'mframes_per_s': round(batch * t / ms / 1e3, 1), 'gbs': round(nbytes / ms / 1e6, 1),
                              'hbm_frac': round(nbytes / ms / 1e6 / peak, 4)}), flush=True)
            del wave, out


if __name__ == '__main__':
    main()
