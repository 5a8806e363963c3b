"""Per-kernel device timing (CUDA events, after warm-up) -- development tool, run under gpurun.
    python tools/bench_kernels.py [--batch 64] [--what logmel,conv]
Writes one JSON object per line to stdout."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import conv, frontend as fe, ops  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--what', default='logmel,conv')
    ap.add_argument('--pair', action='store_true', help='use the cta_group::2 conv kernel where it applies')
    args = ap.parse_args()
    what = args.what.split(',')
    conv.USE_2CTA = args.pair
    B = args.batch
    if 'logmel' in what:
        melW = torch.from_numpy(fe.mel_weight_matrix(32000, 1024, 64, 50, 14000)).cuda()
        bank = fe.MelBankCSR(melW)
        for clips in (64, 512):
            wave = (torch.rand(clips, 320000, device='cuda') - 0.5) * 0.5
            pcm = (wave * 32767).to(torch.int16)
            out = torch.empty((clips, 1, 1001, 64), device='cuda')
            for name, inp, bytes_per_clip in (('f32', wave, 4 * 320000 + 4 * 1001 * 64),
                                              ('i16', pcm, 2 * 320000 + 4 * 1001 * 64)):
                med, best = timeit(lambda: fe.logmel(inp, 320, bank, out=out))
                print(json.dumps({'kernel': 'logmel_' + name, 'clips': clips, 'ms': med, 'best_ms': best,
                                  'Mframes_s': clips * 1001 / med / 1e3,
                                  'GBs': clips * bytes_per_clip / med / 1e6}))
    if 'conv' in what:
        layers = [(1001, 64, 64, 64), (500, 32, 64, 128), (500, 32, 128, 128), (250, 16, 128, 256),
                  (250, 16, 256, 256), (125, 8, 256, 512), (125, 8, 512, 512)]
        for (H, W, Cin, Cout) in layers:
            x = torch.randn(B, H, W, Cin, device='cuda').to(torch.bfloat16)
            dy = torch.randn(B, H, W, Cout, device='cuda').to(torch.bfloat16)
            w = torch.randn(Cout, Cin, 3, 3, device='cuda') * 0.05
            wf, wd = conv.pack_weights(w)
            flops = 2.0 * B * H * W * Cin * Cout * 9
            for name, fn in (('fwd', lambda: conv.conv3x3(x, wf, Cout, want_stats=True)),
                             ('dgrad', lambda: conv.conv3x3(dy, wd, Cin)),
                             ('wgrad', lambda: conv.conv3x3_wgrad(dy, x))):
                med, best = timeit(fn, iters=5, warmup=2)
                print(json.dumps({'kernel': 'conv_' + name, 'B': B, 'H': H, 'W': W, 'Cin': Cin, 'Cout': Cout,
                                  'ms': med, 'best_ms': best, 'TFLOPs': flops / med / 1e9}))
            del x, dy
    if 'bn' in what:
        class _BN(object):
            pass
        for (H, W, C, pool) in [(1001, 64, 64, 1), (1001, 64, 64, 2), (500, 32, 128, 1), (500, 32, 128, 2),
                                (250, 16, 256, 1), (250, 16, 256, 2), (125, 8, 512, 1)]:
            y = torch.randn(B, H, W, C, device='cuda').to(torch.bfloat16)
            bn = torch.nn.BatchNorm2d(C).cuda()
            partial = torch.zeros(4, 2, C, device='cuda')
            partial[0, 0] = 0.1 * B * H * W
            partial[0, 1] = 1.01 * B * H * W
            st = ops.bn_finalize(partial, B * H * W, bn)
            dA = torch.randn(B, H // pool, W // pool, C, device='cuda').to(torch.bfloat16)
            dg, db = torch.empty(C, device='cuda'), torch.empty(C, device='cuda')
            ybytes, abytes = y.numel() * 2, dA.numel() * 2
            med, best = timeit(lambda: ops.bn_relu_pool_fwd(y, st, pool, pool), iters=5, warmup=2)
            print(json.dumps({'kernel': 'bn_fwd', 'H': H, 'W': W, 'C': C, 'pool': pool, 'ms': med,
                              'GBs': (ybytes + abytes) / med / 1e6}))
            med, best = timeit(lambda: ops.bn_relu_pool_bwd(y, dA, st, bn, pool, pool, dg, db), iters=5, warmup=2)
            print(json.dumps({'kernel': 'bn_bwd(reduce+finalize+apply)', 'H': H, 'W': W, 'C': C, 'pool': pool, 'ms': med,
                              'GBs': (3 * ybytes + 2 * abytes) / med / 1e6}))
            del y, dA
    if 'c1' in what:
        H, W, Cout = 1001, 64, 64
        x0 = torch.randn(B, H, W, device='cuda')
        w = torch.randn(Cout, 1, 3, 3, device='cuda') * 0.3
        dy = torch.randn(B, H, W, Cout, device='cuda').to(torch.bfloat16)
        dw = torch.empty(Cout, 1, 3, 3, device='cuda')
        ybytes = dy.numel() * 2
        for name, fn in (('fwd', lambda: ops.conv_c1_fwd(x0, w)), ('wgrad', lambda: ops.conv_c1_wgrad(x0, dy, dw)),
                         ('dgrad', lambda: ops.conv_c1_dgrad(dy, w))):
            med, best = timeit(fn, iters=5, warmup=2)
            print(json.dumps({'kernel': 'conv_c1_' + name, 'B': B, 'ms': med, 'GBs': (ybytes + x0.numel() * 4) / med / 1e6}))
    if 'lin' in what:
        R, C, K = B * 125, 512, 17
        x = torch.randn(R, C, device='cuda')
        Wt = torch.randn(K, C, device='cuda') * 0.05
        bias = torch.zeros(K, device='cuda')
        dout = torch.randn(R, K, device='cuda')
        dW, dbias = torch.empty(K, C, device='cuda'), torch.empty(K, device='cuda')
        for name, fn in (('fwd', lambda: ops.linear_small_fwd(x, Wt, bias)),
                         ('bwd(dx+dw+reduce)', lambda: ops.linear_small_bwd(dout, x, Wt, dW, dbias))):
            med, best = timeit(fn, iters=5, warmup=2)
            print(json.dumps({'kernel': 'linear_small_' + name, 'R': R, 'ms': med}))


if __name__ == '__main__':
    main()
