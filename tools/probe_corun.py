"""Can an HBM-bound BatchNorm kernel run beside the tensor-bound halo conv on the same SMs?  Two streams, per layer
shape of a HALF batch: conv alone, BN forward alone, BN backward (reduce + apply) alone, and each BN beside the conv.
(Feasibility probe for a two-half software pipeline of the trunk; run under gpurun.)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sound_event_detection_dcase2017_task4_b200 import conv as tcconv, ops

dev = torch.device('cuda')
B = int(os.environ.get('B', 128))
shapes = [(1000, 64, 64, 64), (500, 32, 64, 128), (500, 32, 128, 128), (250, 16, 128, 256), (250, 16, 256, 256),
          (125, 8, 256, 512), (125, 8, 512, 512)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fns, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st, fn in fns:
            st.wait_event(e0)
        for st, fn in fns:
            with torch.cuda.stream(st):
                fn()
        for st, fn in fns:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


class BN(object):
    pass


for (h, w, cin, cout) in shapes:
    x = torch.randn(B, h, w, cin, device=dev).bfloat16()
    wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
    wf, wd = tcconv.pack_weights(wt)
    y2 = torch.randn(B, h, w, cout, device=dev).bfloat16()          # the OTHER half's conv output
    bn = torch.nn.BatchNorm2d(cout).to(dev)
    part = torch.rand(64, 2, cout, device=dev) * 100
    st = ops.bn_finalize(part, 64 * 50, bn)
    dA = torch.randn(B, h, w, cout, device=dev).bfloat16()
    conv = lambda: tcconv.conv3x3(x, wf, cout, want_stats=True)
    bnf = lambda: ops.bn_relu_pool_fwd(y2, st, 1, 1)
    bnb = lambda: ops.bn_relu_pool_bwd(y2, dA, st, bn, 1, 1, None, None)
    for _ in range(2):
        conv(); bnf(); bnb()
    tc = timed([(s1, conv)])
    tf = timed([(s2, bnf)])
    tb = timed([(s2, bnb)])
    tcf = timed([(s1, conv), (s2, bnf)])
    tcb = timed([(s1, conv), (s2, bnb)])
    print('H=%4d W=%2d %3d->%3d  conv %.3f  bn_fwd %.3f  bn_bwd %.3f | conv||bn_fwd %.3f (serial %.3f)  conv||bn_bwd %.3f (serial %.3f)'
          % (h, w, cin, cout, tc, tf, tb, tcf, tc + tf, tcb, tc + tb), flush=True)
