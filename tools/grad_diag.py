"""Development diagnostic (run under gpurun): per-parameter gradient error of one train step of the
CUDA path against (a) the pure-fp32 oracle and (b) the oracle with bf16 storage emulation."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import sed, bf16_emulation  # noqa: E402
from sound_event_detection_dcase2017_task4_b200 import losses, models, pytorch_utils  # noqa: E402

CTOR = (32000, 1024, 320, 64, 50, 14000, 17)


def run(name, n_clips, n_samples):
    torch.manual_seed(0)
    ref = sed.build(name)
    emu = bf16_emulation.emulate_bf16_storage(copy.deepcopy(ref))
    torch.manual_seed(0)
    mine = getattr(models, name)(*CTOR).cuda()
    _, wave, target = sed.synthetic_batch(n_clips, n_samples, seed=1234)
    wave, target = torch.from_numpy(wave), torch.from_numpy(target)
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(n_clips))
    outs = {}
    for tag, m in (('fp32', ref), ('emu', emu)):
        m.train()
        torch.manual_seed(1)
        o = m(wave, lam)
        loss = sed.clip_bce(o, {'target': sed.mix_pairs(target, lam)})
        loss.backward()
        outs[tag] = (o['clipwise_output'].detach(), loss.item(), {k: p.grad for k, p in m.named_parameters()})
    mine.train()
    torch.manual_seed(1)
    o = mine(wave.cuda(), lam.cuda())
    loss = losses.clip_bce(o, {'target': pytorch_utils.do_mixup(target.cuda(), lam.cuda())})
    loss.backward()
    clip = o['clipwise_output'].detach().cpu()
    print('== %s  %d clips x %d samples' % (name, n_clips, n_samples))
    for tag in ('fp32', 'emu'):
        c, l, _ = outs[tag]
        print('  vs %-4s: clip rel %.2e  loss rel %.2e' % (tag, ((clip - c).abs() / c.abs()).max().item(),
                                                           abs(loss.item() - l) / abs(l)))
    c, l, _ = outs['emu']
    c0, l0, _ = outs['fp32']
    print('  emu vs fp32: clip rel %.2e' % ((c - c0).abs() / c0.abs()).max().item())
    print('  %-40s %10s %10s %10s %10s' % ('param', 'L2 fp32', 'cos fp32', 'L2 emu', 'emu-fp32'))
    for k, p in mine.named_parameters():
        g0, g1 = outs['fp32'][2][k], outs['emu'][2][k]
        if g0 is None:
            continue
        a = p.grad.cpu().double().flatten()
        b0, b1 = g0.double().flatten(), g1.double().flatten()
        e0 = (a - b0).norm().item() / b0.norm().item()
        e1 = (a - b1).norm().item() / b1.norm().item()
        e01 = (b1 - b0).norm().item() / b0.norm().item()
        cos = torch.dot(a, b0).item() / (a.norm().item() * b0.norm().item())
        print('  %-40s %10.3e %10.5f %10.3e %10.3e' % (k, e0, cos, e1, e01))


if __name__ == '__main__':
    run('Cnn_9layers_FrameAvg', 4, 32000)
    run('Cnn_9layers_FrameAvg', 8, 96000)
    run('Cnn_9layers_Gru_FrameAtt', 4, 32000)
