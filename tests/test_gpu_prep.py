"""-m gpu: bn0 + SpecAugment + mixup stage (csrc/prep.cu) against the oracle's modules run on CPU
(models.py:202-211; torchlibrosa SpecAugmentation; pytorch_utils.do_mixup)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def test_spec_augment_module_bit_exact_with_oracle_rng():
    """Seam A: the stand-alone SpecAugmentation drop-in zeroes exactly the oracle's stripes."""
    from oracle import frontend as ofe
    from sound_event_detection_dcase2017_task4_b200.dropin.torchlibrosa.augmentation import SpecAugmentation
    x = torch.randn(6, 1, 101, 64)
    ref_mod = ofe.SpecAugmentation(time_drop_width=64, time_stripes_num=2, freq_drop_width=8, freq_stripes_num=2)
    mine = SpecAugmentation(time_drop_width=64, time_stripes_num=2, freq_drop_width=8, freq_stripes_num=2)
    ref_mod.train(); mine.train()
    torch.manual_seed(7)
    ref = ref_mod(x.clone())
    after_ref = torch.get_rng_state()
    torch.manual_seed(7)
    got = mine(x.clone().cuda())
    assert torch.equal(got.cpu(), ref)
    assert torch.equal(after_ref, torch.get_rng_state())        # same RNG consumption
    mine.eval()
    assert torch.equal(mine(x.cuda()).cpu(), x)                 # no-op in eval mode


@pytest.mark.parametrize('mixup', [False, True])
def test_bn0_aug_mix_forward_backward(mixup):
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import ops, specaug
    B2, T, M = 6, 101, 64
    g = torch.Generator().manual_seed(0)
    logmel = (torch.randn(B2, T, M, generator=g) * 10 - 30)
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(B2)) if mixup else None
    bn = nn.BatchNorm2d(M)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(M, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(M, generator=g))
    # reference on CPU: transpose / bn0 / transpose, stripes, mixup (models.py:202-211)
    torch.manual_seed(11)
    ts, fs = specaug.draw_spec_augment(B2, T, M)
    x = logmel[:, None].clone().requires_grad_(True)
    bn.train()
    y = bn(x.transpose(1, 3)).transpose(1, 3)
    mask = torch.ones(B2, 1, T, M)
    for n in range(B2):
        for b, w in ts[n]:
            mask[n, :, b:b + w, :] = 0
        for b, w in fs[n]:
            mask[n, :, :, b:b + w] = 0
    y = y * mask
    if mixup:
        y = sed.mix_pairs(y, lam)
    dout = torch.randn(y.shape, generator=g)
    y.backward(dout)
    # device
    import copy
    bn_d = copy.deepcopy(bn).cuda()
    bn_d.running_mean.zero_(); bn_d.running_var.fill_(1.); bn_d.num_batches_tracked.zero_()
    lm = logmel.cuda()
    st = ops.bn_finalize(ops.colstats(lm.view(B2 * T, M)), B2 * T, bn_d)
    t_d = torch.as_tensor(ts, dtype=torch.int32).cuda()
    f_d = torch.as_tensor(fs, dtype=torch.int32).cuda()
    lam_d = lam.cuda() if mixup else None
    out = ops.bn0_aug_mix_fwd(lm, st, t_d, f_d, lam_d)
    assert torch.allclose(out.cpu(), y.detach()[:, 0], rtol=1e-4, atol=1e-4)
    assert torch.equal(out.cpu() == 0, y.detach()[:, 0] == 0) or not mixup      # stripes: exact zeros
    assert torch.allclose(bn_d.running_mean.cpu(), bn.running_mean, rtol=1e-5, atol=1e-5)
    assert torch.allclose(bn_d.running_var.cpu(), bn.running_var, rtol=1e-4)
    dg, db = torch.empty(M, device='cuda'), torch.empty(M, device='cuda')
    ops.bn0_bwd(dout[:, 0].contiguous().cuda(), lm, st, bn_d, t_d, f_d, lam_d, dg, db)
    assert torch.allclose(dg.cpu(), bn.weight.grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(db.cpu(), bn.bias.grad, rtol=1e-3, atol=1e-3)


def test_spec_augment_module_masks_the_gradient_like_upstream():
    """Seam A under autograd: upstream zeroes the stripes with a *tracked* in-place slice assignment,
    so the gradient inside the stripes is zero too.  The drop-in must behave the same when its input is part of
    an autograd graph (the bn0 output of the reference's own models.py:202-207)."""
    from oracle import frontend as ofe
    from sound_event_detection_dcase2017_task4_b200.dropin.torchlibrosa.augmentation import SpecAugmentation
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 1, 101, 64, generator=g)
    w = torch.randn(64, generator=g)
    dout = torch.randn(5, 1, 101, 64, generator=g)
    ref_mod = ofe.SpecAugmentation(time_drop_width=64, time_stripes_num=2, freq_drop_width=8, freq_stripes_num=2)
    mine = SpecAugmentation(time_drop_width=64, time_stripes_num=2, freq_drop_width=8, freq_stripes_num=2)
    ref_mod.train(); mine.train()
    # reference: leaf -> affine (stands for bn0) -> in-place stripes
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    torch.manual_seed(5)
    yr = ref_mod(xr * wr)
    yr.backward(dout)
    xd, wd = x.clone().cuda().requires_grad_(True), w.clone().cuda().requires_grad_(True)
    torch.manual_seed(5)
    yd = mine(xd * wd)
    yd.backward(dout.cuda())
    assert torch.equal(yd.detach().cpu() == 0, yr.detach() == 0)
    assert torch.allclose(yd.detach().cpu(), yr.detach(), rtol=1e-6, atol=1e-6)
    assert torch.equal(xd.grad.cpu() == 0, xr.grad == 0)                    # stripes carry no gradient
    assert torch.allclose(xd.grad.cpu(), xr.grad, rtol=1e-6, atol=1e-6)
    assert torch.allclose(wd.grad.cpu(), wr.grad, rtol=1e-4, atol=1e-4)
    # no graph: plain in-place kernel, same values
    torch.manual_seed(5)
    with torch.no_grad():
        y2 = mine((x * w).cuda())
    assert torch.allclose(y2.cpu(), yr.detach(), rtol=1e-6, atol=1e-6)
