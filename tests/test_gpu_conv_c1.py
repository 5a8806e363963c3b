"""-m gpu: the Cin = 1 first convolution (csrc/conv_c1.cu; ConvBlock1.conv1, models.py:181/:102):
forward + BN statistics, weight gradient, data gradient vs PyTorch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,H,W', [(2, 101, 64), (3, 1001, 64), (1, 7, 64)])
def test_conv_c1_fwd_wgrad_dgrad(B, H, W):
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, W, generator=g).cuda()
    w = (torch.randn(64, 1, 3, 3, generator=g) * 0.3).cuda()
    y, partial = ops.conv_c1_fwd(x, w, want_stats=True)
    ref = F.conv2d(x[:, None], w, padding=1).permute(0, 2, 3, 1)            # (B,H,W,64)
    assert (y.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item()
    s = partial.double().sum(0)
    n = B * H * W
    assert (s[0] / n - ref.double().mean((0, 1, 2))).abs().max().item() <= 1e-4
    assert ((s[1] / n - (ref.double() ** 2).mean((0, 1, 2))).abs() / (ref.double() ** 2).mean((0, 1, 2))).max().item() <= 1e-4
    dy = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).cuda()
    xr = x[:, None].clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    gw = torch.empty_like(w)
    ops.conv_c1_wgrad(x, dy, gw)
    assert (gw - wr.grad).abs().max().item() <= 1e-4 * wr.grad.abs().max().item() + 1e-4
    dx = ops.conv_c1_dgrad(dy, w)
    assert (dx - xr.grad[:, 0]).abs().max().item() <= 1e-4 * xr.grad.abs().max().item() + 1e-5


@pytest.mark.parametrize('B,H,W', [(2, 101, 64), (3, 1001, 64), (1, 7, 64), (5, 33, 32)])
def test_bn_apply_fused_into_c1_wgrad_equals_two_passes(B, H, W):
    """sed_bn_apply_conv_c1_wgrad: block1.bn1's backward apply pass inside the Cin = 1 weight-gradient kernel.  dY and dW
    must equal the two separate passes bit for bit (same arithmetic, same bf16 rounding, same accumulation order), for
    full and ragged 16-row items and image rows beyond H inside a 128-pixel block."""
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, generator=g).cuda()
    y = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).cuda()
    dA = (torch.randn(B, H, W, 64, generator=g) * 0.1).to(torch.bfloat16).cuda()
    bn = torch.nn.BatchNorm2d(64).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(64, generator=g) + 0.5)
        bn.weight[::5] *= -1.0
        bn.bias.copy_(torch.randn(64, generator=g) * 0.3)
    stats = torch.stack([y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))])[None]      # (1, 2, 64) partial
    st = ops.bn_finalize(stats.contiguous(), B * H * W, bn)
    dgam, dbet = torch.empty(64, device='cuda'), torch.empty(64, device='cuda')
    coef = ops.bn_bwd_coef(y, dA, st, bn, 1, 1, dgam, dbet)
    dy_ref = ops.bn_bwd_apply(y, dA, st, coef, 1, 1)
    gw_ref = torch.empty(64, 1, 3, 3, device='cuda')
    ops.conv_c1_wgrad(x, dy_ref, gw_ref)
    gw = torch.full((64, 1, 3, 3), 9.0, device='cuda')
    dy = ops.bn_apply_conv_c1_wgrad(x, y, dA, st, coef, gw)
    assert torch.equal(dy, dy_ref)
    assert torch.equal(gw, gw_ref)
