"""-m gpu: BatchNorm2d(train) + ReLU + avg-pool kernels (csrc/bnpool.cu) through the C ABI against
PyTorch fp32 autograd on the same bf16-valued inputs (models.py:102-113 semantics)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # B, H, W, C, ph, pw, grad_f32
    (2, 101, 64, 64, 1, 1, False),     # block1.bn1 (no pool)
    (2, 101, 64, 64, 2, 2, False),     # block1.bn2 + 2x2 pool, odd H: floor mode drops the last row
    (3, 50, 32, 128, 2, 2, False),
    (2, 25, 16, 256, 2, 2, False),     # odd H again
    (2, 12, 8, 512, 1, 8, True),       # block4.bn2 + mean over the 8 mel bins, fp32 features
    (2, 50, 32, 128, 2, 2, True),      # stand-alone ConvBlock: fp32 gradient into a pooled layer
    # large enough for many dynamically scheduled workers in both size tiers (csrc/bnpool.cu make_bwd_plan)
    (6, 501, 64, 64, 1, 1, False),
    (6, 501, 64, 64, 2, 2, False),
    (24, 125, 8, 512, 1, 8, True),
    (5, 33, 24, 64, 3, 2, False),      # nothing the fast paths take: generic kernels
]


def _setup(B, H, W, C, seed=0):
    g = torch.Generator().manual_seed(seed)
    y = (torch.randn(B, H, W, C, generator=g) * 1.5 + 0.3).to(torch.bfloat16).cuda()
    bn = nn.BatchNorm2d(C).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    return y, bn, g


def _partial(y):
    yf = y.float().reshape(-1, y.shape[-1]).double()
    return torch.stack([yf.sum(0), (yf * yf).sum(0)]).float().unsqueeze(0).contiguous()


@pytest.mark.parametrize('B,H,W,C,ph,pw,gf32', CASES)
def test_bn_relu_pool_forward_backward(B, H, W, C, ph, pw, gf32):
    import copy
    from sound_event_detection_dcase2017_task4_b200 import ops
    y, bn, g = _setup(B, H, W, C)
    ref_bn = copy.deepcopy(bn)
    st = ops.bn_finalize(_partial(y), B * H * W, bn)
    out = ops.bn_relu_pool_fwd(y, st, ph, pw, out_f32=gf32)
    # reference: NCHW fp32 autograd
    y_ref = y.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref_bn.train()
    o_ref = F.avg_pool2d(torch.relu(ref_bn(y_ref)), kernel_size=(ph, pw))
    o_cmp = out.float().permute(0, 3, 1, 2)
    tol = 1e-5 if gf32 else 2.0 ** -8
    assert (o_cmp - o_ref).abs().max().item() <= tol * max(1.0, o_ref.abs().max().item())
    # running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance)
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, atol=1e-5, rtol=1e-5)
    assert torch.allclose(bn.running_var, ref_bn.running_var, atol=1e-5, rtol=1e-5)
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked) == 1
    # backward
    dA = torch.randn(o_ref.shape, generator=g).cuda()
    if not gf32:
        dA = dA.to(torch.bfloat16)
    o_ref.backward(dA.float())
    dgamma, dbeta = torch.empty(C, device='cuda'), torch.empty(C, device='cuda')
    dy = ops.bn_relu_pool_bwd(y, dA.permute(0, 2, 3, 1).contiguous(), st, bn, ph, pw, dgamma, dbeta)
    ref_dy = y_ref.grad.permute(0, 2, 3, 1)
    scale = ref_dy.abs().max().item()
    assert (dy.float() - ref_dy).abs().max().item() <= 2.0 ** -8 * scale + 1e-6
    assert torch.allclose(dgamma, ref_bn.weight.grad, rtol=2e-4, atol=2e-4 * ref_bn.weight.grad.abs().max().item())
    assert torch.allclose(dbeta, ref_bn.bias.grad, rtol=2e-4, atol=2e-4 * ref_bn.bias.grad.abs().max().item())


def test_bn_eval_affine_matches_eval_mode():
    from sound_event_detection_dcase2017_task4_b200 import ops
    y, bn, _ = _setup(2, 20, 16, 128)
    bn.eval()
    st = ops.bn_eval_affine(bn)
    out = ops.bn_relu_pool_fwd(y, st, 2, 2, out_f32=True)
    ref = F.avg_pool2d(torch.relu(bn(y.float().permute(0, 3, 1, 2))), 2)
    assert (out.permute(0, 3, 1, 2) - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_bn_backward_is_bit_reproducible_under_co_scheduling():
    """The backward kernels take their work from an atomic ticket counter, yet every virtual worker owns a fixed range
    and the finalize sums the per-worker rows in fixed order: the result may not depend on which CTA ran what -- not
    even when another kernel occupies part of the machine."""
    from sound_event_detection_dcase2017_task4_b200 import _lib, ops
    B, H, W, C = 8, 501, 64, 64
    y, bn, g = _setup(B, H, W, C, seed=3)
    assert _lib.lib().sed_bn_bwd_workspace_rows(B, H, W, C, 2, 2) > 148
    st = ops.bn_finalize(_partial(y), B * H * W, bn)
    dA = torch.randn(B, H // 2, W // 2, C, generator=g).to(torch.bfloat16).cuda()
    outs = []
    side = torch.cuda.Stream()
    junk = torch.randn(4096, 4096, device='cuda')
    for rep in range(4):
        dgamma, dbeta = torch.empty(C, device='cuda'), torch.empty(C, device='cuda')
        if rep % 2:
            with torch.cuda.stream(side):
                for _ in range(4):
                    junk @ junk                                  # competes for SMs while the BN kernels run
        dy = ops.bn_relu_pool_bwd(y, dA, st, bn, 2, 2, dgamma, dbeta)
        torch.cuda.synchronize()
        outs.append((dy.clone(), dgamma.clone(), dbeta.clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))
    assert int(ops.sched_words(y).abs().sum()) == 0              # the ticket words are left zeroed
