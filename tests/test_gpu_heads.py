"""-m gpu: head / loss / optimizer kernels (csrc/heads.cu, csrc/adam.cu, csrc/prep.cu mix_pairs)
through the C ABI against PyTorch fp32 (models.py:118-149, :221-227, :306-312; losses.py:5-12;
main.py:144-145)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def test_linear_small_fwd_bwd():
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = _g(0)
    R, C, K = 24 * 125, 512, 17
    x = torch.randn(R, C, generator=g).cuda().requires_grad_(True)
    W = (torch.randn(K, C, generator=g) * 0.05).cuda().requires_grad_(True)
    b = torch.randn(K, generator=g).cuda().requires_grad_(True)
    out = ops.linear_small_fwd(x.detach(), W.detach(), b.detach())
    ref = F.linear(x, W, b)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    dout = torch.randn(R, K, generator=g).cuda()
    ref.backward(dout)
    dW, db = torch.empty_like(W), torch.empty_like(b)
    dx = ops.linear_small_bwd(dout, x.detach(), W.detach(), dW, db)
    assert torch.allclose(dx, x.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(dW, W.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(db, b.grad, rtol=1e-4, atol=1e-4)
    dx2 = ops.linear_small_bwd(dout, x.detach(), W.detach(), None, None, dx=dx.clone())   # accumulate
    assert torch.allclose(dx2, 2 * x.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('R,C,K1,K2', [(24 * 125, 512, 17, 17), (32000, 512, 17, 17), (77, 96, 5, 10), (33, 512, 20, 20),
                                       (64, 512, 30, 30)])
def test_linear_pair_fwd(R, C, K1, K2):
    """AttBlock's att and cla maps in one pass over the features (sed_linear_pair_fwd): equal to the two single maps
    and to F.linear; ragged row / column tiles; K1 + K2 > 40 takes the two-launch path."""
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = _g(11)
    x = torch.randn(R, C, generator=g).cuda()
    W1, W2 = (torch.randn(K1, C, generator=g) * 0.05).cuda(), (torch.randn(K2, C, generator=g) * 0.05).cuda()
    b1, b2 = torch.randn(K1, generator=g).cuda(), torch.randn(K2, generator=g).cuda()
    o1, o2 = ops.linear_pair_fwd(x, W1, b1, W2, b2)
    assert torch.allclose(o1, F.linear(x, W1, b1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(o2, F.linear(x, W2, b2), rtol=1e-5, atol=1e-5)
    # same ascending-c fp32 FMA chain per output as the single-map kernel
    assert torch.equal(o1, ops.linear_small_fwd(x, W1, b1)) and torch.equal(o2, ops.linear_small_fwd(x, W2, b2))
    o1n, _ = ops.linear_pair_fwd(x, W1, None, W2, None)
    assert torch.allclose(o1n, F.linear(x, W1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('mode', [0, 1])
def test_head_pool_fwd_bwd(mode):
    """FrameAvg (mean) / FrameMax (max) heads incl. the x8 interpolation (exact copies)."""
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = _g(1)
    B, T, K, ratio = 5, 125, 17, 8
    logit = (torch.randn(B, T, K, generator=g) * 2).cuda().requires_grad_(True)
    prob, clip, argmax, frame = ops.head_pool_fwd(logit.detach(), ratio, mode)
    p_ref = torch.sigmoid(logit)
    f_ref = p_ref[:, :, None, :].expand(B, T, ratio, K).reshape(B, T * ratio, K)
    c_ref = f_ref.mean(1) if mode == 0 else f_ref.max(1)[0]
    assert torch.allclose(prob, p_ref, rtol=1e-6, atol=1e-7)
    assert torch.equal(frame, prob[:, :, None, :].expand(B, T, ratio, K).reshape(B, T * ratio, K))  # bit-exact index map
    assert torch.allclose(clip, c_ref, rtol=1e-5, atol=1e-6)
    dclip = torch.randn(B, K, generator=g).cuda()
    c_ref.backward(dclip)
    dlogit = ops.head_pool_bwd(prob, dclip, argmax, mode)
    assert torch.allclose(dlogit, logit.grad, rtol=1e-4, atol=1e-7)


def test_head_att_fwd_bwd():
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = _g(2)
    B, T, K = 4, 125, 17
    att = (torch.randn(B, T, K, generator=g) * 6).cuda().requires_grad_(True)      # exercises the +-10 clamp
    cla = (torch.randn(B, T, K, generator=g) * 2).cuda().requires_grad_(True)
    clip, norm_att, cla_p, frame = ops.head_att_fwd(att.detach(), cla.detach(), 8, True, 1.0)
    e = torch.exp(torch.clamp(att.transpose(1, 2), -10, 10)) + 1e-6
    na = e / e.sum(dim=2, keepdim=True)
    cp = torch.sigmoid(cla.transpose(1, 2))
    c_ref = (na * cp).sum(dim=2)
    assert torch.allclose(norm_att, na, rtol=1e-5, atol=1e-8)
    assert torch.allclose(cla_p, cp, rtol=1e-6, atol=1e-7)
    assert torch.allclose(clip, c_ref, rtol=1e-5, atol=1e-6)
    assert torch.equal(frame, cla_p.transpose(1, 2)[:, :, None, :].expand(B, T, 8, K).reshape(B, T * 8, K))
    dclip = torch.randn(B, K, generator=g).cuda()
    c_ref.backward(dclip)
    d_att, d_cla = ops.head_att_bwd(att.detach(), norm_att, cla_p, clip, dclip, True, 1.0)
    assert torch.allclose(d_att, att.grad, rtol=1e-3, atol=1e-7)
    assert torch.allclose(d_cla, cla.grad, rtol=1e-4, atol=1e-8)


def test_bce_matches_torch_including_log_clamp():
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = _g(3)
    p = torch.rand(64, 17, generator=g)
    p[0, 0], p[0, 1], p[1, 0] = 0.0, 1.0, 1e-30                       # log clamped at -100
    t = (torch.rand(64, 17, generator=g) < 0.1).float() * torch.rand(64, 17, generator=g)   # soft targets
    pc = p.cuda().requires_grad_(True)
    ref = F.binary_cross_entropy(pc, t.cuda())
    loss, dprob = ops.bce(pc.detach(), t.cuda(), want_grad=True)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    ref.backward()
    inner = (p > 1e-6) & (p < 1 - 1e-6)
    assert torch.allclose(dprob.cpu()[inner], pc.grad.cpu()[inner], rtol=1e-4, atol=1e-7)


def test_mix_pairs_is_do_mixup():
    from sound_event_detection_dcase2017_task4_b200 import ops, pytorch_utils
    g = _g(4)
    x = torch.rand(16, 17, generator=g).cuda()
    lam = torch.rand(16, generator=g).cuda()
    assert torch.equal(ops.mix_pairs(x, lam), pytorch_utils.do_mixup(x, lam))


def test_adam_amsgrad_matches_torch_optim():
    from sound_event_detection_dcase2017_task4_b200 import ops
    g = _g(5)
    n = 100003
    p0 = torch.randn(n, generator=g).cuda()
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0., amsgrad=True)
    p = p0.clone()
    m, v, vmax = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        grad = (torch.randn(n, generator=g) * (0.1 if step % 2 else 3.0)).cuda()
        ref_p.grad = grad.clone()
        opt.step()
        ops.adam_amsgrad_(p, grad, m, v, vmax, 1e-3, 0.9, 0.999, 1e-8, step)
        assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=1e-6), step
