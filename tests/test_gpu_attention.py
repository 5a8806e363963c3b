"""-m gpu: MultiHead self-attention (csrc/attention.cu + tensor-core projections) against the
oracle's MultiHead (restating models.py:587-665), forward and backward, and the dropout kernels."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _modules(seed=0):
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import models
    torch.manual_seed(seed)
    ref = sed.MultiHead(8, 512, 64, 64, 0.2)
    torch.manual_seed(seed)
    mine = models.MultiHead(8, 512, 64, 64, 0.2)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    with torch.no_grad():                                   # non-zero biases so their gradients are exercised
        for m in (ref, mine):
            g = torch.Generator().manual_seed(9)
            for lin in (m.w_qs, m.w_ks, m.w_vs, m.fc):
                lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.1)
    return ref, mine.cuda()


@pytest.mark.parametrize('B,T', [(3, 125), (2, 12), (1, 128)])
def test_multihead_forward_backward_no_dropout(B, T):
    ref, mine = _modules()
    for m in (ref, mine):
        m.train()
        m.dropout.p = 0.0
        m.attention.dropout.p = 0.0
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, 512, generator=g)
    xr = x.clone().requires_grad_(True)
    xm = x.clone().cuda().requires_grad_(True)
    y_ref = ref(xr, xr, xr)
    y = mine(xm, xm, xm)
    assert y.shape == y_ref.shape
    assert (y.detach().cpu() - y_ref.detach()).abs().max().item() <= 1e-4 * y_ref.abs().max().item() + 1e-5
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    y.backward(dy.cuda())
    # backward GEMMs run on plain bf16 operands (fp32 accumulate): 2e-2 relative L2
    def rel(a, b):
        return (a.cpu().double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)
    assert rel(xm.grad, xr.grad) <= 2e-2
    for (k, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        if q.grad is None:
            assert p.grad is None, k                         # layer_norm: registered, never used
            continue
        if k == 'w_ks.bias':
            continue      # adds the same q.b_k to every key of a row: softmax cancels it, the true gradient is 0
        assert rel(p.grad, q.grad) <= 2e-2, k


def test_attention_kernel_exact_fp32_parts():
    """sed_attention_fwd / _bwd alone (fp32 CUDA-core math) vs torch autograd at fp32 tolerance."""
    from sound_event_detection_dcase2017_task4_b200._lib import call, stream_of
    B, T, H, d = 2, 125, 8, 64
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(B * T, 3 * H * d, generator=g).cuda()
    ref_in = qkv.clone().requires_grad_(True)
    q, k, v = [ref_in[:, i * H * d:(i + 1) * H * d].view(B, T, H, d).permute(0, 2, 1, 3) for i in range(3)]
    att = torch.softmax(q @ k.transpose(2, 3) / 8.0, dim=3)
    ctx_ref = (att @ v).permute(0, 2, 1, 3).reshape(B * T, H * d)
    ctx = torch.empty(B * T, H * d, device='cuda')
    probs = torch.empty(B, H, T, T, device='cuda')
    base, ld = qkv.data_ptr(), 3 * H * d
    call('sed_attention_fwd', base, base + 4 * H * d, base + 8 * H * d, ld, ld, ld, B, T, H, d, 8.0, 0.0, 0, 0, 0,
         ctx.data_ptr(), probs.data_ptr(), stream_of(qkv))
    assert torch.allclose(ctx, ctx_ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(probs, att, rtol=1e-4, atol=1e-7)
    dctx = torch.randn(B * T, H * d, generator=g).cuda()
    ctx_ref.backward(dctx)
    dqkv = torch.empty_like(qkv)
    gb = dqkv.data_ptr()
    call('sed_attention_bwd', base, base + 4 * H * d, base + 8 * H * d, ld, ld, ld, B, T, H, d, 8.0, 0.0, 0, 0, 0,
         dctx.data_ptr(), probs.data_ptr(), gb, gb + 4 * H * d, gb + 8 * H * d, stream_of(qkv))
    assert torch.allclose(dqkv, ref_in.grad, rtol=1e-3, atol=1e-5)


def test_dropout_statistics_and_determinism():
    """Dropout(0.1) on the attention and Dropout(0.2)+ReLU on the output: keep rates, 1/(1-p) scaling,
    torch.cuda.manual_seed determinism, eval mode = identity."""
    from sound_event_detection_dcase2017_task4_b200._lib import call, stream_of
    n = 1 << 20
    x = torch.ones(n, device='cuda')
    y = torch.empty_like(x)
    call('sed_dropout_relu_fwd', x.data_ptr(), n, 0.2, 1234, 0, 0, y.data_ptr(), stream_of(x))
    kept = (y > 0).float().mean().item()
    assert abs(kept - 0.8) < 3e-3
    assert torch.all((y == 0) | ((y - 1.25).abs() < 1e-6))
    y2 = torch.empty_like(x)
    call('sed_dropout_relu_fwd', x.data_ptr(), n, 0.2, 1234, 0, 0, y2.data_ptr(), stream_of(x))
    assert torch.equal(y, y2)
    call('sed_dropout_relu_fwd', x.data_ptr(), n, 0.2, 1235, 0, 0, y2.data_ptr(), stream_of(x))
    assert not torch.equal(y, y2)
    # generator state through device memory (CUDA-graph replay): {seed, offset base} words + relative offset
    state = torch.tensor([1234, 5], dtype=torch.int64, device='cuda')
    y3, y4 = torch.empty_like(x), torch.empty_like(x)
    call('sed_dropout_relu_fwd', x.data_ptr(), n, 0.2, 999, 3, state.data_ptr(), y3.data_ptr(), stream_of(x))
    call('sed_dropout_relu_fwd', x.data_ptr(), n, 0.2, 1234, 8, 0, y4.data_ptr(), stream_of(x))
    assert torch.equal(y3, y4)
    dx = torch.empty_like(x)
    call('sed_dropout_relu_bwd', x.data_ptr(), y.data_ptr(), n, 0.2, dx.data_ptr(), stream_of(x))
    assert torch.equal(dx, y)                                 # dy = 1: dx = mask / (1 - p) = y
    ref, mine = _modules()
    xm = torch.randn(2, 125, 512, generator=torch.Generator().manual_seed(3)).cuda()
    mine.train()
    torch.cuda.manual_seed(7)
    a = mine(xm, xm, xm)
    torch.cuda.manual_seed(7)
    b = mine(xm, xm, xm)
    c = mine(xm, xm, xm)
    assert torch.equal(a, b) and not torch.equal(a, c)
    mine.eval(); ref.eval()
    with torch.no_grad():
        e = mine(xm, xm, xm)
        e_ref = ref(xm.cpu(), xm.cpu(), xm.cpu())
    assert (e.cpu() - e_ref).abs().max().item() <= 1e-4 * e_ref.abs().max().item() + 1e-5
