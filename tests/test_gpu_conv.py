"""-m gpu: tcgen05 implicit-GEMM 3x3 convolution (fwd, dgrad, wgrad, BN statistics) through the
C ABI against a plain PyTorch fp32 reference evaluated on the same bf16-rounded operands.
Tolerance: the kernel accumulates in fp32 and rounds once to bf16 (|err| <= 2^-8 * max|y|)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_EPS = 2.0 ** -8          # one rounding to bf16: half an ulp is 2^-9 relative to the binade top

SHAPES = [  # (B, H, W, Cin, Cout)
    (2, 101, 64, 64, 64),      # block1.conv2 on a 1 s clip (H tail: 101 = 50*2 + 1)
    (3, 50, 32, 64, 128),      # block2.conv1
    (2, 50, 32, 128, 128),     # block2.conv2
    (2, 25, 16, 128, 256),     # block3.conv1 (H tail 25 = 3*8 + 1)
    (2, 25, 16, 256, 256),     # block3.conv2
    (2, 12, 8, 256, 512),      # block4.conv1 (two N tiles, H tail)
    (1, 125, 8, 512, 512),     # block4.conv2 at the full 10 s length
    (150, 4, 32, 64, 128),     # more tiles than SMs: persistent loop + accumulator double buffering
]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


@pytest.mark.parametrize('B,H,W,Cin,Cout', SHAPES)
def test_conv_fwd_and_stats(B, H, W, Cin, Cout):
    from sound_event_detection_dcase2017_task4_b200 import conv
    x = _rand((B, H, W, Cin), 1).cuda().to(torch.bfloat16)
    w = (_rand((Cout, Cin, 3, 3), 2) * (2.0 / (9 * Cin)) ** 0.5).cuda()
    wf, _ = conv.pack_weights(w)
    y, stats = conv.conv3x3(x, wf, Cout, want_stats=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    err = (y.float() - ref).abs().max().item()
    assert err <= BF16_EPS * ref.abs().max().item() + 1e-3, err
    s = stats.double().sum(dim=0)                 # (2, Cout)
    n = B * H * W
    mean_ref = ref.double().mean(dim=(0, 1, 2))
    sq_ref = (ref.double() ** 2).mean(dim=(0, 1, 2))
    e0 = (s[0] / n - mean_ref).abs().max().item()
    e1 = ((s[1] / n - sq_ref).abs() / sq_ref).max().item()
    assert e0 <= 2e-4 and e1 <= 2e-4, (e0, e1)


@pytest.mark.parametrize('B,H,W,Cin,Cout', SHAPES)
def test_conv_dgrad(B, H, W, Cin, Cout):
    from sound_event_detection_dcase2017_task4_b200 import conv
    dy = _rand((B, H, W, Cout), 3).cuda().to(torch.bfloat16)
    w = (_rand((Cout, Cin, 3, 3), 4) * (2.0 / (9 * Cout)) ** 0.5).cuda()
    _, wd = conv.pack_weights(w)
    dx = conv.conv3x3(dy, wd, Cin)
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    err = (dx.float() - ref).abs().max().item()
    assert err <= BF16_EPS * ref.abs().max().item() + 1e-3, err


@pytest.mark.parametrize('B,H,W,Cin,Cout', SHAPES)
def test_conv_wgrad(B, H, W, Cin, Cout):
    from sound_event_detection_dcase2017_task4_b200 import conv
    x = _rand((B, H, W, Cin), 5).cuda().to(torch.bfloat16)
    dy = _rand((B, H, W, Cout), 6).cuda().to(torch.bfloat16)
    dw = conv.conv3x3_wgrad(dy, x)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wr = torch.zeros((Cout, Cin, 3, 3), device='cuda', requires_grad=True)
    F.conv2d(xr, wr, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    ref = wr.grad
    err = (dw - ref).abs().max().item()
    assert err <= 1e-3 * ref.abs().max().item() + 1e-3, err


@pytest.mark.parametrize('B,H,W,Cin,Cout,pool', [
    (2, 101, 64, 64, 64, 1),      # block1.conv2 dgrad + bn1 reduce (resident-weight kernel)
    (3, 50, 32, 128, 64, 2),      # block2.conv1 dgrad + block1.bn2 reduce through the 2x2 pool (y has 101 rows: tail)
    (2, 50, 32, 128, 128, 1),     # block2.conv2 dgrad + bn1 reduce
    (2, 25, 16, 256, 128, 2),     # block3.conv1 dgrad + block2.bn2 reduce (y has 50 rows)
    (2, 12, 8, 512, 256, 2),      # block4.conv1 dgrad + block3.bn2 reduce (N = 256 kernel, y has 25 rows: tail)
    (1, 125, 8, 512, 512, 1),     # block4.conv2 dgrad + bn1 reduce
])
def test_conv_dgrad_with_fused_bn_reduce(B, H, W, Cin, Cout, pool):
    """sed_conv3x3_tc_dgrad_bnr: same dX as the plain data gradient, and partial sums of g and g*y for the layer
    below (g = unpool(dX)/pool^2 * relu-mask) equal to a PyTorch fp32 evaluation on the stored bf16 dX."""
    from sound_event_detection_dcase2017_task4_b200 import conv
    cin_eff = Cout                                    # channels of the gradient being produced = channels of y below
    dy = _rand((B, H, W, Cin), 11).cuda().to(torch.bfloat16)
    w = (_rand((Cin, cin_eff, 3, 3), 12) * (2.0 / (9 * Cin)) ** 0.5).cuda()      # layer weight (Cout=Cin here, Cin=cin_eff)
    _, wd = conv.pack_weights(w)
    Hy = H * pool + (1 if pool == 2 and H % 2 == 0 else 0)
    y = _rand((B, Hy, W * pool, cin_eff), 13).cuda().to(torch.bfloat16)

    class St(object):
        pass
    st = St()
    st.scale = (0.5 + torch.rand(cin_eff, device='cuda'))
    st.scale[::3] *= -1.0
    st.shift = torch.randn(cin_eff, device='cuda') * 0.3
    was = conv.USE_KWSTACK
    try:
        conv.USE_KWSTACK = False                      # same accumulation order as the fused kernel: bit-equal dX
        dx_plain = conv.conv3x3(dy, wd, cin_eff)
    finally:
        conv.USE_KWSTACK = was
    dx, partial = conv.conv3x3_dgrad_bnr(dy, wd, cin_eff, y, st, pool)
    assert torch.equal(dx, dx_plain)
    g = dx.float()
    if pool == 2:
        g = g.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2) * 0.25
    yy = y.float()[:, :H * pool]
    g = g * ((yy * st.scale + st.shift) > 0)
    want = torch.stack([g.double().sum(dim=(0, 1, 2)), (g.double() * yy.double()).sum(dim=(0, 1, 2))])
    got = partial.double().sum(dim=0)
    scale = want.abs().max(dim=1, keepdim=True).values + 1e-6
    assert ((got - want).abs() / scale).max().item() <= 2e-4


@pytest.mark.parametrize('B,H,W,Cin,Cout', SHAPES + [
    (3, 50, 32, 128, 64),      # block2.conv1 data gradient (N = 64 with streamed weights)
    (2, 25, 16, 256, 128),     # block3.conv1 data gradient
    (150, 12, 8, 256, 256),    # more pair tiles than CTA pairs: persistent loop + accumulator double buffering
    (1, 3, 64, 64, 64),        # a single pair tile, mostly tail rows
])
def test_conv_cta_pair_kernel(B, H, W, Cin, Cout):
    """tcgen05.mma.cta_group::2 variant (sed_conv3x3_tc2_fwd): same outputs as the single-CTA kernel (identical
    accumulation order per output element => bit-equal), statistics equal to a PyTorch fp32 evaluation."""
    from sound_event_detection_dcase2017_task4_b200 import conv
    x = _rand((B, H, W, Cin), 21).cuda().to(torch.bfloat16)
    w = (_rand((Cout, Cin, 3, 3), 22) * (2.0 / (9 * Cin)) ** 0.5).cuda()
    wf, _ = conv.pack_weights(w)
    was = conv.USE_2CTA, conv.USE_KWSTACK
    try:
        conv.USE_KWSTACK = False              # (the kw-stacked Cout = 64 kernel sums in another order: its own test below)
        conv.USE_2CTA = False
        y1, s1 = conv.conv3x3(x, wf, Cout, want_stats=True)
        conv.USE_2CTA = True
        y2, s2 = conv.conv3x3(x, wf, Cout, want_stats=True)
        torch.cuda.synchronize()
    finally:
        conv.USE_2CTA, conv.USE_KWSTACK = was
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    err = (y2.float() - ref).abs().max().item()
    assert err <= BF16_EPS * ref.abs().max().item() + 1e-3, err
    assert torch.equal(y1, y2)
    a, b = s1.double().sum(0), s2.double().sum(0)
    assert ((a - b).abs() / (a.abs() + 1.0)).max().item() <= 1e-5


@pytest.mark.parametrize('B,H,W,Cin', [
    (2, 101, 64, 64),      # block1.conv2 / its data gradient: an image row spans two warps (mailbox exchange), H tail
    (1, 3, 64, 64),        # a single pair tile, mostly tail rows
    (3, 50, 32, 128),      # block2.conv1 data gradient: a warp is one image row, two 64-channel K blocks
    (2, 25, 16, 64),       # two image rows per warp
    (2, 13, 8, 128),       # four image rows per warp
    (1, 7, 128, 64),       # one image row spans all four warps
    (150, 4, 32, 64),      # more pair tiles than CTA pairs: persistent loop + accumulator double buffering
    (64, 1000, 64, 64),    # the bench layer at a quarter of the batch
])
def test_conv_kw_stacked_kernel(B, H, W, Cin):
    """Cout = 64 kernel with the three kw taps stacked in N (sed_conv3x3_tc2kw_fwd): the kw shift is applied between
    accumulator rows in the epilogue (shuffles + a shared-memory mailbox across warp boundaries).  Against PyTorch fp32 on
    the same bf16 operands, against the unstacked CTA-pair kernel (same values up to the order of three fp32 additions,
    i.e. at most one bf16 ulp on a few outputs) and for its BatchNorm statistics."""
    from sound_event_detection_dcase2017_task4_b200 import _lib, conv
    Cout = 64
    assert _lib.lib().sed_conv3x3_tc2kw_supported(W, Cin, Cout) == 1
    x = _rand((B, H, W, Cin), 41).cuda().to(torch.bfloat16)
    w = (_rand((Cout, Cin, 3, 3), 42) * (2.0 / (9 * Cin)) ** 0.5).cuda()
    wf, _ = conv.pack_weights(w)
    was = conv.USE_KWSTACK, conv.KWSTACK_MIN_CIN
    try:
        conv.USE_KWSTACK, conv.KWSTACK_MIN_CIN = False, 64          # (the product uses it from 128 input channels on)
        y1, s1 = conv.conv3x3(x, wf, Cout, want_stats=True)
        conv.USE_KWSTACK = True
        y2, s2 = conv.conv3x3(x, wf, Cout, want_stats=True)
        y3, s3 = conv.conv3x3(x, wf, Cout, want_stats=True)
        torch.cuda.synchronize()
    finally:
        conv.USE_KWSTACK, conv.KWSTACK_MIN_CIN = was
    assert torch.equal(y2, y3) and torch.equal(s2, s3)                   # deterministic
    d = (y1.float() - y2.float()).abs()
    assert d.max().item() <= BF16_EPS * 2 * y1.float().abs().max().item()
    assert (d > 0).float().mean().item() <= 0.02                         # a rounding flip here and there, nothing else
    if B * H <= 4000:
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
        err = (y2.float() - ref).abs().max().item()
        assert err <= BF16_EPS * ref.abs().max().item() + 1e-3, err
    a, b = s1.double().sum(0), s2.double().sum(0)          # sums of fp32 accumulators that differ in their last bits
    n = B * H * W
    assert ((a[0] - b[0]).abs() / n).max().item() <= 1e-5
    assert ((a[1] - b[1]).abs() / a[1]).max().item() <= 1e-4


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(2, 25, 16, 128, 256), (2, 25, 16, 256, 256), (2, 12, 8, 256, 512),
                                            (1, 125, 8, 512, 512), (40, 25, 16, 256, 256)])
def test_conv_wgrad_cta_pair_kernel(B, H, W, Cin, Cout):
    """cta_group::2 weight-gradient kernel vs the single-CTA one (same accumulation order => bit-equal)."""
    from sound_event_detection_dcase2017_task4_b200 import _lib, conv
    x = _rand((B, H, W, Cin), 31).cuda().to(torch.bfloat16)
    dy = _rand((B, H, W, Cout), 32).cuda().to(torch.bfloat16)
    was = _lib.lib().sed_conv3x3_tc_wgrad_use_pairs(0)
    try:
        single = conv.conv3x3_wgrad(dy, x)
        _lib.lib().sed_conv3x3_tc_wgrad_use_pairs(1)
        pair = conv.conv3x3_wgrad(dy, x)
        torch.cuda.synchronize()
    finally:
        _lib.lib().sed_conv3x3_tc_wgrad_use_pairs(was)
    assert torch.equal(single, pair)


def test_multi_layer_pack_and_unpack_equal_per_layer_entry_points():
    """sed_conv_pack_weights_multi / sed_conv_unpack_wgrad_multi (one launch for all layers of the model) produce the
    same bytes as the per-layer entry points."""
    from sound_event_detection_dcase2017_task4_b200 import conv
    shapes = [(64, 64), (128, 64), (128, 128), (256, 128), (256, 256), (512, 256), (512, 512)]
    ws = [(_rand((co, ci, 3, 3), 50 + i) * 0.05).cuda() for i, (co, ci) in enumerate(shapes)]
    multi = conv.pack_weights_multi(ws, want_dgrad=True)
    for w, (f, d) in zip(ws, multi):
        f1, d1 = conv.pack_weights(w)
        assert torch.equal(f, f1) and torch.equal(d, d1)
    fwd_only = conv.pack_weights_multi(ws[:3], want_dgrad=False)
    assert all(d is None for _, d in fwd_only) and torch.equal(fwd_only[2][0], multi[2][0])
    items, refs = [], []
    for i, (B, H, W, Cin, Cout) in enumerate([(2, 25, 16, 128, 256), (3, 50, 32, 64, 128), (2, 101, 64, 64, 64),
                                               (1, 125, 8, 512, 512)]):
        x = _rand((B, H, W, Cin), 60 + i).cuda().to(torch.bfloat16)
        dy = _rand((B, H, W, Cout), 70 + i).cuda().to(torch.bfloat16)
        refs.append(conv.conv3x3_wgrad(dy, x))
        items.append((conv.conv3x3_wgrad_slabs(dy, x), torch.full((Cout, Cin, 3, 3), 7.0, device='cuda')))
    conv.unpack_wgrad_multi(items)
    for (_, out), ref in zip(items, refs):
        assert torch.equal(out, ref)
