"""Launch wrappers for the tensor-core 3x3 convolution family (csrc/conv_halo2_tc.cu, csrc/conv_halo_tc.cu,
csrc/conv_wgrad_tc.cu, csrc/pack.cu).  Tensors are NHWC bf16 (B, H, W, C)."""
import torch

from . import _lib


def tc_supported(W, Cin, Cout):
    return (W >= 8 and W <= 128 and 128 % W == 0 and Cin >= 64 and Cin % 64 == 0 and
            (Cin == 64 or Cin % 128 == 0) and (Cout in (64, 128, 256, 512)))


def pack_weights(w_oihw, want_fwd=True, want_dgrad=True):
    """fp32 (Cout, Cin, 3, 3) -> (fwd bf16 [Cout, 9*Cin], dgrad bf16 [Cin, 9*Cout])."""
    assert w_oihw.is_cuda and w_oihw.dtype == torch.float32 and tuple(w_oihw.shape[2:]) == (3, 3)
    w = w_oihw.detach().contiguous()
    cout, cin = w.shape[:2]
    fwd = torch.empty((cout, 9 * cin), dtype=torch.bfloat16, device=w.device) if want_fwd else None
    dg = torch.empty((cin, 9 * cout), dtype=torch.bfloat16, device=w.device) if want_dgrad else None
    with torch.cuda.device(w.device):
        _lib.call('sed_conv_pack_weights', w.data_ptr(), cout, cin, _lib.ptr(fwd), _lib.ptr(dg),
                  _lib.stream_of(w))
    return fwd, dg


def _c_array(ctype, values):
    import ctypes
    arr = (ctype * len(values))(*values)
    return arr, ctypes.addressof(arr)


def pack_weights_multi(ws, want_dgrad=True):
    """[fp32 (Cout, Cin, 3, 3)] -> [(fwd bf16 [Cout, 9*Cin], dgrad bf16 [Cin, 9*Cout] or None)], ONE launch for all layers
    (sed_conv_pack_weights_multi; same values as pack_weights per layer)."""
    import ctypes
    if not ws:
        return []
    dev = ws[0].device
    ws = [w.detach().contiguous() for w in ws]
    sizes = [w.numel() for w in ws]
    pool = torch.empty(sum(sizes) * (2 if want_dgrad else 1), dtype=torch.bfloat16, device=dev)
    packs, off = [], 0
    for w, n in zip(ws, sizes):
        cout, cin = w.shape[:2]
        fwd = pool[off:off + n].view(cout, 9 * cin)
        off += n
        dg = None
        if want_dgrad:
            dg = pool[off:off + n].view(cin, 9 * cout)
            off += n
        packs.append((fwd, dg))
    keep = [_c_array(ctypes.c_void_p, [w.data_ptr() for w in ws]),
            _c_array(ctypes.c_int, [w.shape[0] for w in ws]), _c_array(ctypes.c_int, [w.shape[1] for w in ws]),
            _c_array(ctypes.c_void_p, [f.data_ptr() for f, _ in packs]),
            _c_array(ctypes.c_void_p, [0 if d is None else d.data_ptr() for _, d in packs])]
    with torch.cuda.device(dev):
        _lib.call('sed_conv_pack_weights_multi', len(ws), keep[0][1], keep[1][1], keep[2][1], keep[3][1], keep[4][1],
                  _lib.stream_of(ws[0]))
    return packs


def conv3x3_wgrad_slabs(dy, x):
    """The split-K slabs (splits, 9, Cout, Cin) fp32 of dW = dy^T x, not yet folded (see unpack_wgrad_multi)."""
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    assert dy.is_contiguous() and x.is_contiguous() and dy.shape[:3] == x.shape[:3]
    b, h, w, cin = x.shape
    cout = dy.shape[3]
    with torch.cuda.device(x.device):
        splits = _lib.lib().sed_conv3x3_tc_wgrad_splits(b, h, w, cin, cout)
        slabs = torch.empty((splits, 9, cout, cin), dtype=torch.float32, device=x.device)
        _lib.call('sed_conv3x3_tc_wgrad', dy.data_ptr(), x.data_ptr(), slabs.data_ptr(), b, h, w, cin,
                  cout, _lib.stream_of(x))
    return slabs


def unpack_wgrad_multi(items):
    """[(slabs (S, 9, Cout, Cin) fp32, out (Cout, Cin, 3, 3) fp32)]: fold the split-K slabs of up to 8 layers into their
    OIHW gradients with ONE launch (sed_conv_unpack_wgrad_multi), on the current stream."""
    import ctypes
    if not items:
        return
    dev = items[0][0].device
    keep = [_c_array(ctypes.c_void_p, [s.data_ptr() for s, _ in items]),
            _c_array(ctypes.c_int, [s.shape[0] for s, _ in items]),
            _c_array(ctypes.c_int, [s.shape[2] for s, _ in items]), _c_array(ctypes.c_int, [s.shape[3] for s, _ in items]),
            _c_array(ctypes.c_void_p, [o.data_ptr() for _, o in items])]
    for s, o in items:
        assert o.is_contiguous() and o.dtype == torch.float32 and o.numel() == 9 * s.shape[2] * s.shape[3]
    with torch.cuda.device(dev):
        _lib.call('sed_conv_unpack_wgrad_multi', len(items), keep[0][1], keep[1][1], keep[2][1], keep[3][1], keep[4][1],
                  _lib.stream_of(items[0][0]))


# CTA-pair (tcgen05.mma.cta_group::2) kernel: measured 33.9 vs 35.7 ms/step (tools/ab_step.py conv.USE_2CTA=0,1)
USE_2CTA = True
# Cout = 64: the three kw taps stacked in N (csrc/conv_halo2_kw_tc.cu), A/B switch for tools/ab_step.py.  Used from
# KWSTACK_MIN_CIN input channels on: measured at batch 256 (tools/probe_kwstack.py) 128 -> 64 at W = 32: 0.64 -> 0.44 ms
# (944 -> 1370 TFLOP/s); 64 -> 64 at W = 64: 1.20 -> 1.40 ms -- with one K block per tile the 12 instructions of a tile take
# 1150 cycles and the epilogue's 64 shuffles per warp and tile ~1400 (1.21 ms even without the cross-warp mailbox).
USE_KWSTACK = True
KWSTACK_MIN_CIN = 128


def conv3x3(x, wpack, cout, want_stats=False):
    """x (B,H,W,Cin) bf16, wpack [cout, 9*Cin] bf16 -> y (B,H,W,cout) bf16
    [, stats_partial (grid, 2, cout) fp32]."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    b, h, w, cin = x.shape
    assert wpack.shape == (cout, 9 * cin) and wpack.dtype == torch.bfloat16 and wpack.is_contiguous()
    y = torch.empty((b, h, w, cout), dtype=torch.bfloat16, device=x.device)
    stats = None
    pair = USE_2CTA
    with torch.cuda.device(x.device):
        if want_stats:
            grid = getattr(_lib.lib(), 'sed_conv3x3_tc2_grid' if pair else 'sed_conv3x3_tc_grid')(b, h, w, cin, cout)
            stats = torch.empty((grid, 2, cout), dtype=torch.float32, device=x.device)
        entry = 'sed_conv3x3_tc2_fwd' if pair else 'sed_conv3x3_tc_fwd'
        if pair and USE_KWSTACK and cin >= KWSTACK_MIN_CIN and _lib.lib().sed_conv3x3_tc2kw_supported(w, cin, cout):
            entry = 'sed_conv3x3_tc2kw_fwd'
        _lib.call(entry, x.data_ptr(), wpack.data_ptr(), y.data_ptr(),
                  _lib.ptr(stats), b, h, w, cin, cout, _lib.stream_of(x))
    return (y, stats) if want_stats else y


def conv3x3_dgrad_bnr(dy, wpack_dgrad, cin, y_below, st_below, pool):
    """Data gradient fused with the first BN-backward pass of the layer below.

    dy (B,H,W,Cout) bf16, wpack_dgrad [cin, 9*Cout] -> dx (B,H,W,cin) bf16 (= dA of the layer below) and
    gy_partial (grid, 2, cin) fp32: per-CTA sums of g and g*y over that layer's raw output ``y_below``
    (B, Hy, W*pool, cin), g = unpool(dx)/pool^2 * [bn(y) > 0]."""
    assert dy.is_cuda and dy.dtype == torch.bfloat16 and dy.is_contiguous()
    b, h, w, cout = dy.shape
    assert wpack_dgrad.shape == (cin, 9 * cout) and wpack_dgrad.is_contiguous()
    assert y_below.dtype == torch.bfloat16 and y_below.is_contiguous()
    yb, hy, wy, cy = y_below.shape
    assert yb == b and cy == cin and wy == w * pool and hy // pool == h, (y_below.shape, dy.shape, pool)
    dx = torch.empty((b, h, w, cin), dtype=torch.bfloat16, device=dy.device)
    with torch.cuda.device(dy.device):
        grid = _lib.lib().sed_conv3x3_tc_grid(b, h, w, cout, cin)
        partial = torch.empty((grid, 2, cin), dtype=torch.float32, device=dy.device)
        _lib.call('sed_conv3x3_tc_dgrad_bnr', dy.data_ptr(), wpack_dgrad.data_ptr(), dx.data_ptr(), b, h, w, cout,
                  cin, y_below.data_ptr(), hy, st_below.scale.data_ptr(), st_below.shift.data_ptr(), pool,
                  partial.data_ptr(), _lib.stream_of(dy))
    return dx, partial


def conv3x3_wgrad(dy, x, out=None, accumulate=False):
    """dy (B,H,W,Cout) bf16, x (B,H,W,Cin) bf16 -> dW fp32 (Cout, Cin, 3, 3)."""
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    assert dy.is_contiguous() and x.is_contiguous() and dy.shape[:3] == x.shape[:3]
    b, h, w, cin = x.shape
    cout = dy.shape[3]
    with torch.cuda.device(x.device):
        splits = _lib.lib().sed_conv3x3_tc_wgrad_splits(b, h, w, cin, cout)
        slabs = torch.empty((splits, 9, cout, cin), dtype=torch.float32, device=x.device)
        _lib.call('sed_conv3x3_tc_wgrad', dy.data_ptr(), x.data_ptr(), slabs.data_ptr(), b, h, w, cin,
                  cout, _lib.stream_of(x))
        if out is None:
            out = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=x.device)
            accumulate = False
        _lib.call('sed_conv_unpack_wgrad', slabs.data_ptr(), splits, 9 * cout * cin, cout, cin,
                  out.data_ptr(), 1 if accumulate else 0, _lib.stream_of(x))
    return out


def to_bf16(x):
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call('sed_f32_to_bf16', x.data_ptr(), y.data_ptr(), x.numel(), _lib.stream_of(x))
    return y
