"""Thin launch wrappers (one Python function per C-ABI entry point) for the BN / pooling /
prep / first-conv / head / loss / optimizer kernels.  Allocation is PyTorch's caching allocator;
the stream is torch's current stream; nothing here computes on the host."""
import torch

from . import _lib
from ._lib import call, ptr, stream_of

F32 = torch.float32
BF16 = torch.bfloat16


def _dev(t):
    return torch.cuda.device(t.device)


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


# ------------------------------------------------------------------ BatchNorm statistics
class BNState(object):
    """Per-layer, per-step BatchNorm quantities (all fp32 (C,) device tensors)."""
    __slots__ = ('scale', 'shift', 'mean', 'invstd', 'count')


def bn_finalize(partial, count, bn, training_update=True):
    """partial (P, 2, C) -> BNState; updates bn.running_* / num_batches_tracked in place like
    nn.BatchNorm2d in train mode (momentum 0.1, unbiased running var)."""
    P, _, C = partial.shape
    st = BNState()
    buf = _empty((4, C), F32, partial)
    st.scale, st.shift, st.mean, st.invstd = buf[0], buf[1], buf[2], buf[3]
    st.count = float(count)
    upd = training_update and bn.track_running_stats and bn.running_mean is not None
    momentum = 0.1 if bn.momentum is None else bn.momentum
    with _dev(partial):
        call('sed_bn_finalize', partial.data_ptr(), P, C, float(count), ptr(bn.weight), ptr(bn.bias),
             bn.eps, momentum, ptr(bn.running_mean) if upd else 0, ptr(bn.running_var) if upd else 0,
             ptr(bn.num_batches_tracked) if upd else 0, st.scale.data_ptr(), st.shift.data_ptr(),
             st.mean.data_ptr(), st.invstd.data_ptr(), stream_of(partial))
    return st


def bn_eval_affine(bn):
    """Eval mode: scale / shift from the running statistics; mean / invstd are the frozen statistics (the backward
    of an eval-mode forward under autograd treats them as constants)."""
    C = bn.num_features
    st = BNState()
    buf = _empty((4, C), F32, bn.running_mean)
    st.scale, st.shift, st.mean, st.invstd = buf[0], buf[1], buf[2], buf[3]
    st.count = 0.0
    with _dev(buf):
        call('sed_bn_eval_affine', bn.running_mean.data_ptr(), bn.running_var.data_ptr(), ptr(bn.weight),
             ptr(bn.bias), bn.eps, C, st.scale.data_ptr(), st.shift.data_ptr(), st.mean.data_ptr(),
             st.invstd.data_ptr(), stream_of(buf))
    return st


def bn_relu_pool_fwd(y, st, ph, pw, out_f32=False):
    b, h, w, c = y.shape
    out = _empty((b, h // ph, w // pw, c), F32 if out_f32 else BF16, y)
    with _dev(y):
        call('sed_bn_relu_pool_fwd', y.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(), b, h, w, c,
             ph, pw, out.data_ptr(), 1 if out_f32 else 0, stream_of(y))
    return out


_SCHED = {}


def sched_words(t):
    """Zeroed int32 scheduling words per (device, stream) for the dynamically scheduled BatchNorm-backward kernels
    (work tickets + worker-group arrival counters; the kernels leave them zero, see include/sed_b200.h)."""
    key = (t.device.index, stream_of(t))
    w = _SCHED.get(key)
    if w is None:
        w = _SCHED[key] = torch.zeros(_lib.lib().sed_bn_bwd_sched_words(), dtype=torch.int32, device=t.device)
    return w


def bn_bwd_coef(y, dA, st, bn, ph, pw, dgamma, dbeta, gy_partial=None, frozen=False):
    """First half of the BN+ReLU+pool backward: the reduction pass over (y, dA) and the finalize kernel.  Writes the bn
    parameter grads into dgamma / dbeta (may be None) and returns the (3, C) coefficient rows of the apply pass."""
    b, h, w, c = y.shape
    f32 = 1 if dA.dtype == F32 else 0
    with _dev(y):
        coef = _empty((3, c), F32, y)
        s = stream_of(y)
        sched = sched_words(y)
        if gy_partial is None:
            P = _lib.lib().sed_bn_bwd_partials(b, h, w, c, ph, pw)
            partial = _empty((_lib.lib().sed_bn_bwd_workspace_rows(b, h, w, c, ph, pw), 2, c), F32, y)
            call('sed_bn_relu_pool_bwd_reduce', y.data_ptr(), dA.data_ptr(), f32, st.scale.data_ptr(),
                 st.shift.data_ptr(), b, h, w, c, ph, pw, partial.data_ptr(), sched.data_ptr(), s)
        else:
            partial, P = gy_partial, gy_partial.shape[0]
        call('sed_bn_bwd_finalize', partial.data_ptr(), P, c, float(b * h * w), ptr(bn.weight),
             st.invstd.data_ptr(), st.mean.data_ptr(), ptr(dgamma), ptr(dbeta), 0, 1 if frozen else 0,
             coef.data_ptr(), s)
    return coef


def bn_bwd_apply(y, dA, st, coef, ph, pw):
    """Second half: dY (bf16, same shape as y) from y, dA and the coefficient rows."""
    b, h, w, c = y.shape
    f32 = 1 if dA.dtype == F32 else 0
    with _dev(y):
        dy = _empty(y.shape, BF16, y)
        call('sed_bn_relu_pool_bwd_apply', y.data_ptr(), dA.data_ptr(), f32, st.scale.data_ptr(),
             st.shift.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(), coef.data_ptr(), b, h, w, c,
             ph, pw, dy.data_ptr(), sched_words(y).data_ptr(), stream_of(y))
    return dy


def bn_relu_pool_bwd(y, dA, st, bn, ph, pw, dgamma, dbeta, gy_partial=None, frozen=False):
    """Two-pass BN+ReLU+pool backward.  Writes bn parameter grads into dgamma/dbeta (fp32 (C,)
    tensors, may be None) and returns dY (bf16, same shape as y).

    ``gy_partial``: (P, 2, C) partial sums of g and g*y already produced by the data-gradient kernel that wrote
    ``dA`` (conv.conv3x3_dgrad_bnr); the reduction pass over y and dA is then skipped.
    ``frozen``: ``st`` holds running statistics (eval-mode forward): they are constants, dY = gamma * invstd * g."""
    coef = bn_bwd_coef(y, dA, st, bn, ph, pw, dgamma, dbeta, gy_partial=gy_partial, frozen=frozen)
    return bn_bwd_apply(y, dA, st, coef, ph, pw)


def bn_apply_conv_c1_wgrad(x0, y, dA, st, coef, out):
    """The apply pass of block1.bn1's backward fused into the Cin = 1 weight-gradient kernel: returns dY (bf16, for the
    data gradient) and writes dW into ``out`` (Cout, 1, 3, 3).  y, dA: (B, H, W, 64) bf16; x0: (B, H, W) fp32."""
    b, h, wd = x0.shape
    cout = y.shape[3]
    assert dA.dtype == BF16 and y.dtype == BF16 and dA.shape == y.shape
    with _dev(x0):
        dy = _empty(y.shape, BF16, y)
        partial = _empty((_lib.lib().sed_conv_c1_grid(), cout * 9), F32, x0)
        call('sed_bn_apply_conv_c1_wgrad', x0.data_ptr(), y.data_ptr(), dA.data_ptr(), st.scale.data_ptr(),
             st.shift.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(), coef.data_ptr(), dy.data_ptr(),
             partial.data_ptr(), b, h, wd, cout, stream_of(x0))
    reduce_partials(partial, out)
    return dy


# ------------------------------------------------------------------ bn0 + SpecAug + mixup
def colstats(x2d):
    rows, c = x2d.shape
    with _dev(x2d):
        P = _lib.lib().sed_stat_partials()
        partial = _empty((P, 2, c), F32, x2d)
        call('sed_colstats_f32', x2d.data_ptr(), rows, c, partial.data_ptr(), stream_of(x2d))
    return partial


def bn0_aug_mix_fwd(logmel, st, t_stripes, f_stripes, lam):
    """logmel (B2, T, M) fp32 -> (Bout, T, M) fp32; stripes int32 (B2, n, 2) or None; lam fp32 (B2,) or None."""
    b2, t, m = logmel.shape
    bout = b2 // 2 if lam is not None else b2
    out = _empty((bout, t, m), F32, logmel)
    nt = 0 if t_stripes is None else t_stripes.shape[1]
    nf = 0 if f_stripes is None else f_stripes.shape[1]
    with _dev(logmel):
        call('sed_bn0_aug_mix_fwd', logmel.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(),
             ptr(t_stripes), nt, ptr(f_stripes), nf, ptr(lam), b2, t, m, out.data_ptr(), stream_of(logmel))
    return out


def bn0_bwd(dout, logmel, st, bn, t_stripes, f_stripes, lam, dgamma, dbeta):
    b2, t, m = logmel.shape
    nt = 0 if t_stripes is None else t_stripes.shape[1]
    nf = 0 if f_stripes is None else f_stripes.shape[1]
    with _dev(logmel):
        P = _lib.lib().sed_stat_partials()
        partial = _empty((P, 2, m), F32, logmel)
        s = stream_of(logmel)
        call('sed_bn0_bwd_reduce', dout.data_ptr(), logmel.data_ptr(), st.mean.data_ptr(),
             st.invstd.data_ptr(), ptr(t_stripes), nt, ptr(f_stripes), nf, ptr(lam), b2, t, m,
             partial.data_ptr(), s)
        # bn0: d(gamma) = sum d*xhat, d(beta) = sum d  (no dX needed: the waveform takes no gradient)
        call('sed_bn_bwd_finalize', partial.data_ptr(), P, m, float(b2 * t), 0, st.invstd.data_ptr(), 0,
             ptr(dgamma), ptr(dbeta), 0, 0, 0, s)


def spec_augment_(x, t_stripes, f_stripes):
    b, c, t, f = x.shape
    assert x.is_contiguous() and x.dtype == F32
    nt = 0 if t_stripes is None else t_stripes.shape[1]
    nf = 0 if f_stripes is None else f_stripes.shape[1]
    with _dev(x):
        call('sed_spec_augment_f32', x.data_ptr(), b, c, t, f, ptr(t_stripes), nt, ptr(f_stripes), nf,
             stream_of(x))
    return x


def mix_pairs(x, lam):
    """do_mixup (pytorch_utils.py:80-93) of a (B2, n) fp32 matrix: the targets at main.py:246."""
    x, lam = x.contiguous().float(), lam.contiguous().float()
    b2 = x.shape[0]
    n = x.numel() // max(b2, 1)
    out = _empty((b2 // 2,) + tuple(x.shape[1:]), F32, x)
    with _dev(x):
        call('sed_mix_pairs_f32', x.data_ptr(), lam.data_ptr(), b2, n, out.data_ptr(), stream_of(x))
    return out


def reduce_partials(partial, out, scale=1.0, accumulate=False):
    P = partial.shape[0]
    n = partial.numel() // P
    assert out.numel() == n
    with _dev(partial):
        call('sed_reduce_partials', partial.data_ptr(), P, n, out.data_ptr(), 1 if accumulate else 0,
             float(scale), stream_of(partial))
    return out


# ------------------------------------------------------------------ first convolution (Cin = 1)
def conv_c1_fwd(x0, w, want_stats=True):
    """x0 (B,H,W) fp32, w (Cout,1,3,3) fp32 -> y (B,H,W,Cout) bf16 [, partial]."""
    b, h, wd = x0.shape
    cout = w.shape[0]
    y = _empty((b, h, wd, cout), BF16, x0)
    partial = None
    with _dev(x0):
        if want_stats:
            partial = _empty((_lib.lib().sed_conv_c1_grid(), 2, cout), F32, x0)
        call('sed_conv_c1_fwd', x0.data_ptr(), w.data_ptr(), y.data_ptr(), ptr(partial), b, h, wd, cout,
             stream_of(x0))
    return (y, partial) if want_stats else y


def conv_c1_wgrad(x0, dy, out):
    b, h, wd = x0.shape
    cout = dy.shape[3]
    with _dev(x0):
        partial = _empty((_lib.lib().sed_conv_c1_grid(), cout * 9), F32, x0)
        call('sed_conv_c1_wgrad', x0.data_ptr(), dy.data_ptr(), partial.data_ptr(), b, h, wd, cout,
             stream_of(x0))
    return reduce_partials(partial, out)


def conv_c1_dgrad(dy, w):
    b, h, wd, cout = dy.shape
    dx = _empty((b, h, wd), F32, dy)
    with _dev(dy):
        call('sed_conv_c1_dgrad', dy.data_ptr(), w.data_ptr(), dx.data_ptr(), b, h, wd, cout, stream_of(dy))
    return dx


# ------------------------------------------------------------------ heads + loss
def linear_small_fwd(x2d, W, bias):
    r, c = x2d.shape
    k = W.shape[0]
    out = _empty((r, k), F32, x2d)
    with _dev(x2d):
        call('sed_linear_small_fwd', x2d.data_ptr(), W.data_ptr(), ptr(bias), r, c, k, out.data_ptr(),
             stream_of(x2d))
    return out


def linear_pair_fwd(x2d, W1, b1, W2, b2):
    """Two small linear maps of the same input in one pass over x2d: (x W1^T + b1, x W2^T + b2)."""
    r, c = x2d.shape
    k1, k2 = W1.shape[0], W2.shape[0]
    if c > 512 or k1 + k2 > 40:
        return linear_small_fwd(x2d, W1, b1), linear_small_fwd(x2d, W2, b2)
    out1 = _empty((r, k1), F32, x2d)
    out2 = _empty((r, k2), F32, x2d)
    with _dev(x2d):
        call('sed_linear_pair_fwd', x2d.data_ptr(), W1.data_ptr(), ptr(b1), k1, W2.data_ptr(), ptr(b2), k2, r, c,
             out1.data_ptr(), out2.data_ptr(), stream_of(x2d))
    return out1, out2


def linear_small_bwd(dout, x2d, W, dW, dbias, want_dx=True, dx=None):
    """dx given => accumulate into it (dx += dout @ W)."""
    r, c = x2d.shape
    k = W.shape[0]
    acc = 1 if dx is not None else 0
    if dx is None:
        dx = _empty((r, c), F32, x2d) if want_dx else None
    with _dev(x2d):
        P = _lib.lib().sed_linear_partials()
        pw = _empty((P, k, c), F32, x2d) if dW is not None else None
        pb = _empty((P, k), F32, x2d) if dbias is not None else None
        call('sed_linear_small_bwd', dout.data_ptr(), x2d.data_ptr(), W.data_ptr(), r, c, k, ptr(dx), acc,
             ptr(pw), ptr(pb), stream_of(x2d))
    if dW is not None:
        reduce_partials(pw, dW)
    if dbias is not None:
        reduce_partials(pb, dbias)
    return dx


def head_pool_fwd(logit, ratio, mode, want_frame=True):
    b, t, k = logit.shape
    prob = _empty((b, t, k), F32, logit)
    clip = _empty((b, k), F32, logit)
    argmax = _empty((b, k), torch.int32, logit) if mode == 1 else None
    frame = _empty((b, t * ratio, k), F32, logit) if want_frame else None
    with _dev(logit):
        call('sed_head_pool_fwd', logit.data_ptr(), b, t, k, ratio, mode, prob.data_ptr(), clip.data_ptr(),
             ptr(argmax), ptr(frame), stream_of(logit))
    return prob, clip, argmax, frame


def head_pool_bwd(prob, dclip, argmax, mode):
    b, t, k = prob.shape
    dlogit = _empty((b, t, k), F32, prob)
    with _dev(prob):
        call('sed_head_pool_bwd', prob.data_ptr(), dclip.data_ptr(), ptr(argmax), b, t, k, mode,
             dlogit.data_ptr(), stream_of(prob))
    return dlogit


def head_att_fwd(att_logit, cla_logit, ratio, sigmoid_act=True, temperature=1.0, want_frame=True):
    b, t, k = att_logit.shape
    norm_att = _empty((b, k, t), F32, att_logit)
    cla = _empty((b, k, t), F32, att_logit)
    clip = _empty((b, k), F32, att_logit)
    frame = _empty((b, t * ratio, k), F32, att_logit) if want_frame else None
    with _dev(att_logit):
        call('sed_head_att_fwd', att_logit.data_ptr(), cla_logit.data_ptr(), b, t, k, ratio,
             1 if sigmoid_act else 0, float(temperature), norm_att.data_ptr(), cla.data_ptr(), clip.data_ptr(),
             ptr(frame), stream_of(att_logit))
    return clip, norm_att, cla, frame


def head_att_bwd(att_logit, norm_att, cla, clip, dclip, sigmoid_act=True, temperature=1.0):
    b, t, k = att_logit.shape
    d_att = _empty((b, t, k), F32, att_logit)
    d_cla = _empty((b, t, k), F32, att_logit)
    with _dev(att_logit):
        call('sed_head_att_bwd', att_logit.data_ptr(), norm_att.data_ptr(), cla.data_ptr(), clip.data_ptr(),
             dclip.data_ptr(), b, t, k, 1 if sigmoid_act else 0, float(temperature), d_att.data_ptr(),
             d_cla.data_ptr(), stream_of(att_logit))
    return d_att, d_cla


def bce(prob, target, want_grad=True, grad_scale=1.0):
    """mean binary cross entropy (torch semantics) + its gradient wrt prob."""
    assert prob.shape == target.shape and prob.dtype == F32 and target.dtype == F32
    prob, target = prob.contiguous(), target.contiguous()
    loss = _empty((), F32, prob)
    dprob = _empty(prob.shape, F32, prob) if want_grad else None
    with _dev(prob):
        call('sed_bce_fwd_bwd', prob.data_ptr(), target.data_ptr(), prob.numel(), float(grad_scale),
             loss.data_ptr(), ptr(dprob), stream_of(prob))
    return loss, dprob


def adam_amsgrad_(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0,
                  bias_corr=None):
    """``bias_corr``: optional device tensor of two floats (1 - beta1^t, sqrt(1 - beta2^t)) read by the kernel instead
    of the values derived from ``step`` (CUDA-graph replays)."""
    n = param.numel()
    with _dev(param):
        call('sed_adam_amsgrad', param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
             max_exp_avg_sq.data_ptr(), n, lr, beta1, beta2, eps, step, float(grad_scale), ptr(bias_corr),
             stream_of(param))
