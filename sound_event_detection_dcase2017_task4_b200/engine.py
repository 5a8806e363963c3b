"""Composition of the libsedb200 kernels into the model forward / backward.

This is host plumbing only: it sequences C-ABI launches on torch's current stream and keeps
the tensors the backward needs.  It walks a model object that exposes the reference's
sub-module names (``bn0``, ``conv_block1..4`` with ``conv1/conv2/bn1/bn2``, ``fc`` /
``att_block`` / ``gru`` / ``multihead``, see models.py) and reads their parameters directly,
so the same code serves the drop-in ``nn.Module`` classes (through one autograd.Function) and
the benchmark's fused trainer (no autograd at all).

Data layout in HBM: log-mel fp32 (B2, T, 64); conv-1 input fp32
(B, T, 64); every other activation NHWC bf16 (B, H=frames, W=mel bins, C); raw conv outputs
bf16 with fp32 BatchNorm statistics taken from the fp32 accumulators; features after block 4
fp32 time-major (B, T/8, 512).
"""
import torch

from . import conv as tcconv
from . import frontend as fe
from . import ops
from . import specaug


class LayerCtx(object):
    __slots__ = ('conv', 'bn', 'x_in', 'y', 'st', 'ph', 'pw', 'is_c1', 'wd')


class TrunkCtx(object):
    __slots__ = ('logmel', 'st0', 't_stripes', 'f_stripes', 'lam', 'x0', 'layers', 'training', 'bn0')


def _bn_state(partial, count, bn, training):
    return ops.bn_finalize(partial, count, bn) if training else ops.bn_eval_affine(bn)


def _conv_layer(conv_mod, bn_mod, x_in, ph, pw, training, out_f32, keep, packs=None):
    """conv3x3 -> BN(train: batch stats) -> ReLU -> avgpool(ph, pw).  ``packs``: {conv module: (fwd, dgrad) bf16 weight
    packs} made for all layers at once (trunk_forward); without it the layer packs its own."""
    lc = LayerCtx()
    lc.conv, lc.bn, lc.ph, lc.pw = conv_mod, bn_mod, ph, pw
    w = conv_mod.weight
    cout, cin = w.shape[0], w.shape[1]
    if cin == 1:
        lc.is_c1 = True
        b, h, wd = x_in.shape
        y, partial = ops.conv_c1_fwd(x_in, w, want_stats=True)      # (this kernel always reduces its statistics)
        lc.wd = None
    else:
        lc.is_c1 = False
        b, h, wd, _ = x_in.shape
        if not tcconv.tc_supported(wd, cin, cout):
            raise NotImplementedError(
                'tensor-core conv path needs W | 128 (W >= 8), Cin in {64,128k}, Cout in {64,128,256,512}; '
                'got W=%d Cin=%d Cout=%d' % (wd, cin, cout))
        if packs is not None and conv_mod in packs:
            wf, lc.wd = packs[conv_mod]
        else:
            wf, lc.wd = tcconv.pack_weights(w, want_fwd=True, want_dgrad=keep)
        if training:
            y, partial = tcconv.conv3x3(x_in, wf, cout, want_stats=True)
        else:                                   # eval: running statistics, the epilogue skips the batch sums
            y, partial = tcconv.conv3x3(x_in, wf, cout, want_stats=False), None
    st = _bn_state(partial, b * h * wd, bn_mod, training)
    out = ops.bn_relu_pool_fwd(y, st, ph, pw, out_f32=out_f32)
    if keep:
        lc.x_in, lc.y, lc.st = x_in, y, st
    return out, lc


def trunk_forward(model, wave, lam, training, stripes=None, keep=None):
    """waveform (B2, L) fp32/int16 -> features (B, T/8, 512) fp32 time-major, TrunkCtx.

    ``lam``: fp32 (B2,) mixup coefficients or None (models.py:210-211).  ``stripes``: optional
    pre-drawn (t, f) int32 tables; by default they are drawn here from the torch CPU generator
    in the reference's order whenever ``training`` is set (models.py:206-207)."""
    ctx = TrunkCtx()
    keep = training if keep is None else keep       # save activations for trunk_backward
    ctx.training = training
    ctx.bn0 = model.bn0
    hop = model.spectrogram_extractor.stft.hop_length
    lmx = model.logmel_extractor
    bank = lmx.mel_bank()
    logmel = fe.logmel(wave, hop, bank, amin=lmx.amin, ref=lmx.ref)           # (B2, 1, T, M)
    b2, _, t, m = logmel.shape
    logmel = logmel.view(b2, t, m)
    if training:
        st0 = ops.bn_finalize(ops.colstats(logmel.view(b2 * t, m)), b2 * t, model.bn0)
        if stripes is None:
            aug = model.spec_augmenter
            ts, fs = specaug.draw_spec_augment(
                b2, t, m, aug.time_dropper.drop_width, aug.time_dropper.stripes_num,
                aug.freq_dropper.drop_width, aug.freq_dropper.stripes_num)
        else:
            ts, fs = stripes
        dev = logmel.device
        ctx.t_stripes = torch.as_tensor(ts, dtype=torch.int32).to(dev, non_blocking=True)
        ctx.f_stripes = torch.as_tensor(fs, dtype=torch.int32).to(dev, non_blocking=True)
        ctx.lam = None if lam is None else lam.to(dev, torch.float32).contiguous()
    else:
        st0 = ops.bn_eval_affine(model.bn0)
        ctx.t_stripes = ctx.f_stripes = ctx.lam = None
    x = ops.bn0_aug_mix_fwd(logmel, st0, ctx.t_stripes, ctx.f_stripes, ctx.lam)   # (B, T, M) fp32
    ctx.logmel, ctx.st0, ctx.x0 = (logmel, st0, x) if keep else (None, None, None)
    ctx.layers = []
    blocks = (model.conv_block1, model.conv_block2, model.conv_block3, model.conv_block4)
    # bf16 weight packs of all tensor-core layers in one launch (they were 7 launches of ~10 us)
    tc_convs = [c for blk in blocks for c in (blk.conv1, blk.conv2) if c.weight.shape[1] > 1]
    packs = dict(zip(tc_convs, tcconv.pack_weights_multi([c.weight for c in tc_convs], want_dgrad=keep)))
    for bi, blk in enumerate(blocks):
        last = bi == len(blocks) - 1
        x, lc = _conv_layer(blk.conv1, blk.bn1, x, 1, 1, training, False, keep, packs)
        ctx.layers.append(lc)
        wd = x.shape[2]
        # blocks 1-3: avg_pool 2x2; block 4: pool 1x1 then mean over the 8 remaining mel bins
        ph, pw = (1, wd) if last else (2, 2)
        x, lc = _conv_layer(blk.conv2, blk.bn2, x, ph, pw, training, last, keep, packs)
        ctx.layers.append(lc)
    b, tp, one, c = x.shape
    return x.view(b, tp, c), ctx


_SIDE_STREAMS = {}
# Side-stream schedule for the weight gradients (see trunk_backward).  Round 1 measured no gain (40.05 vs 39.8 ms/step)
# because the kernels could not share an SM: the BatchNorm-backward grids were sized as one full wave of 3 CTAs/SM x 80
# registers with up to 16 KB of shared memory, and the weight-gradient CTA took all 227 KB.  Round 2 rebuilt both sides
# for co-residency (csrc/bnpool.cu "backward fast paths", csrc/conv_wgrad_tc.cu kWgradSmemBudget).
OVERLAP_WGRAD = True
# Layers (index into ctx.layers, 0 = block1.conv1) whose weight gradient may go to the side stream.  The 64 -> 64 layer
# (index 1) was excluded while its weight-gradient kernel was bound by shared-memory operand reads (N = 64: beside it the
# BatchNorm reduce pass of layer 0 took 2.4 ms instead of 0.7, tools/timeline.py); with the kh taps stacked in N
# (csrc/conv_wgrad_tc.cu) it is included again: 28.51 -> 28.42, 28.69 -> 28.63 ms/step in a same-box A/B (tools/ab_step.py).
OVERLAP_LAYERS = frozenset((1, 2, 3, 4, 5, 6, 7))
# Fold the reduction pass of each BatchNorm backward into the epilogue of the data-gradient kernel that produces its dA
# (sed_conv3x3_tc_dgrad_bnr).  Correct (tests/test_gpu_conv.py) but measured SLOWER on B200 at batch 256
# (tools/ab_step.py: 42.1 vs 40.1 ms/step): the 8 epilogue warps cannot hide the HBM latency of the y reads that the
# dedicated 24-warp/SM reduction kernels stream at ~5 TB/s, and the epilogue is already the critical path of the
# K = 576 layers.  Off.
FUSE_BN_REDUCE = False
# Layer 1 (Cin = 1): BatchNorm-backward apply fused into the weight-gradient kernel (sed_bn_apply_conv_c1_wgrad).
FUSE_C1_APPLY = True


def _side_stream(device):
    """One auxiliary stream per (device, compute stream) for the weight-gradient kernels."""
    main = torch.cuda.current_stream(device)
    key = (device.index, main.cuda_stream)
    side = _SIDE_STREAMS.get(key)
    if side is None:
        side = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return main, side


def trunk_backward(ctx, dfeat, grad_of, overlap_wgrad=None):
    """dfeat (B, T/8, 512) fp32 -> parameter gradients written through ``grad_of(param)``
    (a callable returning the fp32 tensor to fill, or None to skip that parameter).

    Schedule: the critical chain BN-backward_l -> dgrad_l -> BN-backward_{l-1} ... stays on the current
    stream; the weight gradient of layer l only needs dY_l and the saved input, so it is enqueued on a side
    stream behind an event recorded right after BN-backward_l -- AFTER the data gradient has been enqueued, so
    the chain's tensor kernel is ahead of it in issue order.  The BN-backward kernels are HBM-bound, the
    weight-gradient kernels tensor-pipe-bound: co-resident on the same SMs they overlap instead of adding up.
    The current stream waits for the side stream before returning (the optimizer reads every gradient)."""
    if overlap_wgrad is None:
        overlap_wgrad = OVERLAP_WGRAD
    dA = dfeat.contiguous().view(dfeat.shape[0], dfeat.shape[1], 1, dfeat.shape[2])
    dev = dA.device
    main, side = _side_stream(dev) if overlap_wgrad else (None, None)
    used_side = False
    held = []
    side_slabs, main_slabs = [], []
    gy_partial = None
    for li in range(len(ctx.layers) - 1, -1, -1):
        lc = ctx.layers[li]
        gw = grad_of(lc.conv.weight)
        fuse_c1 = (FUSE_C1_APPLY and lc.is_c1 and gw is not None and dA.dtype == torch.bfloat16
                   and (lc.ph, lc.pw) == (1, 1) and lc.y.shape[3] == 64 and lc.y.shape[2] >= 16)
        if fuse_c1:
            # layer 1: the apply pass runs inside the weight-gradient kernel (one pass over y and dA instead of two)
            coef = ops.bn_bwd_coef(lc.y, dA, lc.st, lc.bn, lc.ph, lc.pw, grad_of(lc.bn.weight), grad_of(lc.bn.bias),
                                   frozen=not ctx.training)
            dy = ops.bn_apply_conv_c1_wgrad(lc.x_in, lc.y, dA, lc.st, coef, gw)
        else:
            dy = ops.bn_relu_pool_bwd(lc.y, dA, lc.st, lc.bn, lc.ph, lc.pw, grad_of(lc.bn.weight),
                                      grad_of(lc.bn.bias), gy_partial=gy_partial, frozen=not ctx.training)
        gy_partial = None
        side_wgrad = overlap_wgrad and gw is not None and not lc.is_c1 and li in OVERLAP_LAYERS
        if side_wgrad:
            ready = torch.cuda.Event()
            ready.record(main)
        x_in = lc.x_in
        if lc.is_c1:
            if gw is not None and not fuse_c1:
                ops.conv_c1_wgrad(x_in, dy, gw)
            dA = ops.conv_c1_dgrad(dy, lc.conv.weight)                 # (B, T, M) fp32
        else:
            below = ctx.layers[li - 1]
            cin = lc.conv.weight.shape[1]
            if FUSE_BN_REDUCE and (below.ph, below.pw) in ((1, 1), (2, 2)) and below.y.shape[2] % below.pw == 0:
                # the data gradient also produces the first pass of the BatchNorm backward of the layer below
                dA, gy_partial = tcconv.conv3x3_dgrad_bnr(dy, lc.wd, cin, below.y, below.st, below.ph)
            else:
                dA = tcconv.conv3x3(dy, lc.wd, cin)                    # bf16 NHWC
            if side_wgrad:
                side.wait_event(ready)
                # dy / x_in belong to the main stream's allocator pool.  They are kept alive in `held` until the main
                # stream has been ordered behind the side stream (end of this function), so whatever reuses their
                # memory later runs after the weight gradient -- no Tensor.record_stream: its deferred frees made the
                # caching allocator grow by ~35 GB and cudaMalloc inside timed steps (tools/hiccup.py: one 130 ms step).
                held.append((dy, x_in))
                used_side = True
                with torch.cuda.stream(side):
                    side_slabs.append((tcconv.conv3x3_wgrad_slabs(dy, x_in), gw))
            elif gw is not None:
                main_slabs.append((tcconv.conv3x3_wgrad_slabs(dy, x_in), gw))
        lc.y = lc.x_in = None
    g0w, g0b = grad_of(ctx.bn0.weight), grad_of(ctx.bn0.bias)
    if g0w is not None or g0b is not None:
        ops.bn0_bwd(dA, ctx.logmel, ctx.st0, ctx.bn0, ctx.t_stripes, ctx.f_stripes, ctx.lam, g0w, g0b)
    # the split-K slabs of all weight gradients are folded into the OIHW gradients by one launch per stream
    tcconv.unpack_wgrad_multi(main_slabs)
    if used_side:
        with torch.cuda.stream(side):
            tcconv.unpack_wgrad_multi(side_slabs)
        done = torch.cuda.Event()
        done.record(side)
        main.wait_event(done)
    del held


# ------------------------------------------------------------------ heads
class HeadCtx(object):
    __slots__ = ('kind', 'feat2d', 'shape', 'prob', 'argmax', 'att_logit', 'norm_att', 'cla', 'clip', 'temporal')


def head_forward(model, feat, ratio, want_frame=True, keep=True):
    """features (B, T', C) -> dict(clipwise_output, framewise_output, embedding), HeadCtx."""
    b, tp, c = feat.shape
    hc = HeadCtx()
    hc.shape = (b, tp, c)
    feat2d = feat.contiguous().view(b * tp, c)
    hc.feat2d = feat2d if keep else None
    pooling = model.pooling
    if pooling in ('avg', 'max'):
        hc.kind = pooling
        fc = model.fc
        logit = ops.linear_small_fwd(feat2d, fc.weight, fc.bias).view(b, tp, -1)
        prob, clip, argmax, frame = ops.head_pool_fwd(logit, ratio, 0 if pooling == 'avg' else 1, want_frame)
        hc.prob, hc.argmax = (prob, argmax) if keep else (None, None)
        emb = feat.transpose(1, 2)
    else:
        hc.kind = 'att'
        ab = model.att_block
        k = ab.att.weight.shape[0]
        att_logit, cla_logit = ops.linear_pair_fwd(feat2d, ab.att.weight.view(k, c), ab.att.bias,
                                                   ab.cla.weight.view(k, c), ab.cla.bias)
        att_logit, cla_logit = att_logit.view(b, tp, k), cla_logit.view(b, tp, k)
        clip, norm_att, cla, frame = ops.head_att_fwd(att_logit, cla_logit, ratio, ab.activation == 'sigmoid',
                                                      ab.temperature, want_frame)
        if keep:
            hc.att_logit, hc.norm_att, hc.cla, hc.clip = att_logit, norm_att, cla, clip
        emb = cla
    return {'framewise_output': frame, 'clipwise_output': clip, 'embedding': emb}, hc


def head_backward(model, hc, dclip, grad_of):
    """dclip (B, K) fp32 -> d features (B, T', C); head parameter grads through grad_of."""
    b, tp, c = hc.shape
    dclip = dclip.contiguous()
    if hc.kind in ('avg', 'max'):
        fc = model.fc
        dlogit = ops.head_pool_bwd(hc.prob, dclip, hc.argmax, 0 if hc.kind == 'avg' else 1)
        k = dlogit.shape[2]
        dfeat = ops.linear_small_bwd(dlogit.view(b * tp, k), hc.feat2d, fc.weight, grad_of(fc.weight),
                                     grad_of(fc.bias))
    else:
        ab = model.att_block
        k = ab.att.weight.shape[0]
        d_att, d_cla = ops.head_att_bwd(hc.att_logit, hc.norm_att, hc.cla, hc.clip, dclip,
                                        ab.activation == 'sigmoid', ab.temperature)
        gw = grad_of(ab.att.weight)
        dfa = ops.linear_small_bwd(d_att.view(b * tp, k), hc.feat2d, ab.att.weight.view(k, c),
                                   None if gw is None else gw.view(k, c), grad_of(ab.att.bias))
        gw = grad_of(ab.cla.weight)
        dfeat = ops.linear_small_bwd(d_cla.view(b * tp, k), hc.feat2d, ab.cla.weight.view(k, c),
                                     None if gw is None else gw.view(k, c), grad_of(ab.cla.bias), dx=dfa)
    return dfeat.view(b, tp, c)
