"""Temporal module between the CNN trunk and the classifier: none, bidirectional GRU
(/root/reference/pytorch/models.py:475, :566) or one multi-head self-attention layer
(models.py:742, :836).  Features are time-major fp32 (B, T', 512) on both sides."""
import torch

from . import _lib
from . import gemm
from . import ops
from ._lib import call, stream_of

F32 = torch.float32


class GruCtx(object):
    __slots__ = ('x2d', 'x3', 'out', 'gates', 'shape')


def _gru_cat(gru):
    """(W_ih (2*3H, I), b_ih (2*3H), W_hh (2, 3H, H), b_hh (2, 3H)) with both directions stacked.
    torch.cat here is parameter plumbing (1.2 M floats), not activation math."""
    w_ih = torch.cat([gru.weight_ih_l0, gru.weight_ih_l0_reverse], dim=0)
    b_ih = torch.cat([gru.bias_ih_l0, gru.bias_ih_l0_reverse], dim=0)
    w_hh = torch.stack([gru.weight_hh_l0, gru.weight_hh_l0_reverse], dim=0).contiguous()
    b_hh = torch.stack([gru.bias_hh_l0, gru.bias_hh_l0_reverse], dim=0).contiguous()
    return w_ih, b_ih, w_hh, b_hh


def gru_forward(gru, feat, keep):
    b, t, c = feat.shape
    h = gru.hidden_size
    if not (gru.num_layers == 1 and gru.bidirectional and gru.batch_first and gru.bias):
        raise NotImplementedError('only the reference GRU configuration is implemented')
    w_ih, b_ih, w_hh, b_hh = _gru_cat(gru)
    x2d = feat.contiguous().view(b * t, c)
    x3 = gemm.split3(x2d, 0)                                     # bf16 (B*T, 3C) = [hi | lo | hi]
    gx = gemm.gemm_nt(x3, gemm.split3(w_ih, 1), b_ih)            # (B*T, 2*3H) fp32, fp32-class accuracy
    out = torch.empty((b, t, 2 * h), dtype=F32, device=feat.device)
    gates = torch.empty((b, t, 2, 4, h), dtype=F32, device=feat.device)
    sync_ws = torch.empty(_lib.lib().sed_gru_workspace_bytes(b, h, 0), dtype=torch.uint8, device=feat.device)
    with torch.cuda.device(feat.device):
        call('sed_gru_fwd', gx.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(), out.data_ptr(), gates.data_ptr(),
             sync_ws.data_ptr(), b, t, h, stream_of(feat))
    ctx = None
    if keep:
        ctx = GruCtx()
        ctx.x2d, ctx.x3, ctx.out, ctx.gates, ctx.shape = x2d, x3, out, gates, (b, t, c, h)
    return out, ctx


def gru_backward(gru, ctx, dout, grad_of):
    b, t, c, h = ctx.shape
    dev = dout.device
    w_ih, _, w_hh, _ = _gru_cat(gru)
    dout = dout.contiguous()
    carry = torch.empty(_lib.lib().sed_gru_workspace_bytes(b, h, 1), dtype=torch.uint8, device=dev)
    dgx = torch.empty((b * t, 2 * 3 * h), dtype=F32, device=dev)
    dgh = torch.empty((b * t, 2 * 3 * h), dtype=F32, device=dev)
    hprev = torch.empty((b * t, 2 * h), dtype=F32, device=dev)
    from . import conv as tcconv
    fused16 = h == 256                      # the persistent kernel also emits the bf16 operands of the GEMMs below
    if fused16:
        dgx16 = torch.empty(dgx.shape, dtype=torch.bfloat16, device=dev)
        dgh16 = torch.empty(dgh.shape, dtype=torch.bfloat16, device=dev)
        hp16 = torch.empty(hprev.shape, dtype=torch.bfloat16, device=dev)
        # ... and per-(batch tile, half) partial sums of dGx / dGh over (batch, time): the bias gradients
        bias_partial = torch.empty((_lib.lib().sed_gru_bwd_bias_rows(b), 2, 2 * 3 * h), dtype=F32, device=dev)
    with torch.cuda.device(dev):
        call('sed_gru_bwd', dout.data_ptr(), ctx.out.data_ptr(), ctx.gates.data_ptr(), w_hh.data_ptr(),
             carry.data_ptr(), dgx.data_ptr(), dgh.data_ptr(), hprev.data_ptr(),
             dgx16.data_ptr() if fused16 else 0, dgh16.data_ptr() if fused16 else 0, hp16.data_ptr() if fused16 else 0,
             bias_partial.data_ptr() if fused16 else 0, b, t, h, stream_of(dout))
    if not fused16:
        dgx16, dgh16, hp16 = tcconv.to_bf16(dgx), tcconv.to_bf16(dgh), tcconv.to_bf16(hprev)
    # X as bf16: the hi third of the [hi | lo | hi] split the forward projection already made (row stride 3 * C)
    x16 = ctx.x3 if ctx.x3 is not None else tcconv.to_bf16(ctx.x2d)
    # weight gradients: dW_ih[d] = dGx_d^T X ; dW_hh[d] = dGh_d^T Hprev_d ; biases = column sums
    if fused16:
        db = ops.reduce_partials(bias_partial, torch.empty((2, 2 * 3 * h), dtype=F32, device=dev))
        db_ih, db_hh = db[0], db[1]
    else:
        db_ih = torch.empty(2 * 3 * h, dtype=F32, device=dev)
        db_hh = torch.empty(2 * 3 * h, dtype=F32, device=dev)
        gemm.colsum(dgx, db_ih)
        gemm.colsum(dgh, db_hh)
    for d, sfx in enumerate(('', '_reverse')):
        g = grad_of(getattr(gru, 'weight_ih_l0' + sfx))
        if g is not None:
            gemm.gemm_tn(dgx16, x16, 3 * h, c, g, a_col=d * 3 * h, ldb=x16.shape[1])
        g = grad_of(getattr(gru, 'weight_hh_l0' + sfx))
        if g is not None:
            gemm.gemm_tn(dgh16, hp16, 3 * h, h, g, a_col=d * 3 * h, b_col=d * h)
        g = grad_of(getattr(gru, 'bias_ih_l0' + sfx))
        if g is not None:
            g.copy_(db_ih[d * 3 * h:(d + 1) * 3 * h])
        g = grad_of(getattr(gru, 'bias_hh_l0' + sfx))
        if g is not None:
            g.copy_(db_hh[d * 3 * h:(d + 1) * 3 * h])
    # input gradient: dX = dGx (B*T, 2*3H) @ W_ih_cat (2*3H, I)
    dx = gemm.gemm_nt(dgx16, gemm.transpose_bf16(w_ih))
    return dx.view(b, t, c)


def forward(model, feat, training, keep):
    kind = model.temporal_kind
    if kind is None:
        return feat, None
    if kind == 'gru':
        return gru_forward(model.gru, feat, keep)
    if kind == 'mha':
        from . import attention
        return attention.multihead_forward(model.multihead, feat, training, keep)
    raise ValueError(kind)


def backward(model, ctx, dfeat, grad_of):
    kind = model.temporal_kind
    if kind is None:
        return dfeat
    if kind == 'gru':
        return gru_backward(model.gru, ctx, dfeat, grad_of)
    if kind == 'mha':
        from . import attention
        return attention.multihead_backward(model.multihead, ctx, dfeat, grad_of)
    raise ValueError(kind)
