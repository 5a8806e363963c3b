"""Multi-head self-attention layer of the CNN-Transformer variants
(/root/reference/pytorch/models.py:611-665 MultiHead, :587-608 ScaledDotProductAttention; call
sites :742, :836 -- always ``multihead(x, x, x)`` with ``mask=None``).

    qkv  = x @ [w_qs; w_ks; w_vs]^T + b      one tensor-core GEMM (3-way bf16 split, fp32-class accuracy)
    ctx  = softmax(q k^T / sqrt(d_k)) [dropout 0.1] v   per (batch, head), one fused tensor-core kernel (warp-level
                                              bf16x3 MMA tiles, csrc/attention_tc.cu), heads addressed in place (no
                                              permute/contiguous copies)
    out  = relu(dropout_0.2(ctx @ fc^T + b)) tensor-core GEMM + one elementwise kernel

No residual and no LayerNorm (both are dead code in the reference; ``layer_norm`` stays a
registered, unused parameter holder).  Features are time-major fp32 (B, T, 512) on both sides.
Dropout masks come from a Philox stream keyed by torch's CUDA generator (seed, offset) and are
recomputed in the backward; they are statistically, not bit-wise, equivalent to torch's.
"""
import torch

from . import conv as tcconv
from . import gemm
from ._lib import call, ptr, stream_of

F32 = torch.float32


class MhaCtx(object):
    __slots__ = ('x2d', 'qkv', 'probs', 'ctx2d', 'y', 'shape', 'p_att', 'p_out', 'seed', 'off_att', 'state_att')


class PhiloxIndirect(object):
    """Generator state through device memory, for CUDA-graph replay (trainer.FusedTrainer): while ``active`` is set the
    kernels read {seed, offset base} from the two int64 device words ``state`` at run time and the offsets handed out here
    are RELATIVE to the step's base (``blocks`` counts the Philox blocks the step has consumed so far); the trainer writes
    the generator's current (seed, offset) into ``state`` before every replay and advances the generator by ``blocks``."""
    __slots__ = ('state', 'blocks')

    def __init__(self, dev):
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)
        self.blocks = 0


INDIRECT = None        # a PhiloxIndirect while a step is captured / run through device-side generator state


def _generator(dev):
    return torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]


def _philox_state(dev, n):
    """(seed, offset in Philox blocks, device state pointer or 0) for n draws.  Direct mode: from torch's CUDA generator
    of ``dev``, which advances by n draws.  Indirect mode (INDIRECT set): offsets relative to the step's base."""
    blocks = (n + 3) // 4
    if INDIRECT is not None:
        off = INDIRECT.blocks
        INDIRECT.blocks += blocks
        return 0, off, INDIRECT.state.data_ptr()
    gen = _generator(dev)
    seed, off = gen.initial_seed(), gen.get_offset()
    gen.set_offset(off + 4 * blocks)
    return seed & 0xFFFFFFFFFFFFFFFF, off // 4, 0


def _check(mh):
    if mh.d_k != 64 or mh.d_v != 64:
        raise NotImplementedError('MultiHead: the B200 attention kernel implements d_k = d_v = 64 '
                                  '(the reference configuration, models.py:702-707)')


def multihead_forward(mh, feat, training, keep):
    """feat (B, T, d_model) fp32 -> (B, T, d_model) fp32, MhaCtx (None unless keep)."""
    _check(mh)
    b, t, c = feat.shape
    h, d = mh.n_head, mh.d_k
    dev = feat.device
    x2d = feat.contiguous().view(b * t, c)
    w_qkv = torch.cat([mh.w_qs.weight, mh.w_ks.weight, mh.w_vs.weight], dim=0)      # parameter plumbing
    b_qkv = torch.cat([mh.w_qs.bias, mh.w_ks.bias, mh.w_vs.bias], dim=0)
    qkv = gemm.linear_x3(x2d, w_qkv, b_qkv)                                          # (B*T, 3*h*d)
    ld = 3 * h * d
    p_att = float(mh.attention.dropout.p) if training else 0.0
    p_out = float(mh.dropout.p) if training else 0.0
    seed, off_att, st_att = _philox_state(dev, b * h * t * 128) if p_att > 0 else (0, 0, 0)   # element = score row * 128 + key
    ctx2d = torch.empty((b * t, h * d), dtype=F32, device=dev)
    probs = torch.empty((b, h, t, t), dtype=F32, device=dev) if keep else None
    temperature = float(mh.attention.temperature)
    with torch.cuda.device(dev):
        base = qkv.data_ptr()
        call('sed_attention_fwd', base, base + 4 * h * d, base + 8 * h * d, ld, ld, ld, b, t, h, d, temperature,
             p_att, seed, off_att, st_att, ctx2d.data_ptr(), ptr(probs), stream_of(feat))
        o = gemm.linear_x3(ctx2d, mh.fc.weight, mh.fc.bias)                           # (B*T, d_model)
        seed_o, off_o, st_o = _philox_state(dev, o.numel()) if p_out > 0 else (0, 0, 0)
        y = torch.empty_like(o)
        call('sed_dropout_relu_fwd', o.data_ptr(), o.numel(), p_out, seed_o, off_o, st_o, y.data_ptr(), stream_of(feat))
    mctx = None
    if keep:
        mctx = MhaCtx()
        mctx.x2d, mctx.qkv, mctx.probs, mctx.ctx2d, mctx.y = x2d, qkv, probs, ctx2d, y
        mctx.shape, mctx.p_att, mctx.p_out, mctx.seed, mctx.off_att = (b, t, c, h, d), p_att, p_out, seed, off_att
        mctx.state_att = st_att
    return y.view(b, t, mh.fc.weight.shape[0]), mctx


def multihead_backward(mh, mctx, dfeat, grad_of):
    """d out (B, T, d_model) -> d feat (B, T, d_model); parameter gradients through grad_of."""
    b, t, c, h, d = mctx.shape
    dev = dfeat.device
    dy = dfeat.contiguous().view(b * t, -1)
    ld = 3 * h * d
    with torch.cuda.device(dev):
        s = stream_of(dfeat)
        do = torch.empty_like(dy)
        call('sed_dropout_relu_bwd', dy.data_ptr(), mctx.y.data_ptr(), dy.numel(), mctx.p_out, do.data_ptr(), s)
        do16, ctx16 = tcconv.to_bf16(do), tcconv.to_bf16(mctx.ctx2d)
        g = grad_of(mh.fc.weight)
        if g is not None:
            gemm.gemm_tn(do16, ctx16, do.shape[1], h * d, g)
        g = grad_of(mh.fc.bias)
        if g is not None:
            gemm.colsum(do, g)
        dctx = gemm.gemm_nt(do16, gemm.transpose_bf16(mh.fc.weight))                  # (B*T, h*d)
        dqkv = torch.empty((b * t, ld), dtype=F32, device=dev)
        qb, gb = mctx.qkv.data_ptr(), dqkv.data_ptr()
        call('sed_attention_bwd', qb, qb + 4 * h * d, qb + 8 * h * d, ld, ld, ld, b, t, h, d,
             float(mh.attention.temperature), mctx.p_att, mctx.seed, mctx.off_att, mctx.state_att, dctx.data_ptr(),
             mctx.probs.data_ptr(), gb, gb + 4 * h * d, gb + 8 * h * d, s)
        dqkv16, x16 = tcconv.to_bf16(dqkv), tcconv.to_bf16(mctx.x2d)
        db = torch.empty(ld, dtype=F32, device=dev)
        gemm.colsum(dqkv, db)
        for i, lin in enumerate((mh.w_qs, mh.w_ks, mh.w_vs)):
            g = grad_of(lin.weight)
            if g is not None:
                gemm.gemm_tn(dqkv16, x16, h * d, c, g, a_col=i * h * d)
            g = grad_of(lin.bias)
            if g is not None:
                g.copy_(db[i * h * d:(i + 1) * h * d])
        w_qkv = torch.cat([mh.w_qs.weight, mh.w_ks.weight, mh.w_vs.weight], dim=0)
        dx = gemm.gemm_nt(dqkv16, gemm.transpose_bf16(w_qkv))                         # (B*T, d_model)
    return dx.view(b, t, c)


class _MultiHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mh, x, keep, *params):
        y, mctx = multihead_forward(mh, x.float(), mh.training, keep)
        ctx.mh, ctx.mctx, ctx.params = mh, mctx, params
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.mctx is None:
            raise RuntimeError('MultiHead backward without saved activations')
        grads = {}

        def grad_of(p):
            if p is None or not p.requires_grad:
                return None
            return grads.setdefault(p, torch.empty_like(p, dtype=F32))

        dx = multihead_backward(ctx.mh, ctx.mctx, dy.float(), grad_of)
        return (None, dx, None) + tuple(grads.get(p) for p in ctx.params)


def multihead_module_forward(mh, q, k, v, mask=None):
    """MultiHead.forward(q, k, v, mask=None) for the stand-alone module (seam B)."""
    if mask is not None:
        raise NotImplementedError('MultiHead: mask is always None in the reference (models.py:659) and is not implemented')
    if not q.is_cuda:
        raise RuntimeError('MultiHead: CUDA tensor required (no CPU path in this package)')
    same = (q.data_ptr() == k.data_ptr() == v.data_ptr()) and q.shape == k.shape == v.shape
    if not same:
        raise NotImplementedError('MultiHead: only self-attention multihead(x, x, x) -- the only form the '
                                  'reference models use (models.py:742, :836) -- is implemented')
    params = [p for p in (mh.w_qs.weight, mh.w_qs.bias, mh.w_ks.weight, mh.w_ks.bias, mh.w_vs.weight, mh.w_vs.bias,
                          mh.fc.weight, mh.fc.bias)]
    keep = torch.is_grad_enabled() and (q.requires_grad or any(p.requires_grad for p in params))
    return _MultiHeadFn.apply(mh, q, keep, *params)
