"""Stand-alone forwards of the reference's building-block modules (seam B: ``ConvBlock``,
``AttBlock``, ``MultiHead`` used outside the seven model classes).  They take and return the
reference's tensor layouts (NCHW fp32 / (B, C, T) fp32) and run the same kernels as the fused
model path; gradients flow through small autograd.Functions."""
import torch

from . import conv as tcconv
from . import engine
from . import ops


class _ConvBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, block, x, ph, pw, keep, w1, w2, g1, b1, g2, b2):
        training = block.training
        b, cin, h, w = x.shape
        if cin == 1:
            x_in = x.reshape(b, h, w).float().contiguous()
        else:
            x_in = tcconv.to_bf16(x.permute(0, 2, 3, 1).float().contiguous())
        a1, l1 = engine._conv_layer(block.conv1, block.bn1, x_in, 1, 1, training, False, keep)
        a2, l2 = engine._conv_layer(block.conv2, block.bn2, a1, ph, pw, training, True, keep)
        ctx.layers = (l1, l2) if keep else None
        ctx.cin = cin
        ctx.frozen = not training
        return a2.permute(0, 3, 1, 2)                       # NCHW view of the fp32 NHWC result

    @staticmethod
    def backward(ctx, dout):
        if ctx.layers is None:
            raise RuntimeError('ConvBlock backward without saved activations')
        l1, l2 = ctx.layers
        grads = {}

        def grad_of(p):
            if p is None or not p.requires_grad:
                return None
            return grads.setdefault(p, torch.empty_like(p, dtype=torch.float32))

        dA = dout.permute(0, 2, 3, 1).float().contiguous()
        dx = None
        for lc in (l2, l1):
            dy = ops.bn_relu_pool_bwd(lc.y, dA, lc.st, lc.bn, lc.ph, lc.pw, grad_of(lc.bn.weight),
                                      grad_of(lc.bn.bias), frozen=ctx.frozen)
            gw = grad_of(lc.conv.weight)
            if lc.is_c1:
                if gw is not None:
                    ops.conv_c1_wgrad(lc.x_in, dy, gw)
                dx = ops.conv_c1_dgrad(dy, lc.conv.weight).unsqueeze(1)
            else:
                if gw is not None:
                    tcconv.conv3x3_wgrad(dy, lc.x_in, out=gw)
                dA = tcconv.conv3x3(dy, lc.wd, lc.conv.weight.shape[1])
                dx = dA
        if ctx.cin != 1:
            dx = dx.float().permute(0, 3, 1, 2)
        blk = l1
        return (None, dx, None, None, None, grads.get(l1.conv.weight), grads.get(l2.conv.weight),
                grads.get(l1.bn.weight), grads.get(l1.bn.bias), grads.get(l2.bn.weight), grads.get(l2.bn.bias))


def conv_block_forward(block, input, pool_size=(2, 2), pool_type='avg'):
    """ConvBlock.forward (models.py:99-115) on an NCHW fp32 CUDA tensor.  Only pool_type='avg' --
    the only mode any reference model uses -- runs on the fused BN+ReLU+pool kernel."""
    if pool_type != 'avg':
        if pool_type in ('max', 'avg+max'):
            raise NotImplementedError("pool_type=%r is never used by the reference models and is not "
                                      "implemented on the B200 path" % pool_type)
        raise Exception('Incorrect argument!')
    if not input.is_cuda:
        raise RuntimeError('ConvBlock: CUDA tensor required (no CPU path in this package)')
    ph, pw = pool_size
    keep = torch.is_grad_enabled() and (input.requires_grad or any(
        p.requires_grad for p in (block.conv1.weight, block.conv2.weight, block.bn1.weight, block.bn2.weight)))
    return _ConvBlockFn.apply(block, input, ph, pw, keep, block.conv1.weight, block.conv2.weight, block.bn1.weight,
                              block.bn1.bias, block.bn2.weight, block.bn2.bias)


class _AttBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ab, x, wa, ba, wc, bc):
        b, c, t = x.shape
        k = wa.shape[0]
        feat2d = x.transpose(1, 2).contiguous().view(b * t, c).float()
        att_logit = ops.linear_small_fwd(feat2d, wa.view(k, c), ba).view(b, t, k)
        cla_logit = ops.linear_small_fwd(feat2d, wc.view(k, c), bc).view(b, t, k)
        clip, norm_att, cla, _ = ops.head_att_fwd(att_logit, cla_logit, 1, ab.activation == 'sigmoid',
                                                  ab.temperature, want_frame=False)
        ctx.ab = ab
        ctx.save_for_backward(feat2d, att_logit, norm_att, cla, clip)
        ctx.shape = (b, c, t, k)
        ctx.mark_non_differentiable(norm_att, cla)
        return clip, norm_att, cla

    @staticmethod
    def backward(ctx, dclip, _dn, _dc):
        feat2d, att_logit, norm_att, cla, clip = ctx.saved_tensors
        b, c, t, k = ctx.shape
        ab = ctx.ab
        d_att, d_cla = ops.head_att_bwd(att_logit, norm_att, cla, clip, dclip.float().contiguous(),
                                        ab.activation == 'sigmoid', ab.temperature)
        gwa, gba = torch.empty(k, c, device=clip.device), torch.empty(k, device=clip.device)
        gwc, gbc = torch.empty(k, c, device=clip.device), torch.empty(k, device=clip.device)
        dx = ops.linear_small_bwd(d_att.view(b * t, k), feat2d, ab.att.weight.view(k, c), gwa, gba)
        dx = ops.linear_small_bwd(d_cla.view(b * t, k), feat2d, ab.cla.weight.view(k, c), gwc, gbc, dx=dx)
        return None, dx.view(b, t, c).transpose(1, 2), gwa.view(k, c, 1), gba, gwc.view(k, c, 1), gbc


def att_block_forward(ab, x):
    """AttBlock.forward (models.py:135-143): x (B, n_in, T) -> (clip (B,K), norm_att (B,K,T), cla (B,K,T)).
    Gradients flow through the clip-level output."""
    if not x.is_cuda:
        raise RuntimeError('AttBlock: CUDA tensor required (no CPU path in this package)')
    return _AttBlockFn.apply(ab, x, ab.att.weight, ab.att.bias, ab.cla.weight, ab.cla.bias)


def multihead_forward(mh, q, k, v, mask=None):
    from . import attention
    return attention.multihead_module_forward(mh, q, k, v, mask)
