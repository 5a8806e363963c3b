"""losses drop-in (/root/reference/pytorch/losses.py): clip_bce / get_loss_func.  The loss is a
torch.autograd.Function over the fused BCE kernel (forward value + analytic gradient in one launch)."""
import torch

from . import ops


class _ClipBCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, target):
        loss, dprob = ops.bce(prob.float(), target.float(), want_grad=True)
        ctx.save_for_backward(dprob)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dprob,) = ctx.saved_tensors
        return dprob * dloss, None


def clip_bce(output_dict, target_dict):
    """Binary cross entropy of clipwise_output (N, classes) against target (N, classes), mean reduced."""
    prob, target = output_dict['clipwise_output'], target_dict['target']
    if not prob.is_cuda:
        raise RuntimeError('clip_bce: CUDA tensors required (no CPU path in this package)')
    return _ClipBCE.apply(prob, target)


def get_loss_func(loss_type):
    if loss_type == 'clip_bce':
        return clip_bce
