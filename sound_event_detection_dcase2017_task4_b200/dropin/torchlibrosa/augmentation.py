"""torchlibrosa.augmentation drop-in: SpecAugmentation / DropStripes (ctor site
/root/reference/pytorch/models.py:176-177, call :206-207).

Same RNG protocol as upstream (bit-exact stripe indices from the torch CPU generator, see
sound_event_detection_dcase2017_task4_b200/specaug.py); the stripes are zeroed in place by ONE
kernel instead of 2*B slice-fill launches per dropper.
"""
import torch
import torch.nn as nn

from sound_event_detection_dcase2017_task4_b200 import ops as _ops
from sound_event_detection_dcase2017_task4_b200 import specaug as _sa


class _DropStripesFn(torch.autograd.Function):
    """In-place stripe zeroing that autograd can see.  Upstream assigns ``e[:, bgn:bgn+d, :] = 0`` -- a tracked
    in-place op, so the gradient is zeroed inside the stripes as well (the stripes sit between bn0 and the first
    convolution: without this, bn0.weight / bn0.bias would collect gradient from masked positions)."""

    @staticmethod
    def forward(ctx, input, table, dim):
        ctx.mark_dirty(input)
        ctx.save_for_backward(table)
        ctx.dim = dim
        if dim == 2:
            _ops.spec_augment_(input, table, None)
        else:
            _ops.spec_augment_(input, None, table)
        return input

    @staticmethod
    def backward(ctx, grad_output):
        (table,) = ctx.saved_tensors
        g = grad_output.contiguous().clone()
        if g.dtype != torch.float32:
            raise RuntimeError('DropStripes backward: expected a float32 gradient')
        if ctx.dim == 2:
            _ops.spec_augment_(g, table, None)
        else:
            _ops.spec_augment_(g, None, table)
        return g, None, None


class DropStripes(nn.Module):
    def __init__(self, dim, drop_width, stripes_num):
        super().__init__()
        assert dim in [2, 3]
        self.dim, self.drop_width, self.stripes_num = dim, drop_width, stripes_num

    def forward(self, input):
        """input: (batch_size, channels, time_steps, freq_bins); modified in place when training."""
        assert input.ndimension() == 4
        if not self.training:
            return input
        if not input.is_cuda:
            raise RuntimeError('DropStripes: CUDA tensor required (no CPU path in this package)')
        stripes = _sa.draw_stripes(input.shape[0], input.shape[self.dim], self.drop_width, self.stripes_num)
        table = torch.from_numpy(stripes).to(input.device, non_blocking=True)
        if not (input.is_contiguous() and input.dtype == torch.float32):
            raise RuntimeError('DropStripes: expected a contiguous float32 tensor')
        if torch.is_grad_enabled() and input.requires_grad:
            return _DropStripesFn.apply(input, table, self.dim)
        if self.dim == 2:
            _ops.spec_augment_(input, table, None)
        else:
            _ops.spec_augment_(input, None, table)
        return input


class SpecAugmentation(nn.Module):
    def __init__(self, time_drop_width, time_stripes_num, freq_drop_width, freq_stripes_num):
        super().__init__()
        self.time_dropper = DropStripes(dim=2, drop_width=time_drop_width, stripes_num=time_stripes_num)
        self.freq_dropper = DropStripes(dim=3, drop_width=freq_drop_width, stripes_num=freq_stripes_num)

    def forward(self, input):
        return self.freq_dropper(self.time_dropper(input))
