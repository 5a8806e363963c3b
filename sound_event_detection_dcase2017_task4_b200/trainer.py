"""The per-step training path of /root/reference/pytorch/main.py:233-258 without autograd:

    lam -> model(waveform, lam) -> do_mixup(target) -> clip_bce -> backward -> Adam(amsgrad).step()

sequenced as libsedb200 launches on torch's current stream (engine.py / temporal.py), with

* every trainable parameter a view into ONE flat fp32 buffer, every gradient a view into a second
  flat buffer of the same layout (dead parameters -- AttBlock.bn_att, MultiHead.layer_norm,
  SURVEY.md 5.9 -- keep zero gradient slots, which leaves them untouched exactly like torch's
  Adam skipping ``grad is None``);
* data parallelism as ONE NCCL all-reduce of the flat gradient buffer per step (replaces
  nn.DataParallel's per-step parameter broadcast + gradient reduce, main.py:138): each rank
  scales its loss gradient by 1/world so the sum is the global-batch mean; BatchNorm statistics
  stay per rank, which is DataParallel's per-replica behaviour;
* Adam + amsgrad as one fused kernel over the flat buffers (main.py:144-145, :258).

The module classes in models.py give the same arithmetic through torch.autograd for the
reference's unmodified main.py; this class is what bench.py times.
"""
import torch

from . import engine
from . import ops
from . import temporal


def shard_bounds(n_items, world_size, rank):
    """Contiguous shard [lo, hi) of ``n_items`` raw clips for ``rank``: equal, EVEN-sized shards so that the mixup
    pairs (2i, 2i+1) of pytorch_utils.do_mixup never straddle two ranks."""
    if n_items % (2 * world_size) != 0:
        raise ValueError('%d clips do not split into %d shards of whole mixup pairs' % (n_items, world_size))
    per = n_items // world_size
    return rank * per, (rank + 1) * per


def exchange_gradients(flat_grad, world_size, group=None):
    """The ONE collective of the step (replaces nn.DataParallel's reduce of per-replica gradients, main.py:138): every
    rank has already scaled its loss gradient by 1/world_size, so the SUM over ranks is the gradient of the global-batch
    mean loss.  Works on any backend torch.distributed offers (NCCL on the GPUs; gloo in the CPU tests)."""
    if world_size > 1:
        torch.distributed.all_reduce(flat_grad, op=torch.distributed.ReduceOp.SUM, group=group)
    return flat_grad


class FusedTrainer(object):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, world_size=1):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.world_size = world_size
        self.group = process_group
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FusedTrainer: the model must live on a CUDA device (no CPU path)')
        total = sum(p.numel() for p in self.params)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat_grad)
        self.exp_avg_sq = torch.zeros_like(self.flat_grad)
        self.max_exp_avg_sq = torch.zeros_like(self.flat_grad)
        self._grad = {}
        off = 0
        for p in self.params:
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view(p.shape)
            self._grad[p] = self.flat_grad[off:off + n].view(p.shape)
            off += n
        self.step_count = 0

    def grad_of(self, p):
        if p is None or not p.requires_grad:
            return None
        return self._grad[p]

    def step(self, wave, target, lam):
        """wave (B2, L) fp32/int16, target (B2, K) fp32, lam (B2,) fp32 or None (all CUDA).
        Returns the rank-local loss as a 0-d CUDA tensor (no host sync here)."""
        model = self.model
        with torch.no_grad():
            feat, tctx = engine.trunk_forward(model, wave, lam, True)
            feat, mctx = temporal.forward(model, feat, True, keep=True)
            out, hctx = engine.head_forward(model, feat, model.interpolate_ratio, want_frame=True, keep=True)
            tgt = ops.mix_pairs(target, lam) if lam is not None else target
            loss, dprob = ops.bce(out['clipwise_output'], tgt, want_grad=True,
                                  grad_scale=1.0 / self.world_size)
            dfeat = engine.head_backward(model, hctx, dprob, self.grad_of)
            dfeat = temporal.backward(model, mctx, dfeat, self.grad_of)
            engine.trunk_backward(tctx, dfeat, self.grad_of)
            exchange_gradients(self.flat_grad, self.world_size, self.group)
            self.step_count += 1
            ops.adam_amsgrad_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq,
                              self.max_exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                              self.step_count)
        self.last_output = out
        return loss
