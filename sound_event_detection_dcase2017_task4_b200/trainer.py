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
import math

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
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, world_size=1, use_graph=False):
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
        # Optional CUDA-graph replay of the step (models without device-side dropout): one graph per set of input
        # buffers; the per-step host inputs (SpecAugment stripes, Adam bias corrections) go through small static device
        # tensors that are refreshed before every replay.  With world_size > 1 the graph ends before the collective
        # (capturing the NCCL all-reduce inside the two-stream graph hung in a 2-GPU trial): forward + backward replay
        # as one graph, then the all-reduce and the one Adam kernel are launched eagerly.
        # The multi-head models' dropout draws (seed, offset) from torch's CUDA generator on the host: under replay the
        # kernels read them from two device words instead (attention.PhiloxIndirect), refreshed before every replay.
        self.use_graph = bool(use_graph)
        self._philox = None
        self._graphs = {}
        self._pool = None
        self._eager_steps = 0
        self.graph_launches = 0           # libsedb200 launches captured in one graph (bench.py: gpu_launches)

    def grad_of(self, p):
        if p is None or not p.requires_grad:
            return None
        return self._grad[p]

    def step(self, wave, target, lam):
        """wave (B2, L) fp32/int16, target (B2, K) fp32, lam (B2,) fp32 or None (all CUDA).
        Returns the rank-local loss as a 0-d CUDA tensor (no host sync here).  In graph mode the returned tensor is
        a static buffer of the graph: read it before the next step."""
        if self.use_graph and self._eager_steps >= 2:
            return self._step_graph(wave, target, lam)
        self._eager_steps += 1
        return self._step_body(wave, target, lam)

    def _step_body(self, wave, target, lam, stripes=None, bias_corr=None):
        loss = self._forward_backward(wave, target, lam, stripes)
        self._update(bias_corr)
        return loss

    def _update(self, bias_corr=None):
        """The collective and the optimizer: all-reduce of the flat gradient, then Adam(amsgrad) over the flat buffers."""
        with torch.no_grad():
            exchange_gradients(self.flat_grad, self.world_size, self.group)
            self.step_count += 1
            ops.adam_amsgrad_(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq,
                              self.max_exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                              self.step_count, bias_corr=bias_corr)

    def _forward_backward(self, wave, target, lam, stripes=None):
        model = self.model
        with torch.no_grad():
            feat, tctx = engine.trunk_forward(model, wave, lam, True, stripes=stripes)
            feat, mctx = temporal.forward(model, feat, True, keep=True)
            out, hctx = engine.head_forward(model, feat, model.interpolate_ratio, want_frame=True, keep=True)
            tgt = ops.mix_pairs(target, lam) if lam is not None else target
            loss, dprob = ops.bce(out['clipwise_output'], tgt, want_grad=True,
                                  grad_scale=1.0 / self.world_size)
            dfeat = engine.head_backward(model, hctx, dprob, self.grad_of)
            dfeat = temporal.backward(model, mctx, dfeat, self.grad_of)
            engine.trunk_backward(tctx, dfeat, self.grad_of)
        self.last_output = out
        return loss

    def _step_graph(self, wave, target, lam):
        from . import _lib, specaug
        import numpy as np
        model = self.model
        dev = self.flat_param.device
        b2, n = wave.shape
        key = (wave.data_ptr(), target.data_ptr(), 0 if lam is None else lam.data_ptr(), tuple(wave.shape), wave.dtype)
        aug = model.spec_augmenter
        hop = model.spectrogram_extractor.stft.hop_length
        t, m = n // hop + 1, model.bn0.num_features
        ts, fs = specaug.draw_spec_augment(b2, t, m, aug.time_dropper.drop_width, aug.time_dropper.stripes_num,
                                           aug.freq_dropper.drop_width, aug.freq_dropper.stripes_num)
        step = self.step_count + 1
        b1, b2 = float(np.float32(self.betas[0])), float(np.float32(self.betas[1]))   # the C ABI takes the betas as float
        bc = np.array([1.0 - math.pow(b1, step), math.sqrt(1.0 - math.pow(b2, step))],
                      dtype=np.float32)                  # the doubles csrc/adam.cu computes, rounded to float like there
        entry = self._graphs.get(key)
        if entry is None:
            entry = {'ts': torch.empty(ts.shape, dtype=torch.int32, device=dev),
                     'fs': torch.empty(fs.shape, dtype=torch.int32, device=dev),
                     'bc': torch.empty(2, dtype=torch.float32, device=dev)}
        # pageable -> device copies are staged by the driver before the call returns: the host arrays may be reused
        entry['ts'].copy_(torch.from_numpy(ts))
        entry['fs'].copy_(torch.from_numpy(fs))
        entry['bc'].copy_(torch.from_numpy(bc))
        mha = getattr(model, 'temporal_kind', None) == 'mha'
        if mha:
            from . import attention
            if self._philox is None:
                self._philox = attention.PhiloxIndirect(dev)
            gen = attention._generator(dev)
            seed, off = gen.initial_seed(), gen.get_offset()
            self._philox.state.copy_(torch.tensor([seed - (1 << 64) if seed >= (1 << 63) else seed, off // 4],
                                                  dtype=torch.int64))
        if 'graph' not in entry:
            n0 = _lib.launch_count()
            count0 = self.step_count
            g = torch.cuda.CUDAGraph()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()   # all graphs of this trainer share one memory pool: they
            if mha:
                attention.INDIRECT, self._philox.blocks = self._philox, 0
            try:
                with torch.cuda.graph(g, pool=self._pool):    # are replayed one at a time and their outputs are read first
                    if self.world_size == 1:
                        loss = self._step_body(wave, target, lam, stripes=(entry['ts'], entry['fs']), bias_corr=entry['bc'])
                    else:
                        loss = self._forward_backward(wave, target, lam, stripes=(entry['ts'], entry['fs']))
            finally:
                if mha:
                    attention.INDIRECT = None
            entry['philox_blocks'] = self._philox.blocks if mha else 0
            self.step_count = count0                       # capture executed nothing
            entry['graph'], entry['loss'], entry['out'] = g, loss, self.last_output
            self.graph_launches = _lib.launch_count() - n0
            self._graphs[key] = entry
        entry['graph'].replay()
        if mha:                                            # the draws of this step, as the eager path would have advanced
            gen.set_offset(off + 4 * entry['philox_blocks'])
        if self.world_size == 1:
            self.step_count += 1
        else:
            self._update()
        self.last_output = entry['out']
        return entry['loss']
