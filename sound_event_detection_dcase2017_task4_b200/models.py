"""Drop-in model zoo (seam B): same class names, constructor signatures, sub-module names,
state_dict keys and torch-RNG consumption as /root/reference/pytorch/models.py, with every
forward/backward computed by libsedb200 kernels (engine.py).

    Model = eval(model_type); Model(sample_rate, window_size, hop_size, mel_bins, fmin, fmax,
                                    classes_num)                     (pytorch/main.py:120-123)
    model(waveform, mixup_lambda) -> {'framewise_output', 'clipwise_output', 'embedding'}

The nn.Conv2d / nn.BatchNorm2d / nn.Linear / nn.GRU children are *parameter containers*:
they give identical parameter registration and initialisation (models.py:15-55) but their
own forward is never called.  The whole forward runs inside ONE torch.autograd.Function whose
backward launches the hand-written backward kernels, so `loss.backward()` / `optimizer.step()`
of the reference's main.py work unchanged.  Gradients flow to the parameters through
``clipwise_output`` only (that is what losses.clip_bce consumes); ``framewise_output`` and
``embedding`` are returned detached.  Activations are kept whenever grad is enabled -- in eval()
mode too, where the BatchNorm backward treats the running statistics as constants.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import engine
from . import temporal
from .dropin.torchlibrosa.stft import Spectrogram, LogmelFilterBank
from .dropin.torchlibrosa.augmentation import SpecAugmentation

__all__ = ['init_layer', 'init_bn', 'init_gru', 'interpolate', 'ConvBlock', 'AttBlock',
           'ScaledDotProductAttention', 'MultiHead', 'Cnn_9layers_FrameMax', 'Cnn_9layers_FrameAvg',
           'Cnn_9layers_FrameAtt', 'Cnn_9layers_Gru_FrameAvg', 'Cnn_9layers_Gru_FrameAtt',
           'Cnn_9layers_Transformer_FrameAvg', 'Cnn_9layers_Transformer_FrameAtt']


# ------------------------------------------------------------------ initialisers (models.py:15-55)
def init_layer(layer):
    nn.init.xavier_uniform_(layer.weight)
    if getattr(layer, 'bias', None) is not None:
        layer.bias.data.fill_(0.)


def init_bn(bn):
    bn.bias.data.fill_(0.)
    bn.weight.data.fill_(1.)


def init_gru(rnn):
    def gate_uniform(block):
        bound = math.sqrt(3.0 / block.shape[1])
        nn.init.uniform_(block, -bound, bound)

    for i in range(rnn.num_layers):
        w_ih, w_hh = getattr(rnn, 'weight_ih_l%d' % i), getattr(rnn, 'weight_hh_l%d' % i)
        hid = w_ih.shape[0] // 3
        for g in range(3):
            gate_uniform(w_ih[g * hid:(g + 1) * hid, :])
        nn.init.constant_(getattr(rnn, 'bias_ih_l%d' % i), 0)
        gate_uniform(w_hh[0:hid, :])
        gate_uniform(w_hh[hid:2 * hid, :])
        nn.init.orthogonal_(w_hh[2 * hid:3 * hid, :])
        nn.init.constant_(getattr(rnn, 'bias_hh_l%d' % i), 0)


def interpolate(x, ratio):
    """(B, T, C) -> (B, T*ratio, C): frame t_out is an exact copy of frame t_out // ratio."""
    b, t, c = x.shape
    return x[:, :, None, :].expand(b, t, ratio, c).reshape(b, t * ratio, c)


# ------------------------------------------------------------------ building blocks
class ConvBlock(nn.Module):
    """Parameter container for conv1/bn1/conv2/bn2 (models.py:72-96).  Used stand-alone it runs
    the same kernels through the functional API in blocks.py."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1),
                               bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1),
                               bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.init_weight()

    def init_weight(self):
        init_layer(self.conv1)
        init_layer(self.conv2)
        init_bn(self.bn1)
        init_bn(self.bn2)

    def forward(self, input, pool_size=(2, 2), pool_type='avg'):
        from . import blocks
        return blocks.conv_block_forward(self, input, pool_size, pool_type)


class AttBlock(nn.Module):
    def __init__(self, n_in, n_out, activation='linear', temperature=1.):
        super().__init__()
        self.activation = activation
        self.temperature = temperature
        self.att = nn.Conv1d(n_in, n_out, kernel_size=1, stride=1, padding=0, bias=True)
        self.cla = nn.Conv1d(n_in, n_out, kernel_size=1, stride=1, padding=0, bias=True)
        self.bn_att = nn.BatchNorm1d(n_out)               # registered but unused, as in the reference
        self.init_weights()

    def init_weights(self):
        init_layer(self.att)
        init_layer(self.cla)
        init_bn(self.bn_att)

    def forward(self, x):
        from . import blocks
        return blocks.att_block_forward(self, x)


class ScaledDotProductAttention(nn.Module):
    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)
        self.softmax = nn.Softmax(dim=2)


class MultiHead(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        self.w_qs.bias.data.fill_(0)
        self.w_ks.bias.data.fill_(0)
        self.w_vs.bias.data.fill_(0)
        self.attention = ScaledDotProductAttention(temperature=np.power(d_k, 0.5))
        self.layer_norm = nn.LayerNorm(d_model)            # registered but unused, as in the reference
        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.fc.bias.data.fill_(0)
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, mask=None):
        from . import blocks
        return blocks.multihead_forward(self, q, k, v, mask)


# ------------------------------------------------------------------ the fused model function
def trainable_tensors(module):
    """The tensors that carry gradient for ``module``, in ``module.parameters()`` order.

    ``nn.DataParallel`` (pytorch/main.py:138) runs forward on *replicas* whose ``_parameters`` dicts are
    empty: ``replicate()`` re-attaches every weight as a plain, non-leaf tensor attribute (the output of its
    Broadcast autograd node) and records it in ``_former_parameters``.  ``replica.parameters()`` therefore
    yields nothing, while ``replica.conv_block1.conv1.weight`` is the tensor the gradient must flow to.  This
    walk reads whichever table the module instance really has, so the same code serves the original module
    (leaf Parameters) and its replicas (broadcast copies)."""
    out, seen = [], set()
    for m in module.modules():
        table = m._parameters
        if getattr(m, '_is_replica', False) and getattr(m, '_former_parameters', None):
            table = m._former_parameters
        for p in table.values():
            if p is not None and p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                out.append(p)
    return out


class _SedFunction(torch.autograd.Function):
    """forward: waveform -> (clipwise, framewise, embedding) with libsedb200 kernels;
    backward: d clipwise -> gradients of every trainable parameter."""

    @staticmethod
    def forward(ctx, model, wave, lam, need_grad, *params):
        training = model.training          # (grad mode is always off inside Function.forward)
        feat, tctx = engine.trunk_forward(model, wave, lam if training else None, training, keep=need_grad)
        feat, mctx = temporal.forward(model, feat, training, keep=need_grad)
        out, hctx = engine.head_forward(model, feat, model.interpolate_ratio, want_frame=True, keep=need_grad)
        ctx.model, ctx.params = model, params
        ctx.saved = (tctx, mctx, hctx) if need_grad else None
        frame, emb = out['framewise_output'], out['embedding']
        if model.temporal_kind == 'mha':
            emb = feat.transpose(1, 2)          # models.py:838: the Transformer variants return the MHA output
        ctx.mark_non_differentiable(frame, emb)
        return out['clipwise_output'], frame, emb

    @staticmethod
    def backward(ctx, dclip, _dframe, _demb):
        if ctx.saved is None:
            raise RuntimeError('backward through a forward that kept no activations (it ran under torch.no_grad() or '
                               'was already backpropagated once)')
        tctx, mctx, hctx = ctx.saved
        ctx.saved = None
        grads = {}

        def grad_of(p):
            if p is None or not p.requires_grad:
                return None
            g = grads.get(p)
            if g is None:
                g = torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
                grads[p] = g
            return g

        model = ctx.model
        dfeat = engine.head_backward(model, hctx, dclip.float(), grad_of)
        dfeat = temporal.backward(model, mctx, dfeat, grad_of)
        engine.trunk_backward(tctx, dfeat, grad_of)
        return (None, None, None, None) + tuple(grads.get(p) for p in ctx.params)


class _Cnn9(nn.Module):
    """Shared trunk + (temporal, pooling) variants; sub-module creation and init order follow
    the seven reference classes so state_dict keys and RNG consumption are identical."""

    temporal_kind = None      # None | 'gru' | 'mha'
    pooling = 'avg'           # 'max' | 'avg' | 'att'
    interpolate_ratio = 8

    def __init__(self, sample_rate, window_size, hop_size, mel_bins, fmin, fmax, classes_num):
        super().__init__()
        self.spectrogram_extractor = Spectrogram(n_fft=window_size, hop_length=hop_size, win_length=window_size,
                                                 window='hann', center=True, pad_mode='reflect',
                                                 freeze_parameters=True)
        self.logmel_extractor = LogmelFilterBank(sr=sample_rate, n_fft=window_size, n_mels=mel_bins, fmin=fmin,
                                                 fmax=fmax, ref=1.0, amin=1e-10, top_db=None,
                                                 freeze_parameters=True)
        self.spec_augmenter = SpecAugmentation(time_drop_width=64, time_stripes_num=2, freq_drop_width=8,
                                               freq_stripes_num=2)
        self.bn0 = nn.BatchNorm2d(64)
        self.conv_block1 = ConvBlock(in_channels=1, out_channels=64)
        self.conv_block2 = ConvBlock(in_channels=64, out_channels=128)
        self.conv_block3 = ConvBlock(in_channels=128, out_channels=256)
        self.conv_block4 = ConvBlock(in_channels=256, out_channels=512)
        if self.temporal_kind == 'gru':
            self.gru = nn.GRU(input_size=512, hidden_size=256, num_layers=1, bias=True, batch_first=True,
                              bidirectional=True)
        elif self.temporal_kind == 'mha':
            self.multihead = MultiHead(8, 512, 64, 64, 0.2)
        if self.pooling == 'att':
            self.att_block = AttBlock(n_in=512, n_out=17, activation='sigmoid')
        else:
            self.fc = nn.Linear(512, classes_num, bias=True)
        self.init_weights()

    def init_weights(self):
        init_bn(self.bn0)
        if self.temporal_kind == 'gru':
            init_gru(self.gru)
        if self.pooling != 'att':
            init_layer(self.fc)

    def forward(self, input, mixup_lambda=None):
        """input: (batch_size, data_length) waveform; mixup_lambda: (batch_size,) or None."""
        if not input.is_cuda:
            raise RuntimeError('%s: CUDA tensors required -- this package has no CPU path'
                               % type(self).__name__)
        params = trainable_tensors(self)      # not self.parameters(): empty on DataParallel replicas
        # activations are kept whenever a gradient can be asked for -- also in eval() mode (frozen-BatchNorm
        # fine-tuning, saliency): the reference modules are differentiable there too
        need_grad = torch.is_grad_enabled() and len(params) > 0
        clip, frame, emb = _SedFunction.apply(self, input, mixup_lambda, need_grad, *params)
        return {'framewise_output': frame, 'clipwise_output': clip, 'embedding': emb}


class Cnn_9layers_FrameMax(_Cnn9):
    pooling = 'max'


class Cnn_9layers_FrameAvg(_Cnn9):
    pooling = 'avg'


class Cnn_9layers_FrameAtt(_Cnn9):
    pooling = 'att'


class Cnn_9layers_Gru_FrameAvg(_Cnn9):
    temporal_kind, pooling = 'gru', 'avg'


class Cnn_9layers_Gru_FrameAtt(_Cnn9):
    temporal_kind, pooling = 'gru', 'att'


class Cnn_9layers_Transformer_FrameAvg(_Cnn9):
    temporal_kind, pooling = 'mha', 'avg'


class Cnn_9layers_Transformer_FrameAtt(_Cnn9):
    temporal_kind, pooling = 'mha', 'att'
