"""CPU restatement (PyTorch fp32) of the reference's model zoo, loss and mixup.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against the unmodified reference
modules by tests/golden/make_golden.py + tests/test_oracle_golden.py.

Follows, function by function:
  xavier/bn/gru inits      /root/reference/pytorch/models.py:15-55
  interpolate              /root/reference/pytorch/models.py:58-69
  ConvBlock                /root/reference/pytorch/models.py:72-115
  AttBlock                 /root/reference/pytorch/models.py:118-149
  MultiHead (+SDPA)        /root/reference/pytorch/models.py:587-665
  Cnn_9layers_* (7)        /root/reference/pytorch/models.py:152-853
  clip_bce                 /root/reference/pytorch/losses.py:5-17
  do_mixup                 /root/reference/pytorch/pytorch_utils.py:80-93
  Mixup.get_lambda         /root/reference/utils/utilities.py:220-242
  train step               /root/reference/pytorch/main.py:233-258

The seven reference classes differ only in the temporal module between the CNN trunk and
the classifier and in how clip-level probabilities are pooled, so they are restated here as
ONE trunk (``Cnn9``) plus a (temporal, pooling) table.  Sub-module names, creation order and
init order follow the reference so that ``state_dict`` keys and the torch RNG stream under
``torch.manual_seed`` are identical (that is what lets golden outputs be compared without
shipping weights).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .frontend import Spectrogram, LogmelFilterBank, SpecAugmentation

SPECS = {
    # name: (temporal module, pooling head)
    'Cnn_9layers_FrameMax': (None, 'max'),
    'Cnn_9layers_FrameAvg': (None, 'avg'),
    'Cnn_9layers_FrameAtt': (None, 'att'),
    'Cnn_9layers_Gru_FrameAvg': ('gru', 'avg'),
    'Cnn_9layers_Gru_FrameAtt': ('gru', 'att'),
    'Cnn_9layers_Transformer_FrameAvg': ('mha', 'avg'),
    'Cnn_9layers_Transformer_FrameAtt': ('mha', 'att'),
}


def xavier_(layer):
    nn.init.xavier_uniform_(layer.weight)
    if getattr(layer, 'bias', None) is not None:
        layer.bias.data.zero_()


def bn_identity_(bn):
    bn.bias.data.zero_()
    bn.weight.data.fill_(1.0)


def gru_init_(rnn):
    """Per-gate uniform(+-sqrt(3/fan_in)) on every gate block except the hidden-to-hidden
    candidate gate, which is orthogonal; all biases zero.  Order: ih (r,z,n), bias_ih,
    hh (r,z,n), bias_hh for the forward direction only -- the reference's loop runs over
    ``num_layers`` and never touches the ``_reverse`` tensors (models.py:44-55)."""
    def uniform_block(t):
        bound = math.sqrt(3.0 / t.shape[1])
        nn.init.uniform_(t, -bound, bound)

    for layer in range(rnn.num_layers):
        w_ih = getattr(rnn, 'weight_ih_l%d' % layer)
        w_hh = getattr(rnn, 'weight_hh_l%d' % layer)
        h = w_ih.shape[0] // 3
        for g in range(3):
            uniform_block(w_ih[g * h:(g + 1) * h, :])
        nn.init.constant_(getattr(rnn, 'bias_ih_l%d' % layer), 0)
        uniform_block(w_hh[0:h, :])
        uniform_block(w_hh[h:2 * h, :])
        nn.init.orthogonal_(w_hh[2 * h:3 * h, :])
        nn.init.constant_(getattr(rnn, 'bias_hh_l%d' % layer), 0)


def repeat_frames(x, ratio):
    """(B, T, C) -> (B, T*ratio, C); frame t_out is a bit-identical copy of t_out // ratio."""
    b, t, c = x.shape
    return x.unsqueeze(2).expand(b, t, ratio, c).reshape(b, t * ratio, c)


def mix_pairs(x, lam):
    """out[i] = x[2i]*lam[2i] + x[2i+1]*lam[2i+1] along dim 0."""
    shape = (-1,) + (1,) * (x.dim() - 1)
    return x[0::2] * lam[0::2].reshape(shape) + x[1::2] * lam[1::2].reshape(shape)


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), stride=(1, 1),
                               padding=(1, 1), bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), stride=(1, 1),
                               padding=(1, 1), bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.bn2 = nn.BatchNorm2d(out_channels)
        xavier_(self.conv1)
        xavier_(self.conv2)
        bn_identity_(self.bn1)
        bn_identity_(self.bn2)

    def forward(self, input, pool_size=(2, 2), pool_type='avg'):
        x = torch.relu(self.bn1(self.conv1(input)))
        x = torch.relu(self.bn2(self.conv2(x)))
        if pool_type == 'avg':
            return F.avg_pool2d(x, kernel_size=pool_size)
        if pool_type == 'max':
            return F.max_pool2d(x, kernel_size=pool_size)
        if pool_type == 'avg+max':
            return F.avg_pool2d(x, kernel_size=pool_size) + F.max_pool2d(x, kernel_size=pool_size)
        raise Exception('Incorrect argument!')


class AttBlock(nn.Module):
    def __init__(self, n_in, n_out, activation='linear', temperature=1.):
        super().__init__()
        self.activation = activation
        self.temperature = temperature
        self.att = nn.Conv1d(n_in, n_out, kernel_size=1, bias=True)
        self.cla = nn.Conv1d(n_in, n_out, kernel_size=1, bias=True)
        self.bn_att = nn.BatchNorm1d(n_out)          # registered, never used (models.py:127)
        xavier_(self.att)
        xavier_(self.cla)
        bn_identity_(self.bn_att)

    def forward(self, x):
        """x (B, n_in, T) -> clip (B, n_out), norm_att (B, n_out, T), cla (B, n_out, T)."""
        e = torch.exp(torch.clamp(self.att(x), -10, 10) / self.temperature) + 1e-6
        norm_att = e / e.sum(dim=2, keepdim=True)
        cla = self.cla(x)
        if self.activation == 'sigmoid':
            cla = torch.sigmoid(cla)
        return (norm_att * cla).sum(dim=2), norm_att, cla


class _SDPA(nn.Module):
    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)


class MultiHead(nn.Module):
    """One multi-head self-attention layer: no residual, no LayerNorm applied (the
    ``layer_norm`` sub-module is registered but dead), output = relu(dropout(fc(heads)))."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_v)))
        for lin in (self.w_qs, self.w_ks, self.w_vs):
            lin.bias.data.zero_()
        self.attention = _SDPA(temperature=np.power(d_k, 0.5))
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.fc.bias.data.zero_()
        self.dropout = nn.Dropout(dropout)

    def forward(self, q, k, v, mask=None):
        b, lq, _ = q.shape
        h, dk, dv = self.n_head, self.d_k, self.d_v

        def split(t, lin, d):                       # (B, L, h*d) -> (h*B, L, d), head-major
            return lin(t).view(b, -1, h, d).permute(2, 0, 1, 3).reshape(h * b, -1, d)

        qh, kh, vh = split(q, self.w_qs, dk), split(k, self.w_ks, dk), split(v, self.w_vs, dv)
        score = torch.bmm(qh, kh.transpose(1, 2)) / self.attention.temperature
        if mask is not None:
            score = score.masked_fill(mask, -np.inf)
        attn = self.attention.dropout(torch.softmax(score, dim=2))
        ctx = torch.bmm(attn, vh).view(h, b, lq, dv).permute(1, 2, 0, 3).reshape(b, lq, h * dv)
        return torch.relu(self.dropout(self.fc(ctx)))


class Cnn9(nn.Module):
    """log-mel -> bn0 -> [SpecAug, mixup] -> 4 ConvBlocks -> freq-mean -> temporal -> head."""

    interpolate_ratio = 8

    def __init__(self, name, sample_rate, window_size, hop_size, mel_bins, fmin, fmax,
                 classes_num):
        super().__init__()
        self.temporal, self.pooling = SPECS[name]
        self.spectrogram_extractor = Spectrogram(
            n_fft=window_size, hop_length=hop_size, win_length=window_size, window='hann',
            center=True, pad_mode='reflect', freeze_parameters=True)
        self.logmel_extractor = LogmelFilterBank(
            sr=sample_rate, n_fft=window_size, n_mels=mel_bins, fmin=fmin, fmax=fmax, ref=1.0,
            amin=1e-10, top_db=None, freeze_parameters=True)
        self.spec_augmenter = SpecAugmentation(time_drop_width=64, time_stripes_num=2,
                                               freq_drop_width=8, freq_stripes_num=2)
        self.bn0 = nn.BatchNorm2d(64)
        self.conv_block1 = ConvBlock(1, 64)
        self.conv_block2 = ConvBlock(64, 128)
        self.conv_block3 = ConvBlock(128, 256)
        self.conv_block4 = ConvBlock(256, 512)
        if self.temporal == 'gru':
            self.gru = nn.GRU(input_size=512, hidden_size=256, num_layers=1, bias=True,
                              batch_first=True, bidirectional=True)
        elif self.temporal == 'mha':
            self.multihead = MultiHead(8, 512, 64, 64, 0.2)
        if self.pooling == 'att':
            self.att_block = AttBlock(n_in=512, n_out=17, activation='sigmoid')
        else:
            self.fc = nn.Linear(512, classes_num, bias=True)
        bn_identity_(self.bn0)
        if self.temporal == 'gru':
            gru_init_(self.gru)
        if self.pooling != 'att':
            xavier_(self.fc)

    def features(self, input, mixup_lambda=None):
        x = self.logmel_extractor(self.spectrogram_extractor(input))      # (B2, 1, T, mel)
        x = self.bn0(x.transpose(1, 3)).transpose(1, 3)
        if self.training:
            x = self.spec_augmenter(x)
            if mixup_lambda is not None:
                x = mix_pairs(x, mixup_lambda)
        x = self.conv_block1(x, pool_size=(2, 2), pool_type='avg')
        x = self.conv_block2(x, pool_size=(2, 2), pool_type='avg')
        x = self.conv_block3(x, pool_size=(2, 2), pool_type='avg')
        x = self.conv_block4(x, pool_size=(1, 1), pool_type='avg')
        return x.mean(dim=3)                                              # (B, 512, T')

    def forward(self, input, mixup_lambda=None):
        x = self.features(input, mixup_lambda)
        if self.temporal == 'gru':
            x = self.gru(x.transpose(1, 2))[0].transpose(1, 2)
        elif self.temporal == 'mha':
            seq = x.transpose(1, 2)
            x = self.multihead(seq, seq, seq).transpose(1, 2)
        if self.pooling == 'att':
            clip, _, cla = self.att_block(x)
            frame = repeat_frames(cla.transpose(1, 2), self.interpolate_ratio)
            embedding = x if self.temporal == 'mha' else cla
        else:
            frame = repeat_frames(torch.sigmoid(self.fc(x.transpose(1, 2))),
                                  self.interpolate_ratio)
            clip = frame.max(dim=1)[0] if self.pooling == 'max' else frame.mean(dim=1)
            embedding = x
        return {'framewise_output': frame, 'clipwise_output': clip, 'embedding': embedding}


def build(name, sample_rate=32000, window_size=1024, hop_size=320, mel_bins=64, fmin=50,
          fmax=14000, classes_num=17):
    return Cnn9(name, sample_rate, window_size, hop_size, mel_bins, fmin, fmax, classes_num)


def clip_bce(output_dict, target_dict):
    return F.binary_cross_entropy(output_dict['clipwise_output'], target_dict['target'])


class MixupLambda(object):
    """numpy-identical lambda stream: per pair one ``RandomState(seed).beta(a, a, 1)[0]``,
    emitted as [lam, 1-lam, ...] float64."""

    def __init__(self, mixup_alpha, random_seed=1234):
        self.alpha = mixup_alpha
        self.rs = np.random.RandomState(random_seed)

    def get_lambda(self, batch_size):
        out = np.empty(batch_size + (batch_size & 1), dtype=np.float64)
        for i in range(0, batch_size, 2):
            lam = self.rs.beta(self.alpha, self.alpha, 1)[0]
            out[i] = lam
            out[i + 1] = 1.0 - lam
        return out


def int16_to_float32(x):
    return (x / 32767.).astype(np.float32)


def synthetic_batch(n_clips, n_samples=320000, classes_num=17, seed=1234):
    """SURVEY.md section 8d synthetic inputs: int16 uniform [-8192, 8191] -> /32767 -> fp32;
    Bernoulli(0.067) targets."""
    rs = np.random.RandomState(seed)
    pcm = rs.randint(-8192, 8192, size=(n_clips, n_samples)).astype(np.int16)
    target = (rs.rand(n_clips, classes_num) < 0.067).astype(np.float32)
    return pcm, int16_to_float32(pcm), target


def train_step(model, optimizer, waveform, target, mixup_lambda=None):
    """One iteration of the loop body at main.py:233-258 (no printing).  Returns the loss."""
    model.train()
    if mixup_lambda is not None:
        out = model(waveform, mixup_lambda)
        tgt = {'target': mix_pairs(target, mixup_lambda)}
    else:
        out = model(waveform, None)
        tgt = {'target': target}
    loss = clip_bce(out, tgt)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss
