"""Mint golden vectors from the UNMODIFIED reference (run in the authoring container only).

    python tests/golden/make_golden.py

Imports /root/reference/pytorch/{models,losses,pytorch_utils}.py and
/root/reference/utils/utilities.py as they are (empty stub modules for the absent
librosa/h5py/sed_eval/matplotlib; the oracle's ``torchlibrosa`` stand-in for the un-vendored
front-end package) and records their outputs on seeded synthetic inputs.  The reference
itself holds no tests or fixtures, so these files are the pin for
``oracle/sed.py``.  The front-end (torchlibrosa) remains *unpinned*: what is recorded for it
is the stand-in's output plus facts checkable against an independent float64 path.

/root/reference does not exist on the GPU box; only the .npz/.json written here travel.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'

MODEL_NAMES = ['Cnn_9layers_FrameMax', 'Cnn_9layers_FrameAvg', 'Cnn_9layers_FrameAtt',
               'Cnn_9layers_Gru_FrameAvg', 'Cnn_9layers_Gru_FrameAtt',
               'Cnn_9layers_Transformer_FrameAvg', 'Cnn_9layers_Transformer_FrameAtt']
CTOR = dict(sample_rate=32000, window_size=1024, hop_size=320, mel_bins=64, fmin=50, fmax=14000,
            classes_num=17)


def import_reference():
    for name in ('librosa', 'h5py', 'sed_eval', 'matplotlib', 'matplotlib.pyplot', 'autoth',
                 'autoth.core'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))          # torchlibrosa stand-in
    sys.path.insert(0, os.path.join(REF, 'utils'))
    sys.path.insert(0, os.path.join(REF, 'pytorch'))
    import models as ref_models
    import losses as ref_losses
    import pytorch_utils as ref_pu
    import utilities as ref_util
    return ref_models, ref_losses, ref_pu, ref_util


def synth(n_clips, n_samples, seed):
    rs = np.random.RandomState(seed)
    pcm = rs.randint(-8192, 8192, size=(n_clips, n_samples)).astype(np.int16)
    target = (rs.rand(n_clips, 17) < 0.067).astype(np.float32)
    return pcm, target


def main():
    torch.set_num_threads(8)
    ref_models, ref_losses, ref_pu, ref_util = import_reference()
    out = {}
    meta = {'reference_commit': '823af881', 'torch': torch.__version__, 'models': {}}

    # ---- Mixup lambda stream (utilities.py:220-242)
    mix = ref_util.Mixup(mixup_alpha=1., random_seed=1234)
    out['mixup_lambda_first32'] = mix.get_lambda(32)
    out['mixup_lambda_next6'] = mix.get_lambda(6)

    # ---- int16_to_float32 on every int16 value (utilities.py:66-67)
    all_pcm = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    out['int16_to_float32_all'] = ref_util.int16_to_float32(all_pcm)

    # ---- interpolate / do_mixup known answers (models.py:58-69, pytorch_utils.py:80-93)
    g = torch.Generator().manual_seed(7)
    xi = torch.randn(3, 5, 17, generator=g)
    out['interp_in'] = xi.numpy()
    out['interp_out'] = ref_models.interpolate(xi, 8).numpy()
    xm = torch.randn(6, 1, 11, 4, generator=g)
    lam = torch.Tensor(ref_util.Mixup(1., 1234).get_lambda(6))
    out['mixup_in'] = xm.numpy()
    out['mixup_lam'] = lam.numpy()
    out['mixup_out'] = ref_pu.do_mixup(xm, lam).numpy()
    tm = torch.rand(6, 17, generator=g)
    out['mixup_target_in'] = tm.numpy()
    out['mixup_target_out'] = ref_pu.do_mixup(tm, lam).numpy()

    # ---- ConvBlock / AttBlock / MultiHead module goldens (seeded init)
    torch.manual_seed(3)
    cb = ref_models.ConvBlock(in_channels=4, out_channels=8)
    xc = torch.randn(2, 4, 13, 10, generator=g)
    cb.train()
    out['convblock_in'] = xc.numpy()
    out['convblock_train_avg22'] = cb(xc, pool_size=(2, 2), pool_type='avg').detach().numpy()
    out['convblock_running_mean1'] = cb.bn1.running_mean.numpy().copy()
    out['convblock_running_var2'] = cb.bn2.running_var.numpy().copy()
    cb.eval()
    out['convblock_eval_avg11'] = cb(xc, pool_size=(1, 1), pool_type='avg').detach().numpy()
    out['convblock_eval_max22'] = cb(xc, pool_size=(2, 2), pool_type='max').detach().numpy()
    out['convblock_eval_avgmax22'] = cb(xc, pool_size=(2, 2), pool_type='avg+max').detach().numpy()

    torch.manual_seed(4)
    ab = ref_models.AttBlock(n_in=16, n_out=5, activation='sigmoid')
    xa = torch.randn(3, 16, 9, generator=g) * 3
    clip, natt, cla = ab(xa)
    out['attblock_in'] = xa.numpy()
    out['attblock_clip'] = clip.detach().numpy()
    out['attblock_norm_att'] = natt.detach().numpy()
    out['attblock_cla'] = cla.detach().numpy()

    torch.manual_seed(5)
    mh = ref_models.MultiHead(4, 32, 8, 8, 0.2)
    mh.eval()
    xq = torch.randn(2, 7, 32, generator=g)
    out['multihead_in'] = xq.numpy()
    out['multihead_eval'] = mh(xq, xq, xq).detach().numpy()

    # ---- whole models: 1 s clips for all seven, 10 s clips for the two BASELINE models
    for name in MODEL_NAMES:
        cls = getattr(ref_models, name)
        torch.manual_seed(0)
        model = cls(**CTOR)
        sd = model.state_dict()
        meta['models'][name] = {
            'state_dict': {k: list(v.shape) for k, v in sd.items()},
            'trainable': int(sum(p.numel() for p in model.parameters() if p.requires_grad)),
            'frozen': int(sum(p.numel() for p in model.parameters() if not p.requires_grad)),
        }
        # a few weight fingerprints (pins init + RNG consumption order)
        fp = {k: float(v.double().sum()) for k, v in sd.items() if v.dtype.is_floating_point}
        meta['models'][name]['weight_sums'] = fp

        lengths = [32000] + ([320000] if name in ('Cnn_9layers_FrameAvg',
                                                  'Cnn_9layers_Gru_FrameAtt') else [])
        for L in lengths:
            tag = '%s/L%d' % (name, L)
            pcm, target = synth(4, L, seed=1234)
            wave = torch.from_numpy((pcm / 32767.).astype(np.float32))
            tgt = torch.from_numpy(target)
            # eval forward (pytorch_utils.forward semantics: eval, no lambda)
            model.eval()
            with torch.no_grad():
                o = model(wave[:2])
            out[tag + '/eval_clip'] = o['clipwise_output'].numpy()
            out[tag + '/eval_frame'] = o['framewise_output'].numpy()
            out[tag + '/eval_emb_sum'] = np.array(float(o['embedding'].double().sum()))
            if 'Transformer' in name:
                continue                      # train-mode dropout is device-Philox: not comparable
            # one training step body (main.py:233-258) with mixup, SpecAug seeded
            import copy
            m2 = copy.deepcopy(model)
            m2.train()
            lam = torch.Tensor(ref_util.Mixup(1., 1234).get_lambda(4))
            torch.manual_seed(1)
            o = m2(wave, lam)
            t2 = {'target': ref_pu.do_mixup(tgt, lam)}
            loss = ref_losses.get_loss_func('clip_bce')(o, t2)
            loss.backward()
            out[tag + '/train_clip'] = o['clipwise_output'].detach().numpy()
            out[tag + '/train_loss'] = np.array(loss.item())
            gn = {k: float(p.grad.double().norm()) for k, p in m2.named_parameters()
                  if p.grad is not None}
            meta.setdefault('grad_norms', {})[tag] = gn
            out[tag + '/bn0_running_mean'] = m2.bn0.running_mean.numpy().copy()
            out[tag + '/bn0_running_var'] = m2.bn0.running_var.numpy().copy()

    np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **out)
    with open(os.path.join(HERE, 'reference_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
