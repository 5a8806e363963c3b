"""oracle/sed.py against golden vectors minted from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import copy

import numpy as np
import pytest
import torch

from oracle import sed

NAMES = list(sed.SPECS)


def _synth(n, L):
    pcm, wave, target = sed.synthetic_batch(n, L, seed=1234)
    return torch.from_numpy(wave), torch.from_numpy(target)


def test_mixup_lambda_stream(golden):
    g, _ = golden
    m = sed.MixupLambda(1., 1234)
    a = m.get_lambda(32)
    assert a[0] == 0.23538938957272115                      # SURVEY.md section 4 known answer
    assert np.array_equal(a, g['mixup_lambda_first32'])
    assert np.array_equal(m.get_lambda(6), g['mixup_lambda_next6'])


def test_int16_to_float32_every_value(golden):
    g, _ = golden
    pcm = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    assert np.array_equal(sed.int16_to_float32(pcm), g['int16_to_float32_all'])


def test_repeat_frames_is_exact_index_map(golden):
    g, _ = golden
    y = sed.repeat_frames(torch.from_numpy(g['interp_in']), 8).numpy()
    assert np.array_equal(y, g['interp_out'])
    assert y.shape[1] == 5 * 8
    for t in range(y.shape[1]):
        assert np.array_equal(y[:, t], g['interp_in'][:, t // 8])


def test_mix_pairs(golden):
    g, _ = golden
    lam = torch.from_numpy(g['mixup_lam'])
    assert np.array_equal(sed.mix_pairs(torch.from_numpy(g['mixup_in']), lam).numpy(), g['mixup_out'])
    assert np.array_equal(sed.mix_pairs(torch.from_numpy(g['mixup_target_in']), lam).numpy(),
                          g['mixup_target_out'])


def test_convblock(golden):
    g, _ = golden
    torch.manual_seed(3)
    cb = sed.ConvBlock(4, 8)
    x = torch.from_numpy(g['convblock_in'])
    cb.train()
    np.testing.assert_allclose(cb(x).detach().numpy(), g['convblock_train_avg22'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(cb.bn1.running_mean.numpy(), g['convblock_running_mean1'], atol=1e-7)
    np.testing.assert_allclose(cb.bn2.running_var.numpy(), g['convblock_running_var2'], atol=1e-7)
    cb.eval()
    np.testing.assert_allclose(cb(x, (1, 1), 'avg').detach().numpy(), g['convblock_eval_avg11'], atol=1e-6)
    np.testing.assert_allclose(cb(x, (2, 2), 'max').detach().numpy(), g['convblock_eval_max22'], atol=1e-6)
    np.testing.assert_allclose(cb(x, (2, 2), 'avg+max').detach().numpy(), g['convblock_eval_avgmax22'], atol=1e-6)
    with pytest.raises(Exception):
        cb(x, (2, 2), 'nope')


def test_attblock(golden):
    g, _ = golden
    torch.manual_seed(4)
    ab = sed.AttBlock(16, 5, activation='sigmoid')
    clip, natt, cla = ab(torch.from_numpy(g['attblock_in']))
    np.testing.assert_allclose(clip.detach().numpy(), g['attblock_clip'], atol=1e-6)
    np.testing.assert_allclose(natt.detach().numpy(), g['attblock_norm_att'], atol=1e-6)
    np.testing.assert_allclose(cla.detach().numpy(), g['attblock_cla'], atol=1e-6)


def test_multihead_eval(golden):
    g, _ = golden
    torch.manual_seed(5)
    mh = sed.MultiHead(4, 32, 8, 8, 0.2).eval()
    x = torch.from_numpy(g['multihead_in'])
    np.testing.assert_allclose(mh(x, x, x).detach().numpy(), g['multihead_eval'], atol=1e-6)


@pytest.mark.parametrize('name', NAMES)
def test_state_dict_contract_and_init(golden, name):
    """Same keys, shapes, parameter counts and -- under torch.manual_seed(0) -- the same
    initial weights as the reference constructor."""
    _, meta = golden
    ref = meta['models'][name]
    torch.manual_seed(0)
    m = sed.build(name)
    sd = m.state_dict()
    assert list(sd.keys()) == sorted(sd.keys(), key=list(sd.keys()).index)
    assert {k: list(v.shape) for k, v in sd.items()} == ref['state_dict']
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == ref['trainable']
    assert sum(p.numel() for p in m.parameters() if not p.requires_grad) == ref['frozen'] == 1083456
    for k, s in ref['weight_sums'].items():
        assert float(sd[k].double().sum()) == pytest.approx(s, rel=1e-9, abs=1e-9), k


@pytest.mark.parametrize('name', NAMES)
def test_model_eval_forward_1s(golden, name):
    g, _ = golden
    torch.manual_seed(0)
    m = sed.build(name).eval()
    wave, _ = _synth(4, 32000)
    with torch.no_grad():
        o = m(wave[:2])
    tag = '%s/L32000' % name
    assert o['framewise_output'].shape == g[tag + '/eval_frame'].shape == (2, 96, 17)
    np.testing.assert_allclose(o['clipwise_output'].numpy(), g[tag + '/eval_clip'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o['framewise_output'].numpy(), g[tag + '/eval_frame'], rtol=0, atol=2e-6)
    assert float(o['embedding'].double().sum()) == pytest.approx(float(g[tag + '/eval_emb_sum']), rel=1e-4, abs=1e-3)


@pytest.mark.parametrize('name', [n for n in NAMES if 'Transformer' not in n])
def test_model_train_step_1s(golden, name):
    g, meta = golden
    torch.manual_seed(0)
    m = sed.build(name).train()
    wave, target = _synth(4, 32000)
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(4))
    torch.manual_seed(1)
    o = m(wave, lam)
    loss = sed.clip_bce(o, {'target': sed.mix_pairs(target, lam)})
    loss.backward()
    tag = '%s/L32000' % name
    np.testing.assert_allclose(o['clipwise_output'].detach().numpy(), g[tag + '/train_clip'], atol=2e-6)
    assert loss.item() == pytest.approx(float(g[tag + '/train_loss']), abs=2e-6)
    np.testing.assert_allclose(m.bn0.running_mean.numpy(), g[tag + '/bn0_running_mean'], atol=1e-5)
    np.testing.assert_allclose(m.bn0.running_var.numpy(), g[tag + '/bn0_running_var'], rtol=1e-5)
    gn = meta['grad_norms'][tag]
    mine = {k: float(p.grad.double().norm()) for k, p in m.named_parameters() if p.grad is not None}
    assert set(mine) == set(gn)            # dead params (bn_att, layer_norm) get no grad in both
    for k in gn:
        assert mine[k] == pytest.approx(gn[k], rel=2e-3, abs=1e-7), k


@pytest.mark.parametrize('name', ['Cnn_9layers_FrameAvg', 'Cnn_9layers_Gru_FrameAtt'])
def test_model_full_length_10s(golden, name):
    g, _ = golden
    torch.manual_seed(0)
    m = sed.build(name)
    wave, target = _synth(4, 320000)
    tag = '%s/L320000' % name
    m.eval()
    with torch.no_grad():
        o = m(wave[:2])
    assert o['framewise_output'].shape == (2, 1000, 17)
    np.testing.assert_allclose(o['clipwise_output'].numpy(), g[tag + '/eval_clip'], atol=2e-6)
    np.testing.assert_allclose(o['framewise_output'].numpy(), g[tag + '/eval_frame'], atol=2e-6)
    m2 = copy.deepcopy(m).train()
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(4))
    torch.manual_seed(1)
    o = m2(wave, lam)
    loss = sed.clip_bce(o, {'target': sed.mix_pairs(target, lam)})
    np.testing.assert_allclose(o['clipwise_output'].detach().numpy(), g[tag + '/train_clip'], atol=2e-6)
    assert loss.item() == pytest.approx(float(g[tag + '/train_loss']), abs=2e-6)
