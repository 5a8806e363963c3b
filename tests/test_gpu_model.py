"""-m gpu: whole-model parity of the drop-in classes (libsedb200 kernels through the C ABI) against
the CPU oracle (oracle/sed.py, itself pinned to the unmodified reference by tests/golden).

Tolerances (BASELINE.json north_star): clip-wise outputs <= 1e-3 relative in eval mode; the
x8 frame interpolation is an exact index map (bit-exact); SpecAugment stripes bit-exact (same
torch CPU RNG stream).  Training-mode outputs / gradients are compared at the tolerances stated
inline (bf16 tensor-core operands, fp32 accumulation, batch statistics of only 4 clips)."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CTOR = (32000, 1024, 320, 64, 50, 14000, 17)
NAMES = ['Cnn_9layers_FrameMax', 'Cnn_9layers_FrameAvg', 'Cnn_9layers_FrameAtt', 'Cnn_9layers_Gru_FrameAvg',
         'Cnn_9layers_Gru_FrameAtt']


def _pair(name):
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import models
    torch.manual_seed(0)
    ref = sed.build(name)
    torch.manual_seed(0)
    mine = getattr(models, name)(*CTOR)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka            # same ctor => same init, same keys
    return ref, mine.cuda()


def _rel(a, b, floor=1e-6):
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item()


@pytest.mark.parametrize('name', NAMES)
def test_eval_forward_1s(name):
    from oracle import sed
    ref, mine = _pair(name)
    _, wave, _ = sed.synthetic_batch(2, 32000, seed=1234)
    wave = torch.from_numpy(wave)
    ref.eval(); mine.eval()
    with torch.no_grad():
        o_ref = ref(wave)
        o = mine(wave.cuda())
    clip, frame = o['clipwise_output'].cpu(), o['framewise_output'].cpu()
    assert frame.shape == o_ref['framewise_output'].shape == (2, 96, 17)
    assert _rel(clip, o_ref['clipwise_output']) <= 1e-3
    assert _rel(frame, o_ref['framewise_output']) <= 2e-3
    for t in range(frame.shape[1]):                              # interpolate: exact copies, t -> t // 8
        assert torch.equal(frame[:, t], frame[:, 8 * (t // 8)])
    assert o['embedding'].shape == o_ref['embedding'].shape


@pytest.mark.parametrize('name', ['Cnn_9layers_FrameAvg', 'Cnn_9layers_Gru_FrameAtt'])
def test_eval_forward_10s_with_trained_like_stats(name):
    """Full-length clips, running statistics moved away from their init values and a sharpened head
    so that probabilities span (0, 1)."""
    from oracle import sed
    ref, mine = _pair(name)
    g = torch.Generator().manual_seed(5)
    sd = ref.state_dict()
    for k in sd:
        if k.endswith('running_mean') and 'bn_att' not in k:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
        if k.endswith('running_var') and 'bn_att' not in k:
            sd[k] = sd[k] * (0.5 + torch.rand(sd[k].shape, generator=g))
        if k in ('fc.weight', 'att_block.cla.weight', 'att_block.att.weight'):
            sd[k] = sd[k] * 8.0
    ref.load_state_dict(sd)
    mine.load_state_dict(sd)                                    # strict: identical key set
    _, wave, _ = sed.synthetic_batch(2, 320000, seed=99)
    wave = torch.from_numpy(wave)
    ref.eval(); mine.eval()
    with torch.no_grad():
        o_ref = ref(wave)
        o = mine(wave.cuda())
    assert o['framewise_output'].shape == (2, 1000, 17)
    c_ref = o_ref['clipwise_output']
    assert c_ref.max() - c_ref.min() > 0.05
    # The 8x sharper head multiplies the logit error of the bf16-operand trunk by 8 (measured 1.9e-3 /
    # 2.7e-3 here vs <= 2.5e-4 at the reference's own initialisation): stress tolerance 4e-3.  The
    # north-star gate (1e-3 at the reference's init) is test_eval_forward_1s / _10s_reference_init.
    assert _rel(o['clipwise_output'].cpu(), c_ref) <= 4e-3


@pytest.mark.parametrize('name', ['Cnn_9layers_FrameAvg', 'Cnn_9layers_Gru_FrameAtt'])
def test_eval_forward_10s_reference_init(name):
    """Full-length 10 s clips at the reference's own initialisation: clip-wise outputs <= 1e-3 rel."""
    from oracle import sed
    ref, mine = _pair(name)
    _, wave, _ = sed.synthetic_batch(2, 320000, seed=99)
    wave = torch.from_numpy(wave)
    ref.eval(); mine.eval()
    with torch.no_grad():
        o_ref = ref(wave)
        o = mine(wave.cuda())
    assert o['framewise_output'].shape == (2, 1000, 17)
    assert _rel(o['clipwise_output'].cpu(), o_ref['clipwise_output']) <= 1e-3
    assert _rel(o['framewise_output'].cpu(), o_ref['framewise_output']) <= 2e-3


@pytest.mark.parametrize('name', NAMES)
def test_train_step_matches_oracle(name):
    from oracle import sed
    ref, mine = _pair(name)
    _, wave, target = sed.synthetic_batch(4, 32000, seed=1234)
    wave, target = torch.from_numpy(wave), torch.from_numpy(target)
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(4))
    ref.train(); mine.train()
    torch.manual_seed(1)
    o_ref = ref(wave, lam)
    loss_ref = sed.clip_bce(o_ref, {'target': sed.mix_pairs(target, lam)})
    loss_ref.backward()

    from sound_event_detection_dcase2017_task4_b200 import losses, pytorch_utils
    torch.manual_seed(1)                                          # same SpecAugment stripes
    o = mine(wave.cuda(), lam.cuda())
    tgt = pytorch_utils.do_mixup(target.cuda(), lam.cuda())
    loss = losses.clip_bce(o, {'target': tgt})
    loss.backward()
    # batch statistics of 4 one-second clips + bf16 conv operands: 3e-3 relative on the outputs
    assert _rel(o['clipwise_output'].detach().cpu(), o_ref['clipwise_output'].detach()) <= 3e-3
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    # BatchNorm running statistics are updated like nn.BatchNorm2d (momentum 0.1, unbiased var)
    assert torch.allclose(mine.bn0.running_mean.cpu(), ref.bn0.running_mean, atol=1e-4, rtol=1e-5)
    assert torch.allclose(mine.bn0.running_var.cpu(), ref.bn0.running_var, rtol=1e-4)
    assert int(mine.bn0.num_batches_tracked) == 1 and int(mine.conv_block4.bn2.num_batches_tracked) == 1
    assert torch.allclose(mine.conv_block2.bn1.running_var.cpu(), ref.conv_block2.bn1.running_var, rtol=5e-3)
    # gradients: relative L2 error per parameter tensor (bf16 MMA operands in dgrad / wgrad)
    ref_g = {k: p.grad for k, p in ref.named_parameters()}
    worst = 0.0
    for k, p in mine.named_parameters():
        if ref_g[k] is None:
            assert p.grad is None, k                             # dead / frozen parameters in both
            continue
        assert p.grad is not None, k
        a, b = p.grad.cpu().double(), ref_g[k].double()
        err = (a - b).norm().item() / max(b.norm().item(), 1e-12)
        worst = max(worst, err)
        assert err <= 3e-2, (k, err)
    print(name, 'worst grad rel-L2', worst)


def test_dataparallel_single_gpu_and_no_cpu_path():
    """main.py:138 wraps the model in nn.DataParallel; the module must survive replicate()."""
    from oracle import sed
    ref, mine = _pair('Cnn_9layers_FrameAvg')
    _, wave, _ = sed.synthetic_batch(2, 32000, seed=1)
    dp = torch.nn.DataParallel(mine)
    dp.eval()
    with torch.no_grad():
        o = dp(torch.from_numpy(wave).cuda())
        o_ref = ref.eval()(torch.from_numpy(wave))
    assert _rel(o['clipwise_output'].cpu(), o_ref['clipwise_output']) <= 1e-3
    with pytest.raises(RuntimeError, match='CUDA'):
        copy.deepcopy(mine).cpu()(torch.from_numpy(wave))
