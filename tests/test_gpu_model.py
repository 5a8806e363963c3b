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


@pytest.mark.parametrize('name', NAMES + ['Cnn_9layers_Transformer_FrameAvg', 'Cnn_9layers_Transformer_FrameAtt'])
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


@pytest.mark.parametrize('name', ['Cnn_9layers_FrameAvg', 'Cnn_9layers_Gru_FrameAtt', 'Cnn_9layers_Transformer_FrameAvg',
                                  'Cnn_9layers_Transformer_FrameAtt'])
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


ALL_NAMES = NAMES + ['Cnn_9layers_Transformer_FrameAvg', 'Cnn_9layers_Transformer_FrameAtt']


def _no_dropout(*models_):
    """MultiHead's two dropouts draw from device Philox streams that no custom kernel can replay
    bit-for-bit: model-level parity runs with p = 0 on both sides; the dropout
    kernels are checked statistically in test_gpu_attention.py."""
    for m in models_:
        if hasattr(m, 'multihead'):
            m.multihead.dropout.p = 0.0
            m.multihead.attention.dropout.p = 0.0


@pytest.mark.parametrize('name', ALL_NAMES)
def test_train_step_matches_oracle(name):
    """One training step (SpecAugment + mixup + batch-statistics BN + clip_bce + backward).

    Forward: clip-wise outputs / loss against the pure-fp32 oracle.
    Gradients: bf16 storage of conv operands flips ReLU masks of near-zero pre-activations, which
    perturbs random-init gradients by 10-25 % (measured ON THE ORACLE ITSELF by
    oracle/bf16_emulation.py, no CUDA involved).  The kernels are therefore required to (i) stay
    inside that band: err(cuda, fp32) <= 2 x err(bf16-emulated oracle, fp32) + 2e-2, and (ii) point
    the same way: cosine >= 0.9; parameters downstream of the trunk (head, GRU, attention) see no
    such noise and must match to 3e-2.  Per-kernel backward parity at tight tolerances lives in
    test_gpu_conv / _bnpool / _conv_c1 / _heads / _gru / _attention."""
    from oracle import sed, bf16_emulation
    ref, mine = _pair(name)
    emu = bf16_emulation.emulate_bf16_storage(copy.deepcopy(ref))
    _no_dropout(ref, emu, mine)
    _, wave, target = sed.synthetic_batch(4, 32000, seed=1234)
    wave, target = torch.from_numpy(wave), torch.from_numpy(target)
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(4))
    grads = {}
    for tag, m in (('ref', ref), ('emu', emu)):
        m.train()
        torch.manual_seed(1)
        o_m = m(wave, lam)
        l_m = sed.clip_bce(o_m, {'target': sed.mix_pairs(target, lam)})
        l_m.backward()
        grads[tag] = {k: p.grad for k, p in m.named_parameters()}
        if tag == 'ref':
            o_ref, loss_ref = o_m, l_m

    from sound_event_detection_dcase2017_task4_b200 import losses, pytorch_utils
    mine.train()
    torch.manual_seed(1)                                          # same SpecAugment stripes
    o = mine(wave.cuda(), lam.cuda())
    tgt = pytorch_utils.do_mixup(target.cuda(), lam.cuda())
    loss = losses.clip_bce(o, {'target': tgt})
    loss.backward()
    # batch statistics of 4 one-second clips + bf16 conv operands: 3e-3 relative on the outputs
    # (1e-2 for FrameMax, whose clip output is a single un-averaged frame)
    tol = 1e-2 if name.endswith('FrameMax') else 3e-3
    assert _rel(o['clipwise_output'].detach().cpu(), o_ref['clipwise_output'].detach()) <= tol
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    # BatchNorm running statistics are updated like nn.BatchNorm2d (momentum 0.1, unbiased var)
    assert torch.allclose(mine.bn0.running_mean.cpu(), ref.bn0.running_mean, atol=1e-4, rtol=1e-5)
    assert torch.allclose(mine.bn0.running_var.cpu(), ref.bn0.running_var, rtol=1e-4)
    assert int(mine.bn0.num_batches_tracked) == 1 and int(mine.conv_block4.bn2.num_batches_tracked) == 1
    assert torch.allclose(mine.conv_block2.bn1.running_var.cpu(), ref.conv_block2.bn1.running_var, rtol=5e-3)
    worst = {}
    for k, p in mine.named_parameters():
        g_ref, g_emu = grads['ref'][k], grads['emu'][k]
        if g_ref is None:
            assert p.grad is None, k                             # dead / frozen parameters in both
            continue
        assert p.grad is not None, k
        if k in ('att_block.att.bias', 'multihead.w_ks.bias'):
            continue          # softmax shift invariance: these gradients are ~0 (clamp / +1e-6 residue)
        a, b, e = p.grad.cpu().double().flatten(), g_ref.double().flatten(), g_emu.double().flatten()
        err = (a - b).norm().item() / max(b.norm().item(), 1e-30)
        band = (e - b).norm().item() / max(b.norm().item(), 1e-30)
        cos = torch.dot(a, b).item() / max(a.norm().item() * b.norm().item(), 1e-30)
        worst[k] = (err, band, cos)
        assert err <= max(3e-2, 2.0 * band + 2e-2) and cos >= 0.9, (k, err, band, cos)
    print(name, 'worst grad rel-L2 %.3e' % max(v[0] for v in worst.values()))


def test_fused_trainer_tracks_oracle_training():
    """8 optimizer steps of the reference loop body (main.py:233-258: lambda, forward, mixup of the
    targets, clip_bce, backward, Adam-amsgrad) through FusedTrainer (no autograd, flat buffers,
    fused Adam kernel) against the oracle driven by torch.optim.Adam: loss trajectories agree."""
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer
    name = 'Cnn_9layers_Gru_FrameAtt'
    ref, mine = _pair(name)
    _, wave, target = sed.synthetic_batch(8, 32000, seed=4321)
    wave, target = torch.from_numpy(wave), torch.from_numpy(target)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0., amsgrad=True)
    mix = sed.MixupLambda(1., 1234)
    torch.manual_seed(3)
    ref_losses = [sed.train_step(ref, opt, wave, target, torch.Tensor(mix.get_lambda(8))).item() for _ in range(8)]
    trainer = FusedTrainer(mine.train(), lr=1e-3)
    mix = sed.MixupLambda(1., 1234)
    torch.manual_seed(3)
    w_d, t_d = wave.cuda(), target.cuda()
    losses_ = [trainer.step(w_d, t_d, torch.Tensor(mix.get_lambda(8)).cuda()).item() for _ in range(8)]
    print('oracle', ref_losses)
    print('cuda  ', losses_)
    assert ref_losses[-1] < ref_losses[0]                         # it is actually learning
    for a, b in zip(losses_, ref_losses):
        assert abs(a - b) <= 5e-2 * abs(b), (losses_, ref_losses)   # bf16-storage noise compounds over the 8 steps
    # the big tensors moved the same way (Adam's early updates are sign-like, so compare displacements)
    torch.manual_seed(0)
    init = sed.build(name).state_dict()
    for (k, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        if p.requires_grad and p.numel() >= 4096 and 'bn_att' not in k and 'layer_norm' not in k:
            dm = (p.detach().cpu() - init[k]).double().flatten()
            dr = (q.detach() - init[k]).double().flatten()
            cos = torch.dot(dm, dr).item() / max(dm.norm().item() * dr.norm().item(), 1e-30)
            assert cos >= 0.7, (k, cos)


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


@pytest.mark.parametrize('name', ['Cnn_9layers_Gru_FrameAtt', 'Cnn_9layers_FrameAvg'])
def test_eval_mode_is_differentiable_like_the_reference(name):
    """the reference modules are differentiable in eval() mode (frozen-BatchNorm fine-tuning, saliency):
    running statistics are constants of the graph, no SpecAugment, no mixup.  Same gradient rule as the train step."""
    from oracle import sed, bf16_emulation
    from sound_event_detection_dcase2017_task4_b200 import losses
    ref, mine = _pair(name)
    g = torch.Generator().manual_seed(11)
    sd = ref.state_dict()
    for k in sd:                                         # running statistics away from (0, 1)
        if k.endswith('running_mean') and 'bn_att' not in k:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
        if k.endswith('running_var') and 'bn_att' not in k:
            sd[k] = sd[k] * (0.5 + torch.rand(sd[k].shape, generator=g))
    ref.load_state_dict(sd)
    mine.load_state_dict(sd)
    emu = bf16_emulation.emulate_bf16_storage(copy.deepcopy(ref))
    _, wave, target = sed.synthetic_batch(4, 32000, seed=77)
    wave, target = torch.from_numpy(wave), torch.from_numpy(target)
    grads = {}
    for tag, m in (('ref', ref), ('emu', emu)):
        m.eval()
        o_m = m(wave)
        sed.clip_bce(o_m, {'target': target}).backward()
        grads[tag] = {k: p.grad for k, p in m.named_parameters()}
        if tag == 'ref':
            o_ref = o_m['clipwise_output'].detach()
    mine.eval()
    o = mine(wave.cuda())
    assert o['clipwise_output'].requires_grad
    losses.clip_bce(o, {'target': target.cuda()}).backward()
    assert _rel(o['clipwise_output'].detach().cpu(), o_ref) <= 2e-3
    assert int(mine.bn0.num_batches_tracked) == 0 and torch.equal(mine.bn0.running_mean.cpu(), sd['bn0.running_mean'])
    for k, p in mine.named_parameters():
        g_ref, g_emu = grads['ref'][k], grads['emu'][k]
        if g_ref is None:
            assert p.grad is None, k
            continue
        if k in ('att_block.att.bias',):
            continue
        a, b, e = p.grad.cpu().double().flatten(), g_ref.double().flatten(), g_emu.double().flatten()
        err = (a - b).norm().item() / max(b.norm().item(), 1e-30)
        band = (e - b).norm().item() / max(b.norm().item(), 1e-30)
        cos = torch.dot(a, b).item() / max(a.norm().item() * b.norm().item(), 1e-30)
        assert err <= max(3e-2, 2.0 * band + 2e-2) and cos >= 0.9, (k, err, band, cos)


@pytest.mark.parametrize('name', ['Cnn_9layers_Gru_FrameAtt', 'Cnn_9layers_Transformer_FrameAvg'])
def test_cuda_graph_replay_equals_eager_steps(name):
    """FusedTrainer(use_graph=True) captures the step once per input-buffer set and replays it; SpecAugment stripes,
    Adam's bias corrections and (Transformer variants) the dropout generator state reach the replay through device
    memory.  Loss trajectory, parameters, optimizer state and the CUDA generator's final offset must be BIT-identical
    to eager launches (same kernels, same order, same Philox counters)."""
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import models
    from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer
    _, wave, target = sed.synthetic_batch(8, 32000, seed=21)
    wave, target = torch.from_numpy(wave).cuda(), torch.from_numpy(target).cuda()
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(8)).cuda()
    runs = []
    for use_graph in (False, True):
        torch.manual_seed(0)
        model = getattr(models, name)(*CTOR).cuda().train()
        tr = FusedTrainer(model, lr=1e-3, use_graph=use_graph)
        torch.manual_seed(9)
        losses = [float(tr.step(wave, target, lam)) for _ in range(7)]
        torch.cuda.synchronize()
        runs.append((losses, tr.flat_param.clone(), tr.max_exp_avg_sq.clone(), int(model.bn0.num_batches_tracked), tr,
                     torch.cuda.default_generators[0].get_offset()))
    assert runs[1][4].use_graph and runs[1][4].graph_launches > 50 and len(runs[1][4]._graphs) == 1
    assert runs[0][0] == runs[1][0], (runs[0][0], runs[1][0])
    assert torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])
    assert runs[0][3] == runs[1][3] == 7
    assert runs[0][5] == runs[1][5]
    if 'Transformer' in name:
        assert len(set(runs[1][0])) == 7                       # a fresh dropout mask per replay, not the captured one
