"""-m gpu: the reference's own loops executed call-for-call on the drop-in modules.

* the training loop body of /root/reference/pytorch/main.py:138-258 (DataParallel wrapper, torch.optim.Adam with
  amsgrad, Mixup, move_data_to_device, model.train(), do_mixup on the targets, loss.backward(), optimizer.step()),
  fed by the drop-in data_generator, against the CPU oracle driven the same way;
* nn.DataParallel with TWO replicas in train mode (the case ADVICE r1 showed broken);
* two FusedTrainer ranks (gloo group, CUDA tensors) against a single-process computation and the oracle;
* train-mode parity at a batch where the batch statistics are stable (32 raw 10 s clips);
* the inference loop of /root/reference/pytorch/pytorch_utils.py:25-77 (``forward``).
"""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTOR = (32000, 1024, 320, 64, 50, 14000, 17)


def _pair(name):
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import models
    torch.manual_seed(0)
    ref = sed.build(name)
    torch.manual_seed(0)
    mine = getattr(models, name)(*CTOR)
    return ref, mine


def _rel(a, b, floor=1e-6):
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item()


def _store(tmp_path, n, samples, frames=None, seed=0):
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    rs = np.random.RandomState(seed)
    wave = rs.randint(-8192, 8192, size=(n, samples)).astype(np.int16)
    target = rs.rand(n, 17) < 0.2
    strong = (rs.rand(n, frames, 17) < 0.1) if frames else None
    path = str(tmp_path / 'store')
    dg.ClipStore.write(path, ['Y%04d.wav' % i for i in range(n)], wave, target, strong)
    return path


def test_main_py_training_loop_body_call_for_call(tmp_path):
    """main.py:138-258 with `--augmentation mixup --batch_size 4`, 6 iterations, on the drop-in modules; the oracle
    (pinned to the unmodified reference) runs the same statements on the CPU.  SpecAugment draws from the torch CPU
    generator on both sides, so it is re-seeded before every forward."""
    import torch.optim as optim
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    from sound_event_detection_dcase2017_task4_b200.losses import get_loss_func
    from sound_event_detection_dcase2017_task4_b200.pytorch_utils import move_data_to_device, do_mixup
    from sound_event_detection_dcase2017_task4_b200.utilities import Mixup
    path = _store(tmp_path, 24, 32000)
    batch_size, iters, device = 4, 6, 'cuda'
    ref, model = _pair('Cnn_9layers_Gru_FrameAtt')
    loss_func = get_loss_func('clip_bce')

    # ---- main.py:138-173
    # (the oracle below normalises over the whole batch, i.e. ONE replica: pin the wrapper to one device on multi-GPU boxes)
    model = torch.nn.DataParallel(model) if torch.cuda.device_count() == 1 else torch.nn.DataParallel(model, device_ids=[0])
    model.to(device)
    optimizer = optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-08, weight_decay=0., amsgrad=True)
    dataset = dg.DCASE2017Task4Dataset()
    train_sampler = dg.TrainSampler(hdf5_path=path, batch_size=batch_size * 2)
    train_loader = torch.utils.data.DataLoader(dataset=dataset, batch_sampler=train_sampler, collate_fn=dg.collate_fn,
                                               num_workers=0, pin_memory=True)
    mixup_augmenter = Mixup(mixup_alpha=1.)
    losses, batches = [], []
    for iteration, batch_data_dict in enumerate(train_loader):                       # main.py:187
        if iteration == iters:
            break
        batches.append({k: np.array(v) for k, v in batch_data_dict.items()})
        batch_data_dict['mixup_lambda'] = mixup_augmenter.get_lambda(batch_size=len(batch_data_dict['waveform']))
        for key in batch_data_dict.keys():                                            # main.py:237-239
            batch_data_dict[key] = move_data_to_device(batch_data_dict[key], device)
        model.train()
        torch.manual_seed(100 + iteration)
        batch_output_dict = model(batch_data_dict['waveform'], batch_data_dict['mixup_lambda'])
        batch_target_dict = {'target': do_mixup(batch_data_dict['target'], batch_data_dict['mixup_lambda'])}
        loss = loss_func(batch_output_dict, batch_target_dict)
        losses.append(float(loss.detach()))                                           # main.py:253 prints it
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    assert model.module.conv_block1.conv1.weight.grad is not None
    assert int(model.module.bn0.num_batches_tracked) == iters
    opt_state = optimizer.state_dict()['state']         # frozen / dead parameters have no entry (their grad is None)
    assert sorted(opt_state[min(opt_state)].keys()) == ['exp_avg', 'exp_avg_sq', 'max_exp_avg_sq', 'step']

    # ---- the oracle, same statements
    opt_ref = optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-08, weight_decay=0., amsgrad=True)
    mix_ref = sed.MixupLambda(1., 1234)
    ref_losses = []
    for iteration, b in enumerate(batches):
        lam = torch.Tensor(mix_ref.get_lambda(len(b['waveform'])))
        ref.train()
        torch.manual_seed(100 + iteration)
        out = ref(torch.Tensor(b['waveform']), lam)
        l_ref = sed.clip_bce(out, {'target': sed.mix_pairs(torch.Tensor(b['target']), lam)})
        ref_losses.append(float(l_ref.detach()))
        opt_ref.zero_grad()
        l_ref.backward()
        opt_ref.step()
    print('drop-in', losses)
    print('oracle ', ref_losses)
    assert abs(losses[0] - ref_losses[0]) <= 2e-3 * abs(ref_losses[0])                # same weights: forward parity
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 3e-2 * abs(b), (losses, ref_losses)
    assert ref_losses[-1] < ref_losses[0] and losses[-1] < losses[0]
    # checkpoint contract (main.py:222-226): the wrapped module's state_dict has the reference's keys
    assert list(model.module.state_dict().keys()) == list(ref.state_dict().keys())


@pytest.mark.parametrize('name', ['Cnn_9layers_Gru_FrameAtt', 'Cnn_9layers_Transformer_FrameAvg'])
def test_dataparallel_two_replicas_train_step(name):
    """nn.DataParallel with two replicas (two GPUs when the box has them, else both on cuda:0) in train mode: forward
    on replicas whose parameters() is empty, clip_bce on the gathered output, loss.backward() through DataParallel's
    Broadcast / Gather nodes.  Gradients must equal the single-device computation of the same two shards."""
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import losses, pytorch_utils
    _, model = _pair(name)
    model = model.cuda()
    for m in (model.spec_augmenter.time_dropper, model.spec_augmenter.freq_dropper):
        m.drop_width = 1                      # zero-width stripes: the replicas' draw order is thread-dependent
    if hasattr(model, 'multihead'):
        model.multihead.dropout.p = 0.0
        model.multihead.attention.dropout.p = 0.0
    _, wave, target = sed.synthetic_batch(8, 32000, seed=5)
    wave, target = torch.from_numpy(wave).cuda(), torch.from_numpy(target).cuda()
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(8)).cuda()
    ids = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    dp = torch.nn.DataParallel(model, device_ids=ids)
    dp.train()
    out = dp(wave, lam)
    assert out['clipwise_output'].shape == (4, 17) and out['clipwise_output'].requires_grad
    loss = losses.clip_bce(out, {'target': pytorch_utils.do_mixup(target, lam)})
    model.zero_grad()
    loss.backward()
    got = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    assert 'conv_block1.conv1.weight' in got and 'bn0.weight' in got
    # single device, the same two shards one after the other (per-replica BatchNorm statistics, mean over 4 x 17)
    model.zero_grad()
    total = 0.0
    for lo in (0, 4):
        o = model(wave[lo:lo + 4], lam[lo:lo + 4])
        l_half = losses.clip_bce(o, {'target': pytorch_utils.do_mixup(target[lo:lo + 4], lam[lo:lo + 4])})
        (0.5 * l_half).backward()
        total += 0.5 * float(l_half)
    assert abs(float(loss) - total) <= 1e-5 * abs(total)
    for k, p in model.named_parameters():
        if p.grad is None:                    # dead parameters (bn_att, layer_norm): DataParallel's Broadcast node hands
            assert k not in got or float(got[k].abs().max()) == 0.0      # back zeros where one device leaves None
            continue
        a, b = got[k].double().flatten(), p.grad.double().flatten()
        assert (a - b).norm().item() <= 1e-4 * max(b.norm().item(), 1e-12), k


def test_two_fused_trainer_ranks_equal_single_process_and_track_the_oracle(tmp_path, run_two_ranks):
    """World size 2 (two processes sharing cuda:0, gloo group, CUDA tensors): each rank steps on its shard with its
    loss scaled by 1/2, ONE all-reduce of the flat gradient, fused Adam.  The reduced gradient must equal (i) the
    same two shard-steps computed in ONE process and averaged -- bit for bit, fp32 addition of two terms commutes --
    and (ii) within the a16 tolerances the CPU oracle on the concatenated batch with per-shard BatchNorm statistics
    (DataParallel semantics, main.py:138)."""
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer
    r0, r1 = run_two_ranks('gpu', tmp_path)
    assert torch.equal(r0['grad'], r1['grad']) and torch.equal(r0['param'], r1['param'])
    assert r0['bounds'] == (0, 8) and r1['bounds'] == (8, 16)
    name = 'Cnn_9layers_Gru_FrameAtt'
    ref, mine = _pair(name)
    mine = mine.cuda().train()
    for m in (mine.spec_augmenter.time_dropper, mine.spec_augmenter.freq_dropper):
        m.drop_width = 1
    _, wave, target = sed.synthetic_batch(16, 32000, seed=77)
    lam = sed.MixupLambda(1., 1234).get_lambda(16).astype(np.float32)
    # (i) one process, no collective: snapshot the parameters, step on each shard from the same start, average
    trainer = FusedTrainer(mine, lr=1e-3, world_size=2)                 # world_size=2 scales the loss by 1/2 ...
    trainer.group = None
    import sound_event_detection_dcase2017_task4_b200.trainer as tr
    start = trainer.flat_param.clone()
    grads, losses_ = [], []
    real_exchange = tr.exchange_gradients
    tr.exchange_gradients = lambda g, w, grp=None: g                     # ... and this run has no peer to add
    try:
        for lo in (0, 8):
            trainer.flat_param.copy_(start)
            trainer.exp_avg.zero_(); trainer.exp_avg_sq.zero_(); trainer.max_exp_avg_sq.zero_(); trainer.step_count = 0
            l_shard = trainer.step(torch.from_numpy(wave[lo:lo + 8]).cuda(), torch.from_numpy(target[lo:lo + 8]).cuda(),
                                   torch.from_numpy(lam[lo:lo + 8]).cuda())
            grads.append(trainer.flat_grad.clone().cpu())
            losses_.append(float(l_shard))
    finally:
        tr.exchange_gradients = real_exchange
    assert torch.equal(grads[0] + grads[1], r0['grad'])
    assert losses_ == [r0['loss'], r1['loss']]
    # (ii) the oracle: mean of the two per-shard mean losses
    ref.train()
    for m in (ref.spec_augmenter.time_dropper, ref.spec_augmenter.freq_dropper):
        m.drop_width = 1
    total = 0
    for lo in (0, 8):
        lam_t = torch.from_numpy(lam[lo:lo + 8])
        out = ref(torch.from_numpy(wave[lo:lo + 8]), lam_t)
        l_ref = sed.clip_bce(out, {'target': sed.mix_pairs(torch.from_numpy(target[lo:lo + 8]), lam_t)})
        (0.5 * l_ref).backward()
        total += 0.5 * float(l_ref)
    assert abs(0.5 * (r0['loss'] + r1['loss']) - total) <= 2e-3 * abs(total)
    off = 0
    for k, p in ref.named_parameters():
        if not p.requires_grad:
            continue
        n = p.numel()
        g = r0['grad'][off:off + n].double()
        off += n
        if p.grad is None or k in ('att_block.att.bias',):
            continue
        b = p.grad.double().flatten()
        cos = torch.dot(g, b).item() / max(g.norm().item() * b.norm().item(), 1e-30)
        err = (g - b).norm().item() / max(b.norm().item(), 1e-30)
        if k.startswith(('gru.', 'att_block.')):
            assert err <= 3e-2, (k, err)                                  # downstream of the bf16 trunk: tight
        else:
            assert cos >= 0.9, (k, cos, err)                              # conv trunk: bf16-storage band
    assert off == r0['grad'].numel()


def test_train_mode_parity_at_a_stable_batch_10s():
    """train-mode forward (SpecAugment + mixup + batch statistics) on 32 raw 10 s clips; clip-wise
    outputs against the fp32 oracle at the north-star tolerance (1e-3 relative), plus the gradient error next to
    the oracle's own bf16-storage emulation at this size (reported; bounded by the DESIGN section 5 rule)."""
    from oracle import sed, bf16_emulation
    from sound_event_detection_dcase2017_task4_b200 import losses, pytorch_utils
    name = 'Cnn_9layers_Gru_FrameAtt'
    ref, mine = _pair(name)
    mine = mine.cuda()
    emu = bf16_emulation.emulate_bf16_storage(copy.deepcopy(ref))
    _, wave, target = sed.synthetic_batch(32, 320000, seed=2024)
    wave, target = torch.from_numpy(wave), torch.from_numpy(target)
    lam = torch.Tensor(sed.MixupLambda(1., 1234).get_lambda(32))
    grads = {}
    for tag, m in (('ref', ref), ('emu', emu)):
        m.train()
        torch.manual_seed(1)
        o_m = m(wave, lam)
        l_m = sed.clip_bce(o_m, {'target': sed.mix_pairs(target, lam)})
        l_m.backward()
        grads[tag] = {k: p.grad for k, p in m.named_parameters()}
        if tag == 'ref':
            o_ref, loss_ref = o_m['clipwise_output'].detach(), float(l_m)
        else:
            o_emu = o_m['clipwise_output'].detach()
    mine.train()
    torch.manual_seed(1)
    o = mine(wave.cuda(), lam.cuda())
    loss = losses.clip_bce(o, {'target': pytorch_utils.do_mixup(target.cuda(), lam.cuda())})
    loss.backward()
    rel = _rel(o['clipwise_output'].detach().cpu(), o_ref)
    rel_emu = _rel(o_emu, o_ref)
    print('train-mode clipwise rel err: cuda %.3e, bf16-emulated oracle %.3e; loss rel err %.3e'
          % (rel, rel_emu, abs(float(loss) - loss_ref) / abs(loss_ref)))
    assert rel <= 1e-3
    assert abs(float(loss) - loss_ref) <= 1e-3 * abs(loss_ref)
    worst = (None, 0.0, 0.0)
    for k, p in mine.named_parameters():
        g_ref, g_emu = grads['ref'][k], grads['emu'][k]
        if g_ref is None or k in ('att_block.att.bias',):
            continue
        a, b, e = p.grad.cpu().double().flatten(), g_ref.double().flatten(), g_emu.double().flatten()
        err = (a - b).norm().item() / max(b.norm().item(), 1e-30)
        band = (e - b).norm().item() / max(b.norm().item(), 1e-30)
        cos = torch.dot(a, b).item() / max(a.norm().item() * b.norm().item(), 1e-30)
        if err > worst[1]:
            worst = (k, err, band)
        assert err <= max(3e-2, 2.0 * band + 2e-2) and cos >= 0.9, (k, err, band, cos)
    print('worst gradient rel-L2 vs fp32 oracle: %s %.3e (bf16-emulated oracle: %.3e)' % worst)


def test_forward_inference_loop_matches_the_oracle_loop(tmp_path):
    """pytorch_utils.forward (pytorch_utils.py:25-77) as evaluate.py:50-54 calls it, on a TestSampler loader with a
    ragged last batch: reference keys / shapes / order, outputs against the oracle run batch by batch, int16 PCM
    batches bit-identical to fp32 ones, optional padding of the 1000 model frames to strong_target's length."""
    from oracle import sed
    from sound_event_detection_dcase2017_task4_b200 import data_generator as dg
    from sound_event_detection_dcase2017_task4_b200 import pytorch_utils as pu
    path = _store(tmp_path, 7, 32000, frames=101, seed=3)
    ref, mine = _pair('Cnn_9layers_Gru_FrameAtt')
    mine = torch.nn.DataParallel(mine.cuda())                    # evaluate.py is handed the DataParallel wrapper
    dataset = dg.DCASE2017Task4Dataset()

    def loader():
        return torch.utils.data.DataLoader(dataset=dataset, batch_sampler=dg.TestSampler(hdf5_path=path, batch_size=3),
                                           collate_fn=dg.collate_fn, num_workers=0, pin_memory=True)

    out = pu.forward(model=mine, data_loader=loader(), return_input=False, return_target=True)
    assert set(out.keys()) == {'audio_name', 'clipwise_output', 'framewise_output', 'target', 'strong_target'}
    assert out['clipwise_output'].shape == (7, 17) and out['framewise_output'].shape == (7, 96, 17)
    assert out['strong_target'].shape == (7, 101, 17) and list(out['audio_name']) == ['Y%04d.wav' % i for i in range(7)]
    ref.eval()
    want_clip, want_frame = [], []
    with torch.no_grad():
        for b in loader():
            o = ref(torch.Tensor(b['waveform']))
            want_clip.append(o['clipwise_output'].numpy())
            want_frame.append(o['framewise_output'].numpy())
    want_clip, want_frame = np.concatenate(want_clip), np.concatenate(want_frame)
    assert np.max(np.abs(out['clipwise_output'] - want_clip) / np.maximum(np.abs(want_clip), 1e-6)) <= 1e-3
    assert np.max(np.abs(out['framewise_output'] - want_frame) / np.maximum(np.abs(want_frame), 1e-6)) <= 2e-3
    # frames_num='target': framewise_output takes strong_target's length (evaluate.py:19 asserts equal shapes)
    padded = pu.forward(model=mine, data_loader=loader(), return_target=True, frames_num='target')
    assert padded['framewise_output'].shape == padded['strong_target'].shape == (7, 101, 17)
    assert np.array_equal(padded['framewise_output'][:, :96], out['framewise_output'])
    for t in range(96, 101):
        assert np.array_equal(padded['framewise_output'][:, t], out['framewise_output'][:, 95])
    # int16 PCM batches (the clip store's native rows): x/32767 on the device, bit-identical outputs
    store = dg.ClipStore(path)
    pcm_batches = [store.gather(np.arange(lo, min(lo + 3, 7))) for lo in range(0, 7, 3)]
    assert pcm_batches[0]['waveform'].dtype == np.int16
    out16 = pu.forward(model=mine, data_loader=pcm_batches, return_input=True)
    assert np.array_equal(out16['clipwise_output'], out['clipwise_output'])
    assert np.array_equal(out16['framewise_output'], out['framewise_output'])
    assert out16['waveform'].dtype == np.int16 and out16['waveform'].shape == (7, 32000)
    assert np.array_equal(pu.pad_framewise_output(out['framewise_output'], 50), out['framewise_output'][:, :50])


@pytest.mark.gpu
def test_two_rank_graph_replay_equals_two_rank_eager(tmp_path, run_two_ranks):
    """World size 2 with ``use_graph=True``: forward + backward replay as one CUDA graph per rank, the all-reduce and
    the Adam kernel follow as eager launches (trainer.FusedTrainer).  Four steps must leave bit-identical parameters,
    gradients and losses to four eager steps, on both ranks."""
    (tmp_path / 'eager').mkdir()
    (tmp_path / 'graph').mkdir()
    e0, e1 = run_two_ranks('gpu4', tmp_path / 'eager')
    g0, g1 = run_two_ranks('gpu4_graph', tmp_path / 'graph')
    assert g0['graphs'] == 1 and g1['graphs'] == 1 and e0['graphs'] == 0
    assert g0['steps'] == 4 and e0['steps'] == 4
    for e, g in ((e0, g0), (e1, g1)):
        assert e['losses'] == g['losses']
        assert torch.equal(e['grad'], g['grad']) and torch.equal(e['param'], g['param'])
    assert torch.equal(g0['param'], g1['param'])
