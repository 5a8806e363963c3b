"""CPU-only tests of the product's host logic (no kernels run): frozen-parameter construction,
SpecAugment RNG replay, C-ABI export surface."""
import ctypes
import os
import re
import time

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frozen_parameters_equal_oracle():
    from oracle import frontend as ofe
    from sound_event_detection_dcase2017_task4_b200 import frontend as pfe
    assert np.array_equal(ofe.slaney_mel_filterbank(32000, 1024, 64, 50, 14000).T,
                          pfe.mel_weight_matrix(32000, 1024, 64, 50, 14000))
    wr, wi = ofe.hann_dft_conv_weights(1024, 1024)
    pr, pi = pfe.dft_conv_weights(1024, 1024)
    assert np.array_equal(wr, pr) and np.array_equal(wi, pi)


def test_mel_csr_is_exact():
    from sound_event_detection_dcase2017_task4_b200 import frontend as pfe
    w = pfe.mel_weight_matrix(32000, 1024, 64, 50, 14000)
    bank = pfe.MelBankCSR(torch.from_numpy(w))
    assert bank.w.numel() == 866 and bank.n_bins == 513 and bank.n_mels == 64
    dense = np.zeros_like(w)
    lo, off, taps = bank.lo.numpy(), bank.off.numpy(), bank.w.numpy()
    for m in range(64):
        dense[lo[m]:lo[m] + off[m + 1] - off[m], m] = taps[off[m]:off[m + 1]]
    assert np.array_equal(dense, w)


def test_int16_fp32_division_matches_reference_conversion():
    """The kernel converts PCM with a correctly-rounded fp32 division; utils/utilities.py:66-67
    divides in float64 and casts.  They agree on every int16 value."""
    pcm = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    ref = (pcm / 32767.).astype(np.float32)
    dev = pcm.astype(np.float32) / np.float32(32767.0)
    assert np.array_equal(ref, dev)


def _round_to_f32(fr):
    """Fraction -> nearest float32, ties to even (normal range): exact, no double rounding."""
    import math
    from fractions import Fraction
    if fr == 0:
        return np.float32(0)
    a = abs(fr)
    e = math.floor(math.log2(a))
    while Fraction(2) ** e > a:
        e -= 1
    while Fraction(2) ** (e + 1) <= a:
        e += 1
    ulp = Fraction(2) ** (e - 23)
    n = a / ulp
    lo = n.numerator // n.denominator
    rem = n - lo
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and lo % 2 == 1):
        lo += 1
    return np.float32((-1 if fr < 0 else 1) * float(lo * ulp))


def test_int16_fma_quotient_is_the_correctly_rounded_division():
    """csrc/logmel.cu Sample<int16_t>::cvt replaces the IEEE division by q0 = v * r, e = fma(-q0, 32767, v),
    q = fma(e, r, q0) with r = fl(1 / 32767).  Replayed here in exact rational arithmetic with one rounding per
    instruction: it equals the reference conversion (utils/utilities.py:66-67) for every int16 value."""
    from fractions import Fraction
    r = _round_to_f32(Fraction(1, 32767))
    assert r == np.float32(1.0) / np.float32(32767.0)
    fr = Fraction(float(r))
    pcm = np.arange(-32768, 32768, dtype=np.int64)
    ref = (pcm / 32767.).astype(np.float32)
    for v, want in zip(pcm.tolist(), ref):
        q0 = Fraction(float(_round_to_f32(Fraction(v) * fr)))
        e = Fraction(float(_round_to_f32(Fraction(v) - q0 * 32767)))
        assert _round_to_f32(e * fr + q0) == want, v


def test_specaug_replay_is_bit_exact_and_advances_generator():
    from oracle import frontend as ofe
    from sound_event_detection_dcase2017_task4_b200 import specaug
    for seed in (1, 7, 2024):
        torch.manual_seed(seed)
        t_ref = ofe.draw_stripes(37, 1001, 64, 2).numpy()
        f_ref = ofe.draw_stripes(37, 64, 8, 2).numpy()
        tail_ref = torch.randint(0, 1 << 30, (4,))
        torch.manual_seed(seed)
        t, f = specaug.draw_spec_augment(37, 1001, 64)
        tail = torch.randint(0, 1 << 30, (4,))
        assert np.array_equal(t, t_ref) and np.array_equal(f, f_ref)      # index-valued: bit exact
        assert torch.equal(tail, tail_ref)                                # global RNG stream preserved
    assert specaug._fast_ok is True


def test_specaug_replay_crosses_mt19937_refill():
    from oracle import frontend as ofe
    from sound_event_detection_dcase2017_task4_b200 import specaug
    torch.manual_seed(3)
    ref = ofe.draw_stripes(400, 1001, 64, 2).numpy()          # 1600 draws > 624-word state
    torch.manual_seed(3)
    got = specaug.draw_stripes(400, 1001, 64, 2)
    assert np.array_equal(got, ref)


def test_specaug_replay_is_fast():
    from sound_event_detection_dcase2017_task4_b200 import specaug
    specaug.draw_spec_augment(4, 1001, 64)
    t0 = time.perf_counter()
    specaug.draw_spec_augment(512, 1001, 64)
    assert time.perf_counter() - t0 < 0.05


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'sed_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sed_[a-z0-9_]+)\s*\(', text)))


def test_c_abi_exports_every_declared_symbol():
    from sound_event_detection_dcase2017_task4_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from sound_event_detection_dcase2017_task4_b200 import build
        build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(handle, name), 'declared in include/sed_b200.h but not exported: %s' % name
    assert sorted(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ set(declared)
    assert handle.sed_abi_version() == 1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'sound_event_detection_dcase2017_task4_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_missing_library_fails_loudly(monkeypatch):
    from sound_event_detection_dcase2017_task4_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libsedb200.so')
    with pytest.raises(RuntimeError, match='no CPU'):
        _lib.lib()


def test_cpu_tensor_is_rejected():
    from sound_event_detection_dcase2017_task4_b200 import frontend as pfe
    bank = pfe.MelBankCSR(torch.from_numpy(pfe.mel_weight_matrix(32000, 1024, 64, 50, 14000)))
    with pytest.raises(RuntimeError, match='CUDA'):
        pfe.logmel(torch.zeros(1, 6400), 320, bank)


def test_launcher_resolves_reference_imports_to_the_dropin():
    """launch.prepare() orders sys.path the way pytorch/main.py:3,18-27 needs so that its bare-name
    imports bind to the B200 implementations and everything else stays the reference's own file."""
    ref = '/root/reference'
    if not os.path.isdir(os.path.join(ref, 'pytorch')):
        pytest.skip('reference checkout not present (GPU box)')
    import subprocess
    import sys
    code = (
        "import sys, os\n"
        "sys.path.insert(0, %r)\n"
        "from sound_event_detection_dcase2017_task4_b200 import launch\n"
        "script = launch.prepare(%r)\n"
        "assert script.endswith('pytorch/main.py') and sys.path[0] == launch.DROPIN_DIR\n"
        "import models, losses, pytorch_utils, config\n"
        "from torchlibrosa.stft import Spectrogram, LogmelFilterBank\n"
        "from torchlibrosa.augmentation import SpecAugmentation\n"
        "for m in (models, losses, pytorch_utils, sys.modules['torchlibrosa']):\n"
        "    assert m.__file__.startswith(launch.DROPIN_DIR), m.__file__\n"
        "assert config.__file__.startswith(%r)\n"
        "names = ['Cnn_9layers_FrameMax', 'Cnn_9layers_FrameAvg', 'Cnn_9layers_FrameAtt', 'Cnn_9layers_Gru_FrameAvg',\n"
        "         'Cnn_9layers_Gru_FrameAtt', 'Cnn_9layers_Transformer_FrameAvg', 'Cnn_9layers_Transformer_FrameAtt',\n"
        "         'ConvBlock', 'AttBlock', 'MultiHead', 'init_layer', 'init_bn', 'init_gru', 'interpolate']\n"
        "assert all(hasattr(models, n) for n in names)\n"
        "assert callable(losses.get_loss_func('clip_bce')) and callable(pytorch_utils.do_mixup)\n"
        "Model = eval('models.' + 'Cnn_9layers_FrameAvg')\n"
        "m = Model(config.sample_rate, config.window_size, config.hop_size, config.mel_bins, config.fmin, config.fmax,\n"
        "          config.classes_num)\n"
        "assert 'spectrogram_extractor.stft.conv_real.weight' in m.state_dict()\n"
        "print('ok')\n" % (ROOT, ref, ref))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stderr[-2000:]


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import it (a product path that routes
    through the oracle would void every parity claim)."""
    pkg = os.path.join(ROOT, 'sound_event_detection_dcase2017_task4_b200')
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                with open(os.path.join(dirpath, f)) as fh:
                    for n, line in enumerate(fh, 1):
                        if re.match(r'\s*(from|import)\s+oracle\b', line):
                            offenders.append('%s:%d' % (os.path.join(dirpath, f), n))
    assert offenders == []


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([x for x in r.stdout.splitlines() if x.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['unit'] == 'clips/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'Cnn_9layers_Gru_FrameAtt' in line['metric'] and 'Cnn_9layers_Gru_FrameAtt' in line['config']['workload']


def test_abi_rejects_bad_arguments_before_touching_the_device():
    """Error contract of the C ABI (include/sed_b200.h): non-zero status, a thread-local message, no exception and --
    for argument errors -- no CUDA call at all, so these run without a GPU."""
    import threading
    from sound_event_detection_dcase2017_task4_b200 import _lib
    L = _lib.lib()

    def msg():
        return (L.sed_last_error_string() or b'').decode()

    assert L.sed_conv3x3_tc_fwd(0, 0, 0, 0, 1, 10, 16, 64, 64, 0) != 0 and 'null pointer' in msg()
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.addressof(buf)
    assert L.sed_conv3x3_tc_fwd(p, p, p, 0, 1, 10, 12, 64, 64, 0) != 0 and 'W=12' in msg()
    assert L.sed_conv3x3_tc_fwd(p, p, p, 0, 1, 10, 16, 96, 64, 0) != 0 and 'Cin=96' in msg()
    assert L.sed_conv3x3_tc2_fwd(p, p, p, 0, 1, 10, 16, 64, 192, 0) != 0 and 'Cout=192' in msg()
    assert L.sed_logmel_f32(p, 1, 400, 320, p, p, p, 64, 1e-10, 0.0, p, 0) != 0 and 'reflect padding' in msg()
    assert L.sed_logmel_f32(p, 1, 32000, 321, p, p, p, 64, 1e-10, 0.0, p, 0) != 0 and 'hop must be even' in msg()
    assert L.sed_vad_count(p, 0, 1, 10, 0, 0, p, 0, p, p, p, 0, 0) != 0 and 'bad shape' in msg()
    assert L.sed_conv3x3_tc_dgrad_bnr(p, p, p, 1, 10, 16, 64, 64, p, 10, p, p, 3, p, 0) != 0 and 'pool=3' in msg()
    with pytest.raises(RuntimeError, match='sed_conv3x3_tc_fwd failed'):
        _lib.call('sed_conv3x3_tc_fwd', 0, 0, 0, 0, 1, 10, 16, 64, 64, 0)
    # the message is per thread
    assert L.sed_conv3x3_tc_fwd(p, p, p, 0, 1, 10, 12, 64, 64, 0) != 0
    seen = []
    t = threading.Thread(target=lambda: seen.append(msg()))
    t.start()
    t.join()
    assert 'W=12' in msg() and 'W=12' not in seen[0]


def test_product_mixup_matches_reference_golden(golden):
    """Row a6: the product's own ``utilities.Mixup`` (utils/utilities.py:220-242) against lambdas minted from
    the UNMODIFIED reference class (tests/golden/make_golden.py); bulk draws equal the per-pair draws."""
    from sound_event_detection_dcase2017_task4_b200.utilities import Mixup, int16_to_float32
    g, _ = golden
    m = Mixup(mixup_alpha=1., random_seed=1234)
    a = m.get_lambda(batch_size=32)
    assert a.dtype == np.float64 and a[0] == 0.23538938957272115
    assert np.array_equal(a, g['mixup_lambda_first32'])
    assert np.array_equal(m.get_lambda(6), g['mixup_lambda_next6'])
    # odd batch sizes round up to whole pairs, like the reference's range(0, n, 2) loop
    ref = np.random.RandomState(7)
    want = []
    for _ in range(3):
        lam = ref.beta(0.5, 0.5, 1)[0]
        want += [lam, 1. - lam]
    assert np.array_equal(Mixup(0.5, 7).get_lambda(5), np.array(want))
    # in-place float32 fill == the cast of move_data_to_device (main.py:238)
    buf = np.empty(32, dtype=np.float32)
    Mixup(1., 1234).fill_lambda(buf)
    assert np.array_equal(buf, g['mixup_lambda_first32'].astype(np.float32))
    pcm = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    assert np.array_equal(int16_to_float32(pcm), g['int16_to_float32_all'])


def _emulated_replica(network):
    """What torch.nn.parallel.replicate() builds for one device (torch 2.11), on CPU: modules shallow-copied by
    ``_replicate_for_data_parallel`` (empty ``_parameters``), every weight re-attached as a plain NON-LEAF tensor
    attribute and listed in ``_former_parameters``."""
    from collections import OrderedDict
    params = list(network.parameters())
    copies = [p * 1 if p.requires_grad else p.detach().clone() for p in params]
    index = {p: i for i, p in enumerate(params)}
    modules = list(network.modules())
    where = {m: i for i, m in enumerate(modules)}
    reps = [m._replicate_for_data_parallel() for m in modules]
    for r in reps:
        r._former_parameters = OrderedDict()
    for i, m in enumerate(modules):
        for key, child in m._modules.items():
            if child is None:
                reps[i]._modules[key] = None
            else:
                setattr(reps[i], key, reps[where[child]])
        for key, p in m._parameters.items():
            if p is None:
                reps[i]._parameters[key] = None
            else:
                setattr(reps[i], key, copies[index[p]])
                reps[i]._former_parameters[key] = copies[index[p]]
        for key, b in m._buffers.items():
            setattr(reps[i], key, None if b is None else b.clone())
    return reps[0], copies


@pytest.mark.parametrize('name', ['Cnn_9layers_Gru_FrameAtt', 'Cnn_9layers_Transformer_FrameAvg'])
def test_trainable_tensors_are_found_on_dataparallel_replicas(name):
    """nn.DataParallel (main.py:138) runs forward on replicas whose parameters() is EMPTY.  The
    model must still hand every weight that needs gradient to its autograd.Function -- otherwise loss.backward()
    (main.py:257) has nothing to differentiate."""
    from sound_event_detection_dcase2017_task4_b200 import models
    torch.manual_seed(0)
    net = getattr(models, name)(32000, 1024, 320, 64, 50, 14000, 17)
    want = [p for p in net.parameters() if p.requires_grad]
    assert [id(p) for p in models.trainable_tensors(net)] == [id(p) for p in want]
    replica, copies = _emulated_replica(net)
    assert list(replica.parameters()) == []                                    # the trap
    got = models.trainable_tensors(replica)
    expect = [c for c, p in zip(copies, net.parameters()) if p.requires_grad]
    assert len(got) == len(want) and [id(t) for t in got] == [id(t) for t in expect]
    assert all((not t.is_leaf) and t.requires_grad for t in got)
    # the tensors the engine reads through attribute access are those same objects
    assert replica.conv_block3.conv2.weight is got[[id(p) for p in want].index(id(net.conv_block3.conv2.weight))]


def test_gradient_exchange_world_size_2_gloo(tmp_path, run_two_ranks):
    """Row a17 / (e) on CPU: two gloo ranks.  Every rank scales its loss by 1/world, ONE all-reduce(sum) of the flat
    gradient then holds the gradient of the global-batch mean loss on every rank -- what nn.DataParallel's reduce
    gives for equal shards (main.py:138).  Shards are contiguous, equal and even-sized (mixup pairs stay together)."""
    r0, r1 = run_two_ranks('cpu', tmp_path)
    want = (r0['local'] + r1['local']) / 2
    assert torch.equal(r0['reduced'], r1['reduced'])
    assert torch.allclose(r0['reduced'], want, rtol=0, atol=1e-7)
    assert r0['bounds'] == (0, 32) and r1['bounds'] == (32, 64)


def test_shard_bounds_keep_mixup_pairs_together():
    from sound_event_detection_dcase2017_task4_b200.trainer import shard_bounds
    for world in (1, 2, 4, 8):
        edges = [shard_bounds(512 * world, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 512 * world
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert all((hi - lo) % 2 == 0 and lo % 2 == 0 for lo, hi in edges)
    with pytest.raises(ValueError):
        shard_bounds(6, 4, 0)
