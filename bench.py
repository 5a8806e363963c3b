"""bench.py -- the hot path's headline metric on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model NAME] [--batch B] [--workload train|logmel|eval]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric "10s-clips/sec train Cnn_9layers_Gru_FrameAtt"): one training step
of Cnn_9layers_Gru_FrameAtt as /root/reference/pytorch/main.py:233-258 runs it with
``--batch_size 256 --augmentation mixup``: 512 raw 32 kHz x 10 s clips per GPU per step go through
log-mel / bn0 / SpecAugment, are mixed pairwise to 256 training samples, 4 ConvBlocks, biGRU,
attention head, clip_bce, backward, one NCCL all-reduce of the flat gradient (N > 1),
Adam(amsgrad).  ``value`` counts RAW 10 s clips consumed per second over all GPUs (the loader
feeds 2 x batch_size clips per iteration under mixup, main.py:151-153); training samples/s is
half of it and is reported as ``train_samples_per_s``.

One JSON line on stdout (rank 0).  ``value``: inputs resident in HBM; ``e2e``: the same step with
the waveforms/targets/lambda copied from pinned host memory inside the timed region and the loss
read back to the host every step.  ``roofline``: the 3x3 tensor-core convolution kernel
(conv3x3_halo2_kernel, CTA pairs: all forward + data-gradient launches of a step) against the measured bf16
peak.  ``cpu_baseline`` / ``--impl reference``: the CPU oracle (oracle/sed.py, a restatement of
the reference's PyTorch modules pinned to the unmodified reference by tests/golden) timed on
this box's host cores on a bounded sample of the same workload.

Other BASELINE.json configurations (the default invocation above is unchanged by these switches):
  --model Cnn_9layers_FrameAvg                            config 2 (batch_size 256 + mixup, 1 GPU)
  --model Cnn_9layers_Transformer_FrameAvg --batch 128    config 4
  --workload logmel                                       config 5's headline point: log-mel only, 512 raw 10 s clips per
                                                          GPU per launch, Mframes/s against the HBM roofline (replicas)
  --workload eval                                         row f3: pytorch_utils.forward over host int16 batches, clips/s
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = 'Cnn_9layers_Gru_FrameAtt'
CTOR = (32000, 1024, 320, 64, 50, 14000, 17)
CLIP_SAMPLES = 320000
UNIT = 'clips/s'


def metric_name(model):
    # shared by both arms: the per-arm batch size is a property of the run (config.batch_size), not of the metric
    return '10s-clips/sec train %s (raw 10 s clips consumed per second, mixup)' % model

def train_config(model, bs, world):
    """`config` of the train workload -- identical in both arms (the reference arm times a bounded SAMPLE of it and
    says so in cpu_baseline.sample)."""
    b2 = 2 * bs
    return {'workload': '%s train step (log-mel, bn0, SpecAugment, mixup, 4 ConvBlocks, %s, clip_bce, backward, '
                        'Adam-amsgrad), batch_size %d + mixup = %d raw 32 kHz x 10 s clips per GPU per step'
                        % (model, head_words(model), bs, b2),
            'batch_size_per_gpu': bs, 'raw_clips_per_gpu_per_step': b2,
            'parallelism': 'dp%d, one NCCL all-reduce of the flat fp32 gradient per step' % world,
            'precision': 'bf16 tensor-core operands / activations, fp32 accumulation, fp32 front-end, '
                         'BN statistics, GRU, heads, loss, optimizer',
            'l2': 'inputs larger than L2 (waveforms %.0f MB per step, activations > 10 GB)'
                  % (b2 * CLIP_SAMPLES * 4 / 1e6)}


def head_words(model):
    temporal = 'biGRU' if '_Gru_' in model else ('multi-head self-attention' if '_Transformer_' in model else 'no temporal module')
    pool = {'Att': 'attention head', 'Avg': 'fc + mean-over-time head', 'Max': 'fc + max-over-time head'}[model[-3:]]
    return '%s, %s' % (temporal, pool)


# 3x3 conv FLOPs per training sample of one pass: layers with Cin >= 64
CONV_TC_LAYERS = [(1001, 64, 64, 64), (500, 32, 64, 128), (500, 32, 128, 128), (250, 16, 128, 256),
                  (250, 16, 256, 256), (125, 8, 256, 512), (125, 8, 512, 512)]


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_burst': p['bf16_tflops'],
                'bf16_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_burst': 1590.0, 'bf16_sustained': 1400.0, 'source': 'fallback'}


def profile_value(name, key='dram_bytes_per_launch_avg'):
    """A per-launch figure from a committed ncu summary under profiles/ (None if absent)."""
    path = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(key)


def conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the conv kernel (average over the 14 forward /
    data-gradient launches of a step) from the committed `ncu --set full` capture, or None."""
    path = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get('dram_bytes_per_launch_avg')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'power_w_max': max(power),
                'samples': len(sm), 'reasons': sorted(reasons)}


def synthetic_rank_batch(b2, rank):
    """int16 uniform [-8192, 8191] -> x/32767 fp32, Bernoulli(0.067) targets,
    RandomState(1234 + rank)."""
    import numpy as np
    rs = np.random.RandomState(1234 + rank)
    pcm = rs.randint(-8192, 8192, size=(b2, CLIP_SAMPLES), dtype=np.int16)
    target = (rs.rand(b2, 17) < 0.067).astype(np.float32)
    return pcm, target


# ----------------------------------------------------------------------------- reference / CPU arm
def cpu_oracle_run(steps, warmup, batch_size, threads, model_name=MODEL):
    """Times the CPU oracle's train step (same model, mixup, Adam-amsgrad) on a bounded sample:
    ``batch_size`` training samples = 2*batch_size raw clips per step.  Returns clips/s."""
    import numpy as np
    import torch
    from oracle import sed
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = sed.build(model_name)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.,
                           amsgrad=True)
    b2 = 2 * batch_size
    _, wave, target = sed.synthetic_batch(b2, CLIP_SAMPLES, seed=1234)
    wave, target = torch.from_numpy(wave), torch.from_numpy(target)
    mix = sed.MixupLambda(1., 1234)
    times = []
    for i in range(warmup + steps):
        lam = torch.Tensor(mix.get_lambda(b2))
        t0 = time.perf_counter()
        sed.train_step(model, opt, wave, target, lam)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return b2 * len(times) / total, total / len(times)


def run_reference(args):
    """The reference arm: the CPU restatement of the reference's own modules (oracle/sed.py; the reference cannot be
    installed -- pure Python + an un-vendored dependency) on the box's host cores, same metric and
    config as our arm, each step a bounded sample (batch_size 8 + mixup = 16 raw clips instead of 512)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    if args.workload != 'train':
        print(json.dumps({'impl': 'reference', 'unavailable': 'the reference arm covers the train workload only'}), flush=True)
        return 0
    threads = os.cpu_count() or 1
    bs = 8 if args.steps <= 24 else (4 if args.steps <= 60 else 2)
    value, sec = cpu_oracle_run(args.steps, args.warmup, bs, threads, args.model)
    line = {
        'impl': 'reference', 'metric': metric_name(args.model), 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': train_config(args.model, args.batch, args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': 'CPU oracle (oracle/sed.py, PyTorch CPU fp32 restatement pinned to the '
                                   'unmodified reference), %d timed steps of batch_size %d + mixup = %d raw clips/step '
                                   '(a bounded sample of the %d-clip step), fp32'
                                   % (args.steps, bs, 2 * bs, 2 * args.batch)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- log-mel-only workload (config 5)
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device; the product path has no CPU fallback '
                           '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    return world, rank, local_rank, dev


def _timed_region(fn, steps, world, dev):
    """EXACTLY `steps` calls between barrier + synchronize pairs; device time, max over ranks."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_logmel(args):
    """BASELINE.json config 5, headline point of the sweep (tools/logmel_sweep.py walks the rest): Spectrogram +
    LogmelFilterBank (models.py:199-200) alone on 512 raw 32 kHz x 10 s clips per GPU per launch.  The path shards by
    clip with no exchange: N GPUs = N replicas on their own clips (`scaling: weak`, no collective)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world, rank, local_rank, dev = _dist_setup()
    from sound_event_detection_dcase2017_task4_b200 import _lib, frontend as fe
    _lib.lib()
    b2 = 2 * args.batch
    bank = fe.MelBankCSR(torch.from_numpy(fe.mel_weight_matrix(*[CTOR[i] for i in (0, 1, 3, 4, 5)])).to(dev))
    pcm, _ = synthetic_rank_batch(b2, rank)
    pcm_host = torch.from_numpy(pcm).pin_memory()
    wave_host = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).pin_memory()
    wave_dev = wave_host.to(dev)
    frames_per_clip = CLIP_SAMPLES // CTOR[2] + 1
    out = torch.empty((b2, 1, frames_per_clip, 64), device=dev)
    out_host = torch.empty((b2, 1, frames_per_clip, 64)).pin_memory()
    stage = torch.empty_like(wave_dev)
    stage16 = torch.empty((b2, CLIP_SAMPLES), dtype=torch.int16, device=dev)

    def resident():
        fe.logmel(wave_dev, CTOR[2], bank, out=out)

    def e2e():
        stage.copy_(wave_host, non_blocking=True)
        fe.logmel(stage, CTOR[2], bank, out=out)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def e2e_i16():
        stage16.copy_(pcm_host, non_blocking=True)
        fe.logmel(stage16, CTOR[2], bank, out=out)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(args.warmup, 3)):
        resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count()
    ms = _timed_region(resident, args.steps, world, dev)
    launches = _lib.launch_count() - n0
    clock_info = clocks.stop() if rank == 0 else None
    stage16.copy_(pcm_host)

    def resident_i16():
        fe.logmel(stage16, CTOR[2], bank, out=out)

    for _ in range(2):
        resident_i16()
    ms_i16 = _timed_region(resident_i16, args.steps, world, dev)
    for _ in range(2):
        e2e()
    ms_e2e = _timed_region(e2e, args.steps, world, dev)
    for _ in range(2):
        e2e_i16()
    ms_e2e16 = _timed_region(e2e_i16, args.steps, world, dev)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import frontend as ofe
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        spec = ofe.Spectrogram(n_fft=1024, hop_length=320, win_length=1024, window='hann', center=True,
                               pad_mode='reflect', freeze_parameters=True)
        lmel = ofe.LogmelFilterBank(sr=32000, n_fft=1024, n_mels=64, fmin=50, fmax=14000, ref=1.0, amin=1e-10,
                                    top_db=None, freeze_parameters=True)
        x = wave_host[:8].clone()
        with torch.no_grad():
            lmel(spec(x))
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 10.0:
                lmel(spec(x))
                reps += 1
            sec = (time.perf_counter() - t0) / reps
        cpu = {'value': 8 * frames_per_clip / sec / 1e6, 'unit': 'Mframes/s', 'cores': threads, 'kind': 'port',
               'sample': 'oracle/frontend.py Spectrogram + LogmelFilterBank (the restated torchlibrosa 0.0.4 conv1d DFT '
                         '+ dense mel matmul), batch of 8 x 10 s clips, %d repeats, %.1f ms each' % (reps, sec * 1e3)}
    if rank == 0:
        peaks = measured_peaks()
        frames = b2 * frames_per_clip * world * args.steps
        bytes_per_launch = b2 * (CLIP_SAMPLES * 4 + frames_per_clip * 64 * 4)
        ms_launch = ms / args.steps
        achieved = bytes_per_launch / ms_launch / 1e6
        line = {
            'metric': 'logmel Mframes/s (Spectrogram + LogmelFilterBank, 32 kHz x 10 s clips)', 'value': frames / ms / 1e3,
            'unit': 'Mframes/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_launch, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': 'log-mel only (BASELINE config 5, sweep point 10 s x %d clips per GPU): framed rFFT-1024, '
                                   'power, sparse mel, 10 log10; fp32 waveform in, fp32 (T, 64) out' % b2,
                       'clips_per_gpu_per_step': b2, 'parallelism': 'replicas x%d, no collective' % world,
                       'l2': 'inputs larger than L2 (%.0f MB of waveforms per launch)' % (b2 * CLIP_SAMPLES * 4 / 1e6)},
            'e2e': {'value': frames / ms_e2e / 1e3, 'unit': 'Mframes/s', 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': int(wave_host.numel() * 4), 'd2h_bytes_per_step': int(out_host.numel() * 4),
                    'input': 'fp32 waveforms in pinned host memory -> device -> kernel -> pinned host log-mel'},
            'resident_int16': {'value': frames / ms_i16 / 1e3, 'unit': 'Mframes/s', 'ms_per_step': ms_i16 / args.steps,
                               'input': 'int16 PCM resident in HBM (x / 32767 fused into the gather, bit-identical output)'},
            'e2e_int16': {'value': frames / ms_e2e16 / 1e3, 'unit': 'Mframes/s', 'ms_per_step': ms_e2e16 / args.steps,
                          'h2d_bytes_per_step': int(pcm_host.numel() * 2), 'd2h_bytes_per_step': int(out_host.numel() * 4)},
            'gpu_launches': launches, 'clocks': clock_info,
            'roofline': {'kernel': 'logmel_warp_kernel<float> (sed_logmel_f32)', 'bound': 'hbm', 'achieved': round(achieved, 1),
                         'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': round(achieved / peaks['hbm_gbs'], 4),
                         'peak_source': peaks['source'] + ' hbm_gbs', 'bytes_per_launch': bytes_per_launch,
                         'avg_launch_ms': round(ms_launch, 4), 'traffic': profile_value('logmel_traffic.json')},
            'cpu_baseline': cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_eval(args):
    """Row f3: the inference loop (pytorch_utils.forward, /root/reference/pytorch/pytorch_utils.py:25-77) over host
    batches of int16 PCM, eval-mode model, outputs drained to host numpy arrays -- 10 s clips per second."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world, rank, local_rank, dev = _dist_setup()
    from sound_event_detection_dcase2017_task4_b200 import _lib, models, pytorch_utils
    _lib.lib()
    bs = args.batch
    torch.manual_seed(0)
    model = getattr(models, args.model)(*CTOR).to(dev)
    pcm, target = synthetic_rank_batch(bs, rank)
    names = np.array(['Y%05d.wav' % i for i in range(bs)])
    batches = [{'audio_name': names, 'waveform': pcm, 'target': target}] * 4

    def loop():
        out = pytorch_utils.forward(model, batches, return_target=True)
        assert out['framewise_output'].shape == (4 * bs, 1000, 17)

    for _ in range(max(2, min(args.warmup, 3))):
        loop()
    n0 = _lib.launch_count()
    ms = _timed_region(loop, args.steps, world, dev)
    launches = _lib.launch_count() - n0
    if rank == 0:
        clips = 4 * bs * world * args.steps
        line = {'metric': '10s-clips/sec inference %s (pytorch_utils.forward, host int16 in, host numpy out)' % args.model,
                'value': clips / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': 'eval-mode forward loop, 4 batches of %d raw 10 s clips per step' % bs,
                           'batch_size_per_gpu': bs},
                'e2e': {'value': clips / (ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(4 * pcm.nbytes),
                        'd2h_bytes_per_step': int(4 * bs * (1000 * 17 + 17) * 4)},
                'gpu_launches': launches}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device; the product path has no CPU fallback '
                           '(use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    from sound_event_detection_dcase2017_task4_b200 import _lib, models
    from sound_event_detection_dcase2017_task4_b200.trainer import FusedTrainer
    from sound_event_detection_dcase2017_task4_b200.utilities import Mixup
    _lib.lib()
    MODEL = args.model

    bs = args.batch
    b2 = 2 * bs
    torch.manual_seed(0)
    model = getattr(models, MODEL)(*CTOR).to(dev)
    model.train()
    use_graph = not args.no_graph
    trainer = FusedTrainer(model, lr=1e-3, world_size=world, use_graph=use_graph)

    pcm, target_np = synthetic_rank_batch(b2, rank)
    wave_host = torch.from_numpy((pcm / np.float32(32767.)).astype(np.float32)).pin_memory()
    target_host = torch.from_numpy(target_np).pin_memory()
    lam_gen = Mixup(1., 1234 + rank)                 # utils/utilities.py:220-242 (host-side numpy stream, main.py:233)
    lam_host = torch.empty(b2, dtype=torch.float32).pin_memory()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    wave_dev = wave_host.to(dev)
    target_dev = target_host.to(dev)
    lam_dev = torch.empty(b2, dtype=torch.float32, device=dev)

    def next_lambda():
        lam_gen.fill_lambda(lam_host.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_step = {}

    def timed(fn, steps, tag=None):
        """EXACTLY `steps` steps between two barrier + synchronize pairs, device time (CUDA events on the compute
        stream); one extra event per step gives the per-step distribution (median / min) without adding a sync."""
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        n0 = _lib.launch_count()
        marks[0].record()
        for i in range(steps):
            fn()
            marks[i + 1].record()
        barrier()
        ms = marks[0].elapsed_time(marks[-1])
        if trainer.use_graph:                              # replays do not pass through the C-ABI launch counter
            n0 -= steps * trainer.graph_launches
        if tag:
            d = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
            per_step[tag] = {'median_ms': round(statistics.median(d), 4), 'min_ms': round(d[0], 4), 'max_ms': round(d[-1], 4)}
        launches = _lib.launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    def step_resident():
        next_lambda()
        lam_dev.copy_(lam_host, non_blocking=True)      # 2 KB of lambdas: host-generated in the reference too
        trainer.step(wave_dev, target_dev, lam_dev)

    stage = {'slot': 0}
    from sound_event_detection_dcase2017_task4_b200.feed import DeviceFeed
    feed = DeviceFeed(dev, slots=2)
    lam_slots = [torch.empty(b2, dtype=torch.float32).pin_memory() for _ in range(2)]

    pcm_host = torch.from_numpy(pcm).pin_memory()          # the on-disk format (int16 PCM, utils/features.py:238-241)
    host_wave = {'cur': wave_host}

    def submit(slot):
        """Step inputs: waveforms + targets + this step's mixup lambdas, pinned host -> device slot."""
        lam_gen.fill_lambda(lam_slots[slot].numpy())
        feed.submit(slot, {'waveform': host_wave['cur'], 'target': target_host, 'lam': lam_slots[slot]})

    def step_e2e():
        """One iteration of main.py:233-258 with HOST inputs: every step copies its own batch (on the copy
        stream, overlapped with the previous step's kernels) and reads its loss back (main.py:253 prints it)."""
        cur = stage['slot']
        submit(cur ^ 1)                                  # next step's batch: H2D in flight during this step
        batch = feed.acquire(cur)
        loss = trainer.step(batch['waveform'], batch['target'], batch['lam'])
        feed.release(cur)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        stage['loss'] = float(loss_host)
        stage['slot'] = cur ^ 1

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, launches = timed(step_resident, args.steps, 'resident')
    clock_info = clocks.stop() if rank == 0 else None
    submit(0)                                            # prologue: first batch (its copy is outside the timed region;
    for _ in range(2):                                   #  each timed step issues exactly one batch copy)
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps, 'e2e')
    # the same loop fed with int16 PCM (x / 32767 fused into the log-mel kernel's gather: bit-identical log-mel)
    host_wave['cur'] = pcm_host
    stage['slot'] = 0
    submit(0)
    for _ in range(2):
        step_e2e()
    ms_e2e_i16, _ = timed(step_e2e, args.steps)

    # ---- roofline leg: per-entry-point device times of one more step (CUDA events on the launch stream)
    roof = None
    shares = None
    # every rank runs these steps (the step contains the all-reduce); only rank 0 records the events
    agg = {}
    reps = 3
    from sound_event_detection_dcase2017_task4_b200 import engine as _engine
    _overlap_was = _engine.OVERLAP_WGRAD
    _engine.OVERLAP_WGRAD = False                        # serial schedule: every kernel timed alone on one stream
    _graph_was, trainer.use_graph = trainer.use_graph, False   # eager launches: the per-call events need them
    for _ in range(reps):
        _lib.PROFILE = [] if rank == 0 else None
        step_resident()
        torch.cuda.synchronize()
        if rank == 0:
            for name, tag, e0, e1 in _lib.PROFILE:
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += e0.elapsed_time(e1)
        _lib.PROFILE = None
    _engine.OVERLAP_WGRAD = _overlap_was
    trainer.use_graph = _graph_was
    if rank == 0:
        peaks = measured_peaks()
        total_ms = sum(v[1] for v in agg.values()) / reps
        shares = {k: {'launches_per_step': v[0] // reps, 'ms_per_step': round(v[1] / reps, 4),
                      'share': round(v[1] / reps / total_ms, 4)}
                  for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        # the conv kernel runs the forward of 7 layers and the data-gradient of 7 (block1.conv2's
        # dgrad feeds the Cin=1 layer; block1.conv1 itself is a separate direct kernel)
        fl = [2.0 * h * w * ci * co * 9 for (h, w, ci, co) in CONV_TC_LAYERS]
        flops_step = bs * (sum(fl) + sum(fl))           # fwd + dgrad of all 7 tensor-core layers
        conv_keys = [k for k in ('sed_conv3x3_tc2_fwd', 'sed_conv3x3_tc2kw_fwd', 'sed_conv3x3_tc_fwd') if k in agg]
        n_conv = sum(agg[k][0] for k in conv_keys) / reps
        conv_ms = sum(agg[k][1] for k in conv_keys) / reps
        achieved = flops_step / (conv_ms * 1e-3) / 1e12
        roof = {'kernel': 'conv3x3_halo2_kernel + conv3x3_halo2_kw_kernel (sed_conv3x3_tc2_fwd / sed_conv3x3_tc2kw_fwd, CTA pairs: '
                          'forward + data-gradient launches)',
                'bound': 'tensor', 'achieved': round(achieved, 2), 'peak': peaks['bf16_sustained'],
                'unit': 'TFLOP/s', 'frac': round(achieved / peaks['bf16_sustained'], 4),
                'peak_source': peaks['source'] + ' bf16_tflops_sustained (kernel timed inside a long step)',
                'peak_burst': peaks['bf16_burst'], 'frac_of_burst': round(achieved / peaks['bf16_burst'], 4),
                'launches_per_step': n_conv, 'avg_launch_ms': round(conv_ms / n_conv, 4),
                'flop_per_launch_avg': flops_step / n_conv, 'traffic': conv_traffic(),
                'share_of_step': round(conv_ms / total_ms, 4),
                'timing': 'CUDA events around every launch of three more steps on the compute stream'}

    # ---- log-mel leg of the metric ("logmel Mframes/s"): the front-end kernel alone on this step's 512 resident clips
    logmel = None
    if rank == 0:
        from sound_event_detection_dcase2017_task4_b200 import frontend as _fe
        bank = model.logmel_extractor.mel_bank()
        lm_out = torch.empty((b2, 1, CLIP_SAMPLES // 320 + 1, 64), device=dev)
        for _ in range(3):
            _fe.logmel(wave_dev, 320, bank, out=lm_out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            _fe.logmel(wave_dev, 320, bank, out=lm_out)
        e1.record()
        torch.cuda.synchronize()
        lm_ms = e0.elapsed_time(e1) / 10
        frames = b2 * (CLIP_SAMPLES // 320 + 1)
        lm_bytes = b2 * (CLIP_SAMPLES * 4 + (CLIP_SAMPLES // 320 + 1) * 64 * 4)
        logmel = {'mframes_per_s': round(frames / lm_ms / 1e3, 1), 'ms': round(lm_ms, 4),
                  'clips': b2, 'achieved_gbs': round(lm_bytes / lm_ms / 1e6, 1), 'hbm_peak_gbs': measured_peaks()['hbm_gbs'],
                  'hbm_frac': round(lm_bytes / lm_ms / 1e6 / measured_peaks()['hbm_gbs'], 4),
                  'bytes_per_clip': lm_bytes // b2,
                  'note': 'fp32 in + fp32 log-mel out (SURVEY 8d), timed right after the training steps (clocks under the power cap); '
                          'the kernel is instruction-issue / latency bound, not HBM bound (profiles/r02_logmel_ncu.md); '
                          '`bench.py --workload logmel` times it alone'}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, sec = cpu_oracle_run(20, 2, 8, threads, MODEL)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': 'CPU oracle (oracle/sed.py), 2 warm-up + 20 timed steps of batch_size 8 + mixup '
                         '(16 raw 10 s clips/step), %.2f s/step = %.0f s of CPU work' % (sec, 20 * sec)}

    if rank == 0:
        clips = b2 * world * args.steps
        value = clips / (ms * 1e-3)
        e2e_value = clips / (ms_e2e * 1e-3)
        line = {
            'metric': metric_name(MODEL), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'train_samples_per_s': value / 2,
            'config': dict(train_config(MODEL, bs, world),
                           launch='CUDA graph replay, one graph per input-buffer set (%d kernels per step%s)' % (
                               trainer.graph_launches, '' if world == 1 else '; all-reduce and Adam launched eagerly after it')
                           if trainer.use_graph else 'eager launches through the C ABI'),
            'per_step_ms': per_step,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': int(wave_host.numel() * 4 + target_host.numel() * 4 + b2 * 4),
                    'd2h_bytes_per_step': 4, 'loss': stage.get('loss'),
                    'overlap': 'batch i+1 copied on a copy stream while step i computes (feed.DeviceFeed)',
                    'input': 'fp32 waveforms in pinned host memory (what the reference loader hands to main.py:238)'},
            'e2e_int16': {'value': clips / (ms_e2e_i16 * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e_i16 / args.steps,
                          'h2d_bytes_per_step': int(pcm_host.numel() * 2 + target_host.numel() * 4 + b2 * 4),
                          'd2h_bytes_per_step': 4,
                          'input': 'int16 PCM in pinned host memory (the HDF5 on-disk format); x/32767 fused on device'},
            'gpu_launches': launches,
            'clocks': clock_info,
            'roofline': roof,
            'logmel': logmel,
            'cpu_baseline': cpu,
            'kernel_shares': shares,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='batch_size per GPU (raw clips = 2x under mixup)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of the CUDA-graph replay')
    ap.add_argument('--model', default=MODEL, help='any of the seven Cnn_9layers_* classes (default: the metric\'s model)')
    ap.add_argument('--workload', default='train', choices=['train', 'logmel', 'eval'])
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world != args.gpus and world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch ourselves under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29513', os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.workload == 'logmel':
        return run_logmel(args)
    if args.workload == 'eval':
        return run_eval(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
