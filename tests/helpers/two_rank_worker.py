"""Worker of the world-size-2 tests (tests/test_gpu_reference_loop.py, tests/test_host_logic.py): one process per
rank, gloo rendezvous on 127.0.0.1.  mode 'cpu': only the gradient exchange + sharding helpers (no CUDA).  mode
'gpu': both ranks share cuda:0 and run one FusedTrainer step on their shard of a common synthetic batch.  modes
'gpu4' / 'gpu4_graph': four steps on the same device buffers, eager / with the CUDA-graph replay of forward + backward
(the all-reduce and Adam stay eager at world size > 1)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    mode, out_dir = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sound_event_detection_dcase2017_task4_b200 import trainer as tr
    if mode == 'cpu':
        n = 1000
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(n, generator=g, dtype=torch.float32)           # this rank's gradient of ITS mean loss
        flat = local / world                                               # what the step does: loss scaled by 1/world
        tr.exchange_gradients(flat, world)
        lo, hi = tr.shard_bounds(64, world, rank)
        torch.save({'local': local, 'reduced': flat, 'bounds': (lo, hi)}, os.path.join(out_dir, 'rank%d.pt' % rank))
    else:
        from oracle import sed
        from sound_event_detection_dcase2017_task4_b200 import models
        name = 'Cnn_9layers_Gru_FrameAtt'
        torch.cuda.set_device(0)
        torch.manual_seed(0)
        model = getattr(models, name)(32000, 1024, 320, 64, 50, 14000, 17).cuda().train()
        for m in (model.spec_augmenter.time_dropper, model.spec_augmenter.freq_dropper):
            m.drop_width = 1                                               # zero-width stripes: ranks draw independently
        trainer = tr.FusedTrainer(model, lr=1e-3, world_size=world, use_graph=(mode == 'gpu4_graph'))
        _, wave, target = sed.synthetic_batch(16, 32000, seed=77)
        lam = sed.MixupLambda(1., 1234).get_lambda(16).astype(np.float32)
        lo, hi = tr.shard_bounds(16, world, rank)
        dev_in = (torch.from_numpy(wave[lo:hi]).cuda(), torch.from_numpy(target[lo:hi]).cuda(),
                  torch.from_numpy(lam[lo:hi]).cuda())
        losses = []
        for _ in range(1 if mode == 'gpu' else 4):
            loss = trainer.step(*dev_in)
            losses.append(float(loss))                                     # read before the next replay overwrites it
        torch.cuda.synchronize()
        torch.save({'loss': losses[0], 'losses': losses, 'grad': trainer.flat_grad.cpu(),
                    'param': trainer.flat_param.cpu(), 'bounds': (lo, hi), 'graphs': len(trainer._graphs),
                    'steps': trainer.step_count}, os.path.join(out_dir, 'rank%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
