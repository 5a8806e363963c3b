"""Run the reference's *unmodified* ``pytorch/main.py`` on top of the B200 path.

    python -m sound_event_detection_dcase2017_task4_b200.launch /path/to/reference train \
        --dataset_dir=... --workspace=... --holdout_fold=1 --model_type=Cnn_9layers_Gru_FrameAtt \
        --loss_type=clip_bce --augmentation=mixup --learning_rate=1e-3 --batch_size=32 \
        --resume_iteration=0 --stop_iteration=50000 --cuda          (flags: runme.sh:19-21)

``main.py`` resolves its collaborators by bare module name (pytorch/main.py:3, :18-27):
``sys.path[0]`` is the script directory and ``sys.path[0]/../utils`` is inserted at position 1.
``prepare()`` arranges ``sys.path`` so that

    models, losses, pytorch_utils          -> dropin/  (this package; seam B of SURVEY 8b)
    data_generator                         -> dropin/  (memory-mapped clip store; HDF5 packs still work via h5py)
    torchlibrosa.{stft,augmentation}       -> dropin/torchlibrosa  (seam A; pytorch/models.py:10-11)
    evaluate                               -> <reference>/pytorch   (untouched, off the hot path)
    config, utilities, vad, ...            -> <reference>/utils     (untouched)

and ``main()`` then executes the reference file's bytes with ``runpy`` -- nothing under the reference tree
is edited, copied or monkey-patched.
"""
import os
import runpy
import sys

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dropin')
SHADOWED = ('models', 'losses', 'pytorch_utils', 'torchlibrosa', 'data_generator')


def prepare(reference_root):
    """Put the drop-in directory first and the reference's own directories behind it on ``sys.path``."""
    ref_pytorch = os.path.join(reference_root, 'pytorch')
    ref_utils = os.path.join(reference_root, 'utils')
    for d in (ref_pytorch, ref_utils):
        if not os.path.isdir(d):
            raise FileNotFoundError('not a checkout of the reference: %s is missing' % d)
    repo_root = os.path.dirname(os.path.dirname(DROPIN_DIR))
    if repo_root not in sys.path:
        sys.path.append(repo_root)
    for name in list(sys.modules):
        if name.split('.')[0] in SHADOWED:
            origin = getattr(sys.modules[name], '__file__', '') or ''
            if not origin.startswith(DROPIN_DIR):
                del sys.modules[name]
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') not in (DROPIN_DIR, ref_pytorch, ref_utils)]
    # main.py:3 inserts sys.path[0]/../utils at index 1; index 0 must therefore be the drop-in directory
    sys.path.insert(0, DROPIN_DIR)
    sys.path.insert(1, ref_utils)
    sys.path.insert(2, ref_pytorch)
    return os.path.join(ref_pytorch, 'main.py')


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit('usage: python -m sound_event_detection_dcase2017_task4_b200.launch <reference root> '
                         '<main.py arguments...>')
    script = prepare(argv[0])
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
