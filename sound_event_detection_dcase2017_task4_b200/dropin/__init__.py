"""Directory to put on ``sys.path`` (ahead of the reference's own ``pytorch/`` directory) so that the
reference's bare-name imports resolve to the B200 implementations:

    from torchlibrosa.stft import Spectrogram, LogmelFilterBank      (pytorch/models.py:10)
    from torchlibrosa.augmentation import SpecAugmentation            (pytorch/models.py:11)
    from models import *  /  from losses import get_loss_func  /  from pytorch_utils import ...
                                                                       (pytorch/main.py:21-27)
See INTEGRATION.md and sound_event_detection_dcase2017_task4_b200/launch.py.
"""
