"""seam B shim: `import losses` resolves here when dropin/ is on sys.path (pytorch/main.py:21-27)."""
from sound_event_detection_dcase2017_task4_b200.losses import *  # noqa: F401,F403
