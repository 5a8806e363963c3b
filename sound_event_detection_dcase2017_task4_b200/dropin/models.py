"""seam B shim: `import models` resolves here when dropin/ is on sys.path (pytorch/main.py:21-27)."""
from sound_event_detection_dcase2017_task4_b200.models import *  # noqa: F401,F403
