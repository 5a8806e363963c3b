"""seam shim: `import data_generator` resolves here when dropin/ is on sys.path (pytorch/main.py:25-26)."""
from sound_event_detection_dcase2017_task4_b200.data_generator import *  # noqa: F401,F403
from sound_event_detection_dcase2017_task4_b200.data_generator import (  # noqa: F401
    DCASE2017Task4Dataset, TrainSampler, TestSampler, collate_fn)
