"""oracle stand-in: torchlibrosa.augmentation (see oracle/frontend.py)."""
from .stft import _m

DropStripes = _m.DropStripes
SpecAugmentation = _m.SpecAugmentation
