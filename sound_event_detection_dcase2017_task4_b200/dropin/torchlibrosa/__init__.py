"""Drop-in for the third-party ``torchlibrosa`` package (seam A): same class
names and constructor signatures, same state_dict keys, B200 kernels underneath."""
from . import stft, augmentation   # noqa: F401
__version__ = '0.0.4+sedb200'
