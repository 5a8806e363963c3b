"""oracle stand-in: torchlibrosa.stft (see oracle/frontend.py)."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location(
    '_oracle_frontend', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                     'frontend.py'))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
STFT = _m.STFT
Spectrogram = _m.Spectrogram
LogmelFilterBank = _m.LogmelFilterBank
