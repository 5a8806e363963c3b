"""sed-b200: B200-native (sm_100a) kernels behind the per-clip training path of
qiuqiangkong/sound_event_detection_dcase2017_task4.

Python here is host plumbing over PyTorch tensors; the arithmetic lives in
``libsedb200.so`` (hand-written CUDA, C ABI declared in include/sed_b200.h).
There is no CPU fallback: every op raises if the library or a CUDA device is missing.
"""
__version__ = '0.1.0'
