"""CPU oracle for the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package, and there only as the
checker or as the timed CPU baseline.  The product package
(``sound_event_detection_dcase2017_task4_b200``) never imports it.

What it restates (plain PyTorch-CPU fp32 / numpy, written from the published behaviour, not
vendored):

* ``oracle.frontend``  -- the un-vendored third-party ``torchlibrosa==0.0.4`` classes the
  reference imports at ``pytorch/models.py:10-11`` (``Spectrogram``, ``LogmelFilterBank``,
  ``SpecAugmentation``) plus the librosa Slaney mel filterbank they depend on.
  PARITY UNPINNED at this boundary: torchlibrosa is not in /root/reference, not installed
  and not downloadable, and the reference holds no golden vectors.  Self-consistency is
  checked against an independent float64 ``torch.stft`` + analytic mel bank
  (tests/test_oracle_frontend.py).
* ``oracle.sed``       -- the seven ``Cnn_9layers_*`` models, ``ConvBlock``, ``AttBlock``,
  ``MultiHead``, ``clip_bce``, ``do_mixup``, ``Mixup`` of ``pytorch/models.py``,
  ``pytorch/losses.py``, ``pytorch/pytorch_utils.py``, ``utils/utilities.py``.
  PINNED: validated against the unmodified reference modules imported in the authoring
  container (tests/golden/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).
* ``oracle.torchlibrosa`` -- a stand-in package so the unmodified reference ``models.py``
  can be imported (only by the golden generator / tests).
"""
