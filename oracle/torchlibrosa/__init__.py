"""Stand-in for the un-vendored ``torchlibrosa`` package (TEST INFRASTRUCTURE).

Putting ``/root/repo/oracle`` on ``sys.path`` lets the unmodified reference
``pytorch/models.py`` (``from torchlibrosa.stft import ...``, models.py:10-11) import.
"""
