"""Host-side pieces of /root/reference/utils/utilities.py that sit ON the training step
(SURVEY.md section 8, rows a6 / f1): the mixup coefficient generator consumed at
/root/reference/pytorch/main.py:172-173, :233-235 and the PCM scaling of the loader
(/root/reference/utils/utilities.py:66-67).

These are numpy-stream contracts, not device math: ``Mixup`` must emit exactly the reference's
``RandomState(seed).beta(alpha, alpha, 1)[0]`` sequence (first value for alpha = 1, seed = 1234:
0.23538938957272115), because the lambdas decide every mixed sample and target of the run.
"""
import numpy as np


class Mixup(object):
    def __init__(self, mixup_alpha, random_seed=1234):
        """Mixup coefficient generator (utilities.py:220-226)."""
        self.mixup_alpha = mixup_alpha
        self.random_state = np.random.RandomState(random_seed)

    def get_lambda(self, batch_size):
        """(batch_size,) float64 ``[lam_0, 1 - lam_0, lam_1, 1 - lam_1, ...]`` (utilities.py:228-242).

        One bulk draw: the legacy ``RandomState.beta`` fills its output element by element from the
        same bit stream, so ``beta(a, a, n)`` equals n successive ``beta(a, a, 1)[0]`` calls
        (pinned against the reference in tests/test_host_logic.py)."""
        pairs = (batch_size + 1) // 2
        lam = self.random_state.beta(self.mixup_alpha, self.mixup_alpha, pairs)
        out = np.empty(2 * pairs, dtype=np.float64)
        out[0::2] = lam
        out[1::2] = 1. - lam
        return out

    def fill_lambda(self, out):
        """Writes the next ``len(out)`` coefficients into a (pinned) float32 host array in place:
        the cast ``move_data_to_device`` applies at main.py:238 (float64 -> torch.Tensor float32)."""
        out[...] = self.get_lambda(out.shape[0])
        return out


def int16_to_float32(x):
    return (x / 32767.).astype(np.float32)


def float32_to_int16(x):
    assert np.max(np.abs(x)) <= 1.
    return (x * 32767.).astype(np.int16)
