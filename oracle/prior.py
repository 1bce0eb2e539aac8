"""Oracle: beta-binomial alignment prior.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``roar/collections/tts/parts/utils/tts_dataset_utils.py:128-149``
(``logbeta``, ``logcombinations``, ``logbetabinom``, ``beta_binomial_prior_distribution``)
with the same float32 ``torch.special.gammaln`` arithmetic, and ``:69-92``
(``BetaBinomialInterpolator``).  PINNED: ``tests/golden/prior_*.npz`` hold outputs of the
reference function itself (imported by file path, ``tests/golden/make_golden.py``).
``prior_f64`` is the exact (float64 ``lgamma``) value of the same formula, used to size
the float32 noise of the reference.
"""
import functools

import numpy as np
import torch
from scipy import ndimage
from scipy.special import gammaln as _gammaln64
from torch.special import gammaln


def _logbeta(x, y):
    return gammaln(x) + gammaln(y) - gammaln(x + y)


def _logcombinations(n, k):
    return gammaln(n + 1) - gammaln(k + 1) - gammaln(n - k + 1)


def _logbetabinom(n, a, b, x):
    return _logcombinations(n, x) + _logbeta(x + a, n - x + b) - _logbeta(a, b)


def beta_binomial_prior_distribution(phoneme_count, mel_count, scaling_factor=1.0):
    """-> float32 ``[mel_count, phoneme_count]``."""
    x = torch.arange(0, phoneme_count).reshape(1, -1)
    y = torch.arange(1, mel_count + 1).reshape(-1, 1)
    a = scaling_factor * y
    b = scaling_factor * (mel_count + 1 - y)
    n = torch.FloatTensor([phoneme_count - 1])
    return _logbetabinom(n, a, b, x).exp().numpy()


def prior_f64(phoneme_count, mel_count, scaling_factor=1.0):
    """Same formula in float64 (no float32 lgamma noise)."""
    k = np.arange(0, phoneme_count, dtype=np.float64)[None, :]
    y = np.arange(1, mel_count + 1, dtype=np.float64)[:, None]
    a = scaling_factor * y
    b = scaling_factor * (mel_count + 1 - y)
    n = float(phoneme_count - 1)
    lc = _gammaln64(n + 1) - _gammaln64(k + 1) - _gammaln64(n - k + 1)
    lb1 = _gammaln64(k + a) + _gammaln64(n - k + b) - _gammaln64(n + a + b)
    lb2 = _gammaln64(a) + _gammaln64(b) - _gammaln64(a + b)
    return np.exp(lc + lb1 - lb2)


class BetaBinomialInterpolator:
    """``tts_dataset_utils.py:69-92``."""

    def __init__(self, round_mel_len_to=50, round_text_len_to=10, cache_size=500):
        self.round_mel_len_to = round_mel_len_to
        self.round_text_len_to = round_text_len_to
        self.bank = functools.lru_cache(maxsize=cache_size)(beta_binomial_prior_distribution)

    @staticmethod
    def round(val, to):
        return max(1, int(np.round((val + 1) / to))) * to

    def __call__(self, w, h):
        bw = self.round(w, to=self.round_mel_len_to)
        bh = self.round(h, to=self.round_text_len_to)
        ret = ndimage.zoom(self.bank(bw, bh).T, zoom=(w / bw, h / bh), order=1)
        assert ret.shape[0] == w, ret.shape
        assert ret.shape[1] == h, ret.shape
        return ret
