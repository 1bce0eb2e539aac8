"""Oracle: sample-rate conversion.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference resamples with ``librosa.core.resample(samples, orig_sr=, target_sr=)``
(``roar/collections/asr/parts/preprocessing/segment.py:68-75``), default ``res_type="soxr_hq"`` -- the soxr
library, a third-party dependency that is neither under ``/root/reference`` nor in this image, and whose filter
design is not restated here.  **Parity unpinned**: the oracle for the GPU resampler is
``scipy.signal.resample_poly`` (polyphase Kaiser-windowed sinc at the same band limit), i.e. it pins the KERNEL's
arithmetic, not its agreement with soxr.  ``band_limited_check`` states what the two designs must share: a tone
well inside the pass band keeps its amplitude and frequency.
"""
import math

import numpy as np


def resample(y, orig_sr: int, target_sr: int):
    from scipy.signal import resample_poly
    g = math.gcd(int(orig_sr), int(target_sr))
    return resample_poly(np.asarray(y, dtype=np.float32), int(target_sr) // g, int(orig_sr) // g)


def band_limited_check(y_out, target_sr: int, f_tone: float, amp: float, tol=0.01) -> bool:
    """A pure tone of frequency ``f_tone`` (<< Nyquist) and amplitude ``amp`` survives with both intact."""
    n = len(y_out)
    seg = np.asarray(y_out[n // 4: 3 * n // 4], dtype=np.float64)
    spec = np.abs(np.fft.rfft(seg * np.hanning(len(seg))))
    f_peak = np.argmax(spec) * target_sr / len(seg)
    a = 2 * spec.max() / np.hanning(len(seg)).sum()
    return abs(f_peak - f_tone) <= target_sr / len(seg) * 1.5 and abs(a / amp - 1) <= tol
