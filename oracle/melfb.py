"""Oracle: ``librosa.filters.mel`` (librosa 0.10.x) restated in NumPy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference call sites: ``roar/collections/tts/data/dataset.py:305-314`` (slaney norm,
default) and ``roar/collections/asr/parts/preprocessing/features.py:297-308``
(``norm=mel_norm``).  librosa itself is not vendored by the reference: parity unpinned;
cross-checked against ``torchaudio.functional.melscale_fbanks`` in tests.
"""
import numpy as np

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(freq):
    """Slaney mel scale (``librosa.hz_to_mel(htk=False)``)."""
    freq = np.asanyarray(freq, dtype=np.float64)
    mels = freq / _F_SP
    if freq.ndim:
        log_t = freq >= _MIN_LOG_HZ
        mels[log_t] = _MIN_LOG_MEL + np.log(freq[log_t] / _MIN_LOG_HZ) / _LOGSTEP
    elif freq >= _MIN_LOG_HZ:
        mels = _MIN_LOG_MEL + np.log(freq / _MIN_LOG_HZ) / _LOGSTEP
    return mels


def mel_to_hz(mels):
    mels = np.asanyarray(mels, dtype=np.float64)
    freqs = _F_SP * mels
    if mels.ndim:
        log_t = mels >= _MIN_LOG_MEL
        freqs[log_t] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels[log_t] - _MIN_LOG_MEL))
    elif mels >= _MIN_LOG_MEL:
        freqs = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels - _MIN_LOG_MEL))
    return freqs


def mel_frequencies(n_mels, fmin, fmax):
    return mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels))


def mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None, norm="slaney"):
    """-> float32 ``[n_mels, 1 + n_fft//2]``; same rounding sequence as librosa
    (float32 weights filled from float64 ramps, then an in-place float32*float64 scale)."""
    if fmax is None:
        fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        raise ValueError(f"unsupported norm {norm!r}")
    return weights
