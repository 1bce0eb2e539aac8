"""Oracle for ``librosa.effects.trim`` as called by ``AudioSegment.__init__``
(``roar/collections/asr/parts/preprocessing/segment.py:76-88``; ``TTSDataset`` forwards ``trim``,
``trim_ref``, ``trim_top_db``, ``trim_frame_length``, ``trim_hop_length``, ``dataset.py:285-291,613-617``).

TEST INFRASTRUCTURE ONLY.  **parity unpinned**: librosa is not in this image; this restates librosa
0.10.x ``effects.trim`` -> ``_signal_to_frame_nonsilent`` -> ``feature.rms`` (``center=True``,
``pad_mode="constant"``) -> ``amplitude_to_db`` / ``power_to_db`` (``amin=1e-5``, ``top_db=None``).
"""
import numpy as np


def rms(y, frame_length=2048, hop_length=512):
    y = np.asarray(y, dtype=np.float32)
    pad = frame_length // 2
    yp = np.pad(y, (pad, pad), mode="constant")
    T = 1 + (len(yp) - frame_length) // hop_length
    frames = np.lib.stride_tricks.as_strided(yp, shape=(frame_length, T),
                                             strides=(yp.strides[0], hop_length * yp.strides[0]))
    power = np.mean(np.abs(frames) ** 2, axis=-2, keepdims=True)
    return np.sqrt(power)


def nonsilent_frames(y, top_db=60, ref=np.max, frame_length=2048, hop_length=512):
    mse = rms(y, frame_length, hop_length)[0]
    magnitude = np.abs(mse)
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    amin = 1e-5 ** 2
    power = np.square(magnitude)
    db = 10.0 * np.log10(np.maximum(amin, power)) - 10.0 * np.log10(np.maximum(amin, ref_value ** 2))
    return db > -top_db


def trim(y, top_db=60, ref=np.max, frame_length=2048, hop_length=512):
    """-> (y[start:end], (start, end))"""
    y = np.asarray(y)
    non_silent = nonsilent_frames(y, top_db, ref, frame_length, hop_length)
    nz = np.flatnonzero(non_silent)
    if nz.size > 0:
        start = int(nz[0] * hop_length)
        end = min(y.shape[-1], int((nz[-1] + 1) * hop_length))
    else:
        start, end = 0, 0
    return y[start:end], (start, end)
