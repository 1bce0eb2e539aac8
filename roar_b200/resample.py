"""Sample-rate conversion on the GPU.

The reference resamples inside ``AudioSegment.__init__`` (``asr/parts/preprocessing/segment.py:68-75``) with
``librosa.core.resample(samples, orig_sr=, target_sr=)``, whose default ``res_type`` is the third-party
``soxr_hq`` resampler (not in ``/root/reference``, not in this image).  Here the conversion is a polyphase FIR
with the arithmetic of ``scipy.signal.resample_poly`` -- rational factor ``up / down``, Kaiser(5.0)-windowed sinc
low-pass at the lower Nyquist frequency, ``10 * max(up, down)`` taps per side -- run as one kernel over a packed
batch (``roar_sup_resample``).  Same band limit as soxr_hq, different filter: outputs agree with
``scipy.signal.resample_poly`` to float32 rounding, and with the reference only to the extent two good low-pass
designs agree (**parity unpinned**, stated in DESIGN.md).
"""
import ctypes
import math
from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

from . import _lib


@lru_cache(maxsize=32)
def plan(orig_sr: int, target_sr: int) -> Tuple[int, int, np.ndarray, int, int]:
    """-> (up, down, taps float32 [2 * half_len + 1] (already * up), n_pre_pad, n_pre_remove) exactly as
    ``scipy.signal.resample_poly(x, up, down)`` lays its filter out for float32 input."""
    from scipy.signal import firwin
    g = math.gcd(int(orig_sr), int(target_sr))
    up, down = int(target_sr) // g, int(orig_sr) // g
    if up == down == 1:
        return 1, 1, np.ones(1, np.float32), 0, 0
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(np.float32)
    h *= up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    return up, down, h, n_pre_pad, n_pre_remove


def out_len(n_in, up: int, down: int):
    """``ceil(n_in * up / down)`` (resample_poly's output length)."""
    n = np.asarray(n_in, dtype=np.int64) * up
    return n // down + (n % down != 0)


def resample_batch(ex, batch, orig_sr, target_sr: int):
    """Packed batch -> packed batch at ``target_sr`` (new device buffer, 16-byte aligned starts).  ``orig_sr`` is
    one rate for the whole batch or an array with one rate per utterance (a mixed corpus): one kernel launch per
    distinct source rate, utterances already at ``target_sr`` are copied through."""
    from .extractor import PackedBatch, pack_layout
    n = batch.n_utts
    srs = np.full(n, int(orig_sr), dtype=np.int64) if np.ndim(orig_sr) == 0 else np.asarray(orig_sr, dtype=np.int64)
    if n == 0 or bool((srs == int(target_sr)).all()):
        return batch
    lens = np.empty(n, dtype=np.int64)
    for sr in np.unique(srs):
        up, down = plan(int(sr), int(target_sr))[:2]
        m = srs == sr
        lens[m] = out_len(batch.lens_host[m], up, down)
    offs, total = pack_layout(lens)
    out = torch.zeros(total, dtype=torch.float32, device=ex.device)
    d_oo = ex._up(offs, np.int64)
    stream = ctypes.c_void_p(torch.cuda.current_stream(ex.device).cuda_stream)
    for sr in np.unique(srs):
        up, down, taps, n_pre_pad, n_pre_remove = plan(int(sr), int(target_sr))
        part = np.where(srs == sr, lens, 0)                  # other rates: nothing to do in this launch
        d_taps, d_ol = ex._up(taps, np.float32), ex._up(part, np.int32)
        _lib.check(ex.lib.roar_sup_resample(
            ex._h, ctypes.c_void_p(batch.audio.data_ptr()), ctypes.c_void_p(batch.sample_off.data_ptr()),
            ctypes.c_void_p(batch.sample_len.data_ptr()), n, int(part.max()), up, down,
            ctypes.c_void_p(d_taps.data_ptr()), len(taps), n_pre_pad, n_pre_remove, ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(d_oo.data_ptr()), ctypes.c_void_p(d_ol.data_ptr()), stream))
        ex.kernel_launches += 1
    return PackedBatch(out, d_oo, ex._up(lens, np.int32), lens, offs)
