"""Native host-side I/O of the extraction run (``libroar_sup.so``, C++ threads, GIL released by ctypes).

* wav decoding into caller-owned (pinned) staging buffers -- the ``AudioSegment.from_file`` step
  (``asr/parts/preprocessing/segment.py:156-278``) for RIFF/WAVE input: 16-bit mono PCM stays int16 and is
  converted on the GPU (``roar_sup_pcm16_to_f32``), every other encoding is decoded to float32 mono here;
* the ``.pt`` cache writer -- ``torch.save(tensor, path)`` of ``dataset.py:656-657, 704-708, 752-753`` as a
  stored zip archive built in one buffer per file, written by a thread pool, temp name + rename.
"""
import ctypes
import os
from typing import List, Sequence

import numpy as np
import torch

from . import _lib


class WavInfo(ctypes.Structure):
    """``roar_wav_info`` (include/roar_sup.h)."""
    _fields_ = [("sample_rate", ctypes.c_int32), ("channels", ctypes.c_int32), ("bits", ctypes.c_int32),
                ("format", ctypes.c_int32), ("n_frames", ctypes.c_int64), ("data_offset", ctypes.c_int64)]


WAV_INFO_DTYPE = np.dtype([("sample_rate", np.int32), ("channels", np.int32), ("bits", np.int32),
                           ("format", np.int32), ("n_frames", np.int64), ("data_offset", np.int64)])
assert WAV_INFO_DTYPE.itemsize == ctypes.sizeof(WavInfo)


def default_threads() -> int:
    return max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))


def _c_paths(paths: Sequence[str]):
    arr = (ctypes.c_char_p * len(paths))()
    arr[:] = [os.fsencode(p) for p in paths]
    return arr


def wav_probe(paths: Sequence[str], threads: int = 0) -> np.ndarray:
    """-> structured array (``WAV_INFO_DTYPE``) per file; raises naming the first undecodable file."""
    lib = _lib.load()
    info = np.zeros(len(paths), dtype=WAV_INFO_DTYPE)
    if len(paths) == 0:
        return info
    bad = lib.roar_sup_wav_probe_batch(_c_paths(paths), len(paths), info.ctypes.data_as(ctypes.c_void_p),
                                       threads or default_threads())
    if bad:
        raise _lib.RoarSupError(f"{bad} audio file(s) could not be decoded: {lib.roar_sup_last_error().decode()}")
    return info


def is_pcm16_mono(info: np.ndarray) -> np.ndarray:
    return (info["format"] == 1) & (info["bits"] == 16) & (info["channels"] == 1)


def wav_read(paths: Sequence[str], info: np.ndarray, first_frame: np.ndarray, n_frames: np.ndarray,
             dst: torch.Tensor, dst_off: np.ndarray, channel: int = -1, threads: int = 0) -> None:
    """Decode frame ranges of ``paths`` into ``dst`` (CPU tensor, int16 -> raw PCM, float32 -> decoded mono)
    at element offsets ``dst_off``."""
    lib = _lib.load()
    assert not dst.is_cuda and dst.is_contiguous() and dst.dtype in (torch.int16, torch.float32)
    n = len(paths)
    if n == 0:
        return
    ff = np.ascontiguousarray(first_frame, dtype=np.int64)
    nf = np.ascontiguousarray(n_frames, dtype=np.int64)
    do = np.ascontiguousarray(dst_off, dtype=np.int64)
    assert int((do + nf).max()) <= dst.numel()
    info = np.ascontiguousarray(info)
    bad = lib.roar_sup_wav_read_batch(_c_paths(paths), info.ctypes.data_as(ctypes.c_void_p), n,
                                      ff.ctypes.data_as(ctypes.c_void_p), nf.ctypes.data_as(ctypes.c_void_p),
                                      int(channel), 1 if dst.dtype == torch.int16 else 0,
                                      ctypes.c_void_p(dst.data_ptr()), do.ctypes.data_as(ctypes.c_void_p),
                                      threads or default_threads())
    if bad:
        raise _lib.RoarSupError(f"{bad} audio file(s) could not be read: {lib.roar_sup_last_error().decode()}")


def load_wav(path: str, offset: float = 0.0, duration: float = 0.0, channel: int = -1):
    """One file -> (float32 mono numpy array, sample_rate); ``offset`` / ``duration`` in seconds like
    ``AudioSegment.from_file`` (``segment.py:218-224``)."""
    info = wav_probe([path], 1)
    sr = int(info["sample_rate"][0])
    first = int(offset * sr) if offset > 0 else 0
    first = min(first, int(info["n_frames"][0]))
    count = int(info["n_frames"][0]) - first
    if duration > 0:
        count = min(count, int(duration * sr))
    out = torch.empty(count, dtype=torch.float32)
    wav_read([path], info, np.array([first]), np.array([count]), out, np.array([0]), channel, 1)
    return out.numpy(), sr


def pt_write_batch(base: torch.Tensor, elem_off: np.ndarray, shapes: List[Sequence[int]], paths: Sequence[str],
                   threads: int = 0) -> None:
    """Write ``len(paths)`` ``.pt`` files: file i = float32 tensor of ``shapes[i]`` (rank 1..3) starting at
    element ``elem_off[i]`` of the flat CPU tensor ``base``."""
    lib = _lib.load()
    assert not base.is_cuda and base.dtype == torch.float32 and base.is_contiguous()
    n = len(paths)
    if n == 0:
        return
    rank = np.array([len(s) for s in shapes], dtype=np.int32)
    shape3 = np.zeros((n, 3), dtype=np.int64)
    for i, s in enumerate(shapes):
        shape3[i, :len(s)] = s
    eo = np.ascontiguousarray(elem_off, dtype=np.int64)
    bad = lib.roar_sup_pt_write_batch(ctypes.c_void_p(base.data_ptr()), n, eo.ctypes.data_as(ctypes.c_void_p),
                                      rank.ctypes.data_as(ctypes.c_void_p), shape3.ctypes.data_as(ctypes.c_void_p),
                                      _c_paths(paths), threads or default_threads())
    if bad:
        raise _lib.RoarSupError(f"{bad} cache file(s) could not be written: {lib.roar_sup_last_error().decode()}")
