"""Oracle: process pool running the CPU restatement over many synthetic utterances.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): used by ``tests/`` (parity at BASELINE.json's
configuration sizes) and ``scripts/``.  One utterance per task, the waveform regenerated inside the worker from
``(seed, utt_id, n_samples, sr, speaker)`` so that nothing but small tuples crosses the process boundary --
the same way ``DataLoader(batch_size=1, num_workers=N)`` drives ``TTSDataset.__getitem__`` in the reference
(``scripts/dataset_processing/tts/extract_sup_data.py:66-71``).  The banded Viterbi (bit-identical to the dense
one, ``tests/test_oracle.py``) keeps a few hundred utterances within seconds.
"""
import multiprocessing as mp
import os

import numpy as np

FMIN, FMAX = 65.40639132514966, 2093.004522404789


def quantize_pcm16(y: np.ndarray) -> np.ndarray:
    """float waveform -> the int16 samples a 16-bit wav file of it holds."""
    return np.clip(np.rint(np.asarray(y, dtype=np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def synth(task):
    """task = dict(seed, utt_id, n_samples, sr, speaker, pcm16) -> float32 waveform (``pcm16``: the exact
    ``x / 2**15`` values of its 16-bit quantisation, ``segment.py:140-153``)."""
    from roar_b200 import synth as S
    y = S.synth_utterance(task["seed"], task["utt_id"], task["n_samples"], task["sr"], task["speaker"])
    if task.get("pcm16"):
        y = quantize_pcm16(y).astype(np.float32) / np.float32(32768.0)
    return y


def _init():
    import torch
    torch.set_num_threads(1)


def _work(task):
    from . import extract as oextract
    from . import pyin as opyin
    from . import spec as ospec
    y = synth(task)
    cfg = task.get("cfg", {})
    sr = task["sr"]
    n_fft, hop, win = cfg.get("n_fft", 1024), cfg.get("hop_length", 256), cfg.get("win_length", 1024)
    if task.get("what", "all") == "all":
        return oextract.extract_utterance(y, task["text_len"], sr=sr, n_fft=n_fft, hop_length=hop, win_length=win,
                                          n_mels=cfg.get("n_mels", 80), fmin=cfg.get("fmin", 0.0),
                                          fmax=cfg.get("fmax", 8000.0), dense_viterbi=False)
    f0, vf, vp = opyin.pyin(y, FMIN, FMAX, sr=sr, frame_length=win, fill_na=0.0)
    out = dict(pitch=f0.astype(np.float32), voiced_mask=vf.astype(np.float32), p_voiced=vp.astype(np.float32))
    if task.get("what") == "pyin+logmel":
        out["log_mel"] = ospec.get_log_mel(y, task["fb"], n_fft, hop, win).numpy()
    return out


def run(tasks, processes=None):
    """-> list of oracle outputs, one per task, in order."""
    from . import pyin as opyin
    processes = min(processes or os.cpu_count() or 1, max(1, len(tasks)))
    # compile the numba kernels once in the parent so forked workers inherit them
    opyin.pyin(np.zeros(4096, np.float32), FMIN, FMAX, sr=22050, frame_length=1024, fill_na=0.0)
    if processes == 1:
        return [_work(t) for t in tasks]
    with mp.get_context("fork").Pool(processes, initializer=_init) as pool:
        return pool.map(_work, tasks, chunksize=1)


def tasks_for(corpus: str, manifest, what="all", cfg=None, pcm16=False, fb=None):
    from roar_b200 import synth as S
    spec_ = S.CORPORA[corpus]
    out = []
    for u in manifest:
        t = dict(seed=spec_["seed"], utt_id=u.utt_id, n_samples=u.n_samples, sr=spec_["sr"], speaker=u.speaker,
                 text_len=u.text_len, what=what, cfg=cfg or {}, pcm16=pcm16)
        if fb is not None:
            t["fb"] = fb
        out.append(t)
    return out
