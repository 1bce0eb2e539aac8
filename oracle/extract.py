"""Oracle: one utterance through the reference's extraction sequence (CPU baseline + checker).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``TTSDataset.__getitem__`` (``roar/collections/tts/data/dataset.py:643-755``) including its
redundancy -- the STFT runs three times per utterance (log-mel ``:656``, the prior's ``mel_len``
``:669-670``, energy ``:751``), pyin uses the dense Viterbi librosa uses -- so that timing it is a
faithful CPU baseline ("kind": "port").
"""
import numpy as np

from . import melfb, prior, pyin, spec


def extract_utterance(audio, text_len, sr=22050, n_fft=1024, hop_length=256, win_length=1024, n_mels=80,
                      fmin=0.0, fmax=8000.0, pitch_fmin=65.40639132514966, pitch_fmax=2093.004522404789,
                      fb=None, dense_viterbi=True):
    if fb is None:
        fb = melfb.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    log_mel = spec.get_log_mel(audio, fb, n_fft, hop_length, win_length)                  # dataset.py:656
    mel_len = spec.get_log_mel(audio, fb, n_fft, hop_length, win_length).shape[2]         # dataset.py:669-670
    align_prior = prior.beta_binomial_prior_distribution(text_len, mel_len)               # dataset.py:676-678
    f0, vflag, vprob = pyin.pyin(np.asarray(audio), pitch_fmin, pitch_fmax, sr=sr,         # dataset.py:696-703
                                 frame_length=win_length, fill_na=0.0, dense_viterbi=dense_viterbi)
    energy = spec.get_energy(audio, n_fft, hop_length, win_length)                        # dataset.py:751-753
    return dict(log_mel=log_mel.numpy(), align_prior_matrix=align_prior,
                pitch=f0.astype(np.float32), voiced_mask=vflag.astype(np.float32),
                p_voiced=vprob.astype(np.float32), energy=energy.numpy())


def _worker_init():
    import torch
    torch.set_num_threads(1)


def _worker(args):
    seed, utt_id, n_samples, sr, speaker, text_len, kw = args
    from roar_b200 import synth
    y = synth.synth_utterance(seed, utt_id, n_samples, sr, speaker)
    out = extract_utterance(y, text_len, sr=sr, **kw)
    p = out["pitch"]
    return n_samples / sr, float(p[p != 0].sum()), int((p != 0).sum())


def timed_cpu_extraction(corpus="C2", first=0, count=8, processes=None, **kw):
    """Run ``count`` utterances of ``corpus`` through the reference-style CPU path on ``processes``
    worker processes (one utterance per task, like ``DataLoader(batch_size=1, num_workers=N)``,
    ``extract_sup_data.py:66-71``).  -> (audio_seconds, wall_seconds, processes)."""
    import multiprocessing as mp
    import os
    import time

    from roar_b200 import synth
    processes = processes or os.cpu_count() or 1
    spec_ = synth.CORPORA[corpus]
    man = synth.corpus_manifest(corpus, first + count)[first:first + count]
    tasks = [(spec_["seed"], u.utt_id, u.n_samples, spec_["sr"], u.speaker, u.text_len, kw) for u in man]
    # compile the numba kernels once in the parent so forked workers inherit them
    pyin.pyin(np.zeros(8192, np.float32), 65.4, 2093.0, sr=spec_["sr"], frame_length=kw.get("win_length", 1024),
              fill_na=0.0, dense_viterbi=True)
    ctx = mp.get_context("fork")
    with ctx.Pool(processes, initializer=_worker_init) as pool:
        pool.map(_worker, tasks[:min(len(tasks), processes)][:0])  # spin the workers up (no work)
        t0 = time.perf_counter()
        res = pool.map(_worker, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    return float(sum(r[0] for r in res)), wall, processes
