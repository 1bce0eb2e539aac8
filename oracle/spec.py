"""Oracle: ``TTSDataset.get_spec`` / ``get_log_mel`` / energy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``roar/collections/tts/data/dataset.py``:
  * ``self.stft`` lambda ``:324-333`` -- ``torch.stft(center=True (default), reflect pad,
    window=hann_window(win_length, periodic=False), return_complex=True)``
  * ``get_spec`` ``:524-530`` -- ``sqrt(re^2 + im^2 + 1e-9)``
  * ``get_log_mel`` ``:532-537`` -- ``log(clamp(fb @ spec, min=finfo(float32).tiny))``
  * energy ``:751-753`` -- ``torch.linalg.norm(spec, axis=0)``
The torch calls are the ones the reference itself makes (torch is its dependency).
``spec_f64`` is an independent float64 framing+rFFT used to size the fp32 error budget.
"""
import numpy as np
import torch

from .melfb import mel_filterbank

EPSILON = 1e-9  # dataset.py:60-67


def stft_complex(audio, n_fft, hop_length, win_length):
    x = torch.as_tensor(np.asarray(audio, dtype=np.float32))
    window = torch.hann_window(win_length, periodic=False).to(torch.float)
    return torch.stft(
        input=x,
        n_fft=n_fft,
        hop_length=hop_length,
        win_length=win_length,
        window=window,
        return_complex=True,
    )


def get_spec(audio, n_fft=1024, hop_length=256, win_length=1024):
    """-> torch float32 ``[1 + n_fft//2, T]``, ``T = 1 + L // hop``."""
    spec = torch.view_as_real(stft_complex(audio, n_fft, hop_length, win_length))
    return torch.sqrt(spec.pow(2).sum(-1) + EPSILON)


def get_log_mel(audio, fb, n_fft=1024, hop_length=256, win_length=1024):
    """``fb``: float32 ``[n_mels, F]`` -> torch float32 ``[1, n_mels, T]`` (the reference's
    ``fb`` carries a leading unit dim, ``dataset.py:305-314``)."""
    spec = get_spec(audio, n_fft, hop_length, win_length)
    fbt = torch.as_tensor(fb, dtype=torch.float).unsqueeze(0)
    mel = torch.matmul(fbt.to(spec.dtype), spec)
    return torch.log(torch.clamp(mel, min=torch.finfo(mel.dtype).tiny))


def get_energy(audio, n_fft=1024, hop_length=256, win_length=1024):
    """-> torch float32 ``[T]`` (L2 norm over all linear bins, floor included)."""
    spec = get_spec(audio, n_fft, hop_length, win_length)
    return torch.linalg.norm(spec.squeeze(0), axis=0).float()


def log_mel_energy(audio, sr=22050, n_fft=1024, hop_length=256, win_length=1024,
                   n_mels=80, fmin=0.0, fmax=8000.0, fb=None):
    if fb is None:
        fb = mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    return (get_log_mel(audio, fb, n_fft, hop_length, win_length).numpy(),
            get_energy(audio, n_fft, hop_length, win_length).numpy())


def spec_f64(audio, n_fft=1024, hop_length=256, win_length=1024):
    """Independent float64 STFT magnitude (manual reflect pad + framing + rfft)."""
    y = np.asarray(audio, dtype=np.float64)
    pad = n_fft // 2
    yp = np.pad(y, (pad, pad), mode="reflect")
    n = np.arange(win_length, dtype=np.float64)
    w = 0.5 * (1.0 - np.cos(2.0 * np.pi * n / (win_length - 1)))
    wfull = np.zeros(n_fft)
    left = (n_fft - win_length) // 2
    wfull[left : left + win_length] = w
    T = 1 + len(y) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(T)[:, None]
    X = np.fft.rfft(yp[idx] * wfull[None, :], axis=1)
    return np.sqrt(X.real ** 2 + X.imag ** 2 + EPSILON).T
