"""Drop-in mirrors of the reference's mel preprocessor modules, backed by ``roar_fbank_forward``.

* ``FilterbankFeatures``  <- ``roar/collections/asr/parts/preprocessing/features.py:196-461``
* ``AudioToMelSpectrogramPreprocessor`` <- ``.../audio_preprocessing.py:90-290`` (kwargs-only
  ``forward(input_signal=, length=)`` like the reference's ``@typecheck``-ed module)

Same constructor arguments, same outputs ``(features [B, nfilt, T'], seq_len [B])``.
``use_grads=True`` (the mel-loss preprocessors of JETS / HiFi-GAN / BigVGAN / RoarTTS,
``tts/models/jets.py:175-177``) is differentiable through ``roar_fbank_backward`` for
``normalize=None, preemph=None`` -- those models' configuration; training-time dither / narrow-band
augmentation, ``frame_splicing > 1`` and ``linear_spec=True`` raise instead of silently taking another
path.
"""
import ctypes
import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .config import FLOAT32_EPS, FLOAT32_TINY, SupConfig

CONSTANT = 1e-5


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class FilterbankFeatures(nn.Module):
    """Featurizer that converts wavs to Mel Spectrograms (CUDA, sm_100a)."""

    def __init__(self, sample_rate=16000, n_window_size=320, n_window_stride=160, window="hann",
                 normalize="per_feature", n_fft=None, preemph=0.97, nfilt=64, lowfreq=0, highfreq=None,
                 log=True, log_zero_guard_type="add", log_zero_guard_value=2 ** -24, dither=CONSTANT,
                 pad_to=16, max_duration=16.7, frame_splicing=1, exact_pad=False, pad_value=0,
                 mag_power=2.0, use_grads=False, rng=None, nb_augmentation_prob=0.0, nb_max_freq=4000,
                 mel_norm="slaney", stft_exact_pad=False, stft_conv=False):
        super().__init__()
        if exact_pad and n_window_stride % 2 == 1:
            raise NotImplementedError(
                f"{self} received exact_pad == True, but hop_size was odd. If audio_length % hop_size == 0. Then the "
                "returned spectrogram would not be of length audio_length // hop_size. Please use an even hop_size.")
        if (n_window_size is None or n_window_stride is None or not isinstance(n_window_size, int)
                or not isinstance(n_window_stride, int) or n_window_size <= 0 or n_window_stride <= 0):
            raise ValueError(f"{self} got an invalid value for either n_window_size or "
                             f"n_window_stride. Both must be positive ints.")
        if log_zero_guard_type not in ["add", "clamp"]:
            raise ValueError(f"{self} received {log_zero_guard_type} for the log_zero_guard_type parameter. "
                             f"It must be either 'add' or 'clamp'.")
        if use_grads and (preemph is not None or normalize is not None):
            raise NotImplementedError("use_grads=True is differentiated for normalize=None, preemph=None (the "
                                      "reference's mel-loss configurations: jets.py:175-177, hifigan.py:56-58)")
        if isinstance(normalize, dict) and not ("fixed_mean" in normalize and "fixed_std" in normalize):
            raise ValueError("a normalize dict needs 'fixed_mean' and 'fixed_std' (features.py:50-57)")
        if use_grads and (frame_splicing != 1 or pad_to == "max"):
            raise NotImplementedError("use_grads=True is differentiated for frame_splicing=1 and an integer pad_to")
        # Options the kernel does not fuse -- frame_splicing > 1, a fixed mean / std table, pad_to="max" -- run the
        # kernel without normalisation / padding and finish with the reference's own tensor statements
        # (features.py:434-460) on the device: they are shape plumbing, not arithmetic worth a kernel.
        self._post = frame_splicing != 1 or isinstance(normalize, dict) or pad_to == "max"
        self.win_length = n_window_size
        self.hop_length = n_window_stride
        self.n_fft = n_fft or 2 ** math.ceil(math.log2(self.win_length))
        self.stft_pad_amount = (self.n_fft - self.hop_length) // 2 if exact_pad else None
        self.normalize = normalize
        self.log = log
        self.dither = dither
        self.frame_splicing = frame_splicing
        self.nfilt = nfilt
        self.preemph = preemph
        self.pad_to = pad_to
        self.pad_value = pad_value
        self.mag_power = mag_power
        self.use_grads = use_grads
        self.nb_augmentation_prob = nb_augmentation_prob
        self.log_zero_guard_type = log_zero_guard_type
        self.log_zero_guard_value = log_zero_guard_value
        self.sample_rate = sample_rate
        guard = self.log_zero_guard_value_fn(None)
        self._cfg = SupConfig(
            sample_rate=sample_rate, n_fft=self.n_fft, win_length=self.win_length, hop_length=self.hop_length,
            window=window if window in ("hann", "hamming", "blackman", "bartlett") else "none",
            n_mels=nfilt, lowfreq=lowfreq, highfreq=highfreq or sample_rate / 2, mel_norm=mel_norm,
            spec_floor=CONSTANT if use_grads else 0.0, mag_power=mag_power, log_mode=(log_zero_guard_type if log else None),
            log_guard=guard, exact_pad=exact_pad, preemph=preemph,
            normalize=normalize if normalize in ("per_feature", "all_features") and not self._post else None,
            pad_value=pad_value, pad_to=int(pad_to) if pad_to and not self._post else 0,
            pyin=False)      # mel-only handle: no pYIN tables, none of pYIN's geometry limits
        self._c = self._cfg.to_c()
        self._lib = _lib.load()
        self._h = None
        self._ws = None
        fb = np.zeros((nfilt, self.n_fft // 2 + 1), dtype=np.float32)
        _lib.check(self._lib.roar_sup_host_mel_filterbank(ctypes.byref(self._c), fb.ctypes.data_as(ctypes.c_void_p)))
        self.register_buffer("fb", torch.from_numpy(fb).unsqueeze(0))
        win = np.zeros(self.n_fft, dtype=np.float32)
        _lib.check(self._lib.roar_sup_host_window(ctypes.byref(self._c), win.ctypes.data_as(ctypes.c_void_p)))
        left = (self.n_fft - self.win_length) // 2
        self.register_buffer("window", torch.from_numpy(win[left:left + self.win_length].copy()))
        max_length = self.get_seq_len(torch.tensor(max_duration * sample_rate, dtype=torch.float))
        max_pad = pad_to - (max_length % pad_to) if (pad_to != "max" and pad_to > 0) else 0   # features.py:341-343
        self.max_length = max_length + max_pad

    def log_zero_guard_value_fn(self, x):
        if isinstance(self.log_zero_guard_value, str):
            if self.log_zero_guard_value == "tiny":
                return FLOAT32_TINY
            if self.log_zero_guard_value == "eps":
                return FLOAT32_EPS
            raise ValueError(f"{self} received {self.log_zero_guard_value} for the log_zero_guard_type parameter. "
                             f"It must be either a number, 'tiny', or 'eps'")
        return self.log_zero_guard_value

    def get_seq_len(self, seq_len):
        pad_amount = self.stft_pad_amount * 2 if self.stft_pad_amount is not None else self.n_fft // 2 * 2
        seq_len = torch.floor_divide((seq_len + pad_amount - self.n_fft), self.hop_length) + 1
        return seq_len.to(dtype=torch.long)

    @property
    def filter_banks(self):
        return self.fb

    def _handle(self, device):
        if self._h is None:
            h = ctypes.c_void_p()
            _lib.check(self._lib.roar_sup_create(ctypes.byref(self._c), device.index or 0, ctypes.byref(h)))
            self._h = h
            self._dev = device
        elif self._dev != device:
            raise _lib.RoarSupError("FilterbankFeatures handle is bound to one device")
        return self._h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.roar_sup_destroy(self._h)
        except Exception:
            pass

    def _run_forward(self, x, lens):
        B, Lmax = x.shape
        h = self._handle(x.device)
        Tpad = int(self._lib.roar_fbank_out_frames(h, Lmax))
        out = torch.empty(B, self.nfilt, Tpad, dtype=torch.float32, device=x.device)
        out_len = torch.empty(B, dtype=torch.int64, device=x.device)
        ws = self._workspace(h, B, x.device)
        _lib.check(self._lib.roar_fbank_forward(
            h, _ptr(x), _ptr(lens), B, Lmax, _ptr(out), _ptr(out_len), _ptr(ws), ws.numel(),
            ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return out, out_len

    def _workspace(self, h, B, device):
        need = int(self._lib.roar_fbank_workspace_bytes(h, B))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _run_backward(self, x, lens, grad_out):
        B, Lmax = x.shape
        h = self._handle(x.device)
        gx = torch.empty_like(x)
        ws = self._workspace(h, B, x.device)
        _lib.check(self._lib.roar_fbank_backward(
            h, _ptr(x), _ptr(lens), B, Lmax, _ptr(grad_out.contiguous().float()), _ptr(gx), _ptr(ws),
            ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return gx

    def forward(self, x, seq_len, linear_spec=False):
        if linear_spec:
            raise NotImplementedError("linear_spec=True is outside the accelerated path")
        if self.training and (self.dither > 0 or self.nb_augmentation_prob > 0.0):
            raise NotImplementedError("training-time dither / narrow-band augmentation are outside the accelerated "
                                      "path: call .eval() or construct with dither=0.0")
        if not x.is_cuda:
            raise _lib.RoarSupError("roar_b200.FilterbankFeatures needs CUDA tensors (no CPU fallback)")
        lens = seq_len.to(device=x.device, dtype=torch.int64).contiguous()
        if self.normalize == "per_feature" and bool((self.get_seq_len(lens) == 1).any()):
            raise ValueError(
                "normalize_batch with `per_feature` normalize_type received a tensor of length 1. This will result "
                "in torch.std() returning nan. Make sure your audio length has enough samples for a single "
                "feature (ex. at least `hop_length` for Mel Spectrograms).")
        if self.use_grads and torch.is_grad_enabled() and x.requires_grad:
            out, out_len = _FbankFunction.apply(x, lens, self)
            return out, out_len
        with torch.no_grad():
            out, out_len = self._run_forward(x.contiguous().float(), lens)
            return (self._finish(out, out_len), out_len) if self._post else (out, out_len)

    def _finish(self, x, seq_len):
        """features.py:434-460 for the options the kernel leaves out.  ``x`` is the kernel's log-mel: valid frames
        un-normalised, frames beyond ``seq_len`` already ``pad_value``, no padding of the time axis."""
        if self.frame_splicing > 1:
            # splice_frames (features.py:83-95): cat([x[:, :, :n], x[:, :, n:]], dim=2) is x itself, so the
            # reference stacks `frame_splicing` copies of x along the feature axis
            x = torch.cat([x] * self.frame_splicing, dim=1)
        valid = (torch.arange(x.shape[2], device=x.device)[None, :] < seq_len[:, None]).unsqueeze(1)   # [B, 1, T]
        if isinstance(self.normalize, dict):
            mean = torch.tensor(self.normalize["fixed_mean"], device=x.device, dtype=x.dtype).view(x.shape[0], x.shape[1], 1)
            std = torch.tensor(self.normalize["fixed_std"], device=x.device, dtype=x.dtype).view(x.shape[0], x.shape[1], 1)
            x = (x - mean) / std
        elif self.normalize in ("per_feature", "all_features"):
            dims = (2,) if self.normalize == "per_feature" else (1, 2)
            n = valid.sum(dim=dims, keepdim=True).to(x.dtype) * (1 if self.normalize == "per_feature" else x.shape[1])
            xm = torch.where(valid, x, torch.zeros_like(x))
            mean = xm.sum(dim=dims, keepdim=True) / n
            var = (torch.where(valid, x - mean, torch.zeros_like(x)) ** 2).sum(dim=dims, keepdim=True) / (n - 1)
            x = (x - mean) / (torch.sqrt(var) + CONSTANT)
        x = x.masked_fill(~valid, self.pad_value)
        if self.pad_to == "max":
            x = nn.functional.pad(x, (0, int(self.max_length) - x.size(-1)), value=self.pad_value)
        elif self.pad_to and self.pad_to > 0:
            pad_amt = x.size(-1) % self.pad_to
            if pad_amt != 0:
                x = nn.functional.pad(x, (0, self.pad_to - pad_amt), value=self.pad_value)
        return x


class _FbankFunction(torch.autograd.Function):
    """autograd through the CUDA preprocessor: forward = roar_fbank_forward, backward = roar_fbank_backward
    (the forward spectrum is recomputed in the backward kernel, nothing but the audio is saved)."""

    @staticmethod
    def forward(ctx, x, lens, module):
        xc = x.contiguous().float()
        out, out_len = module._run_forward(xc, lens)
        ctx.save_for_backward(xc, lens)
        ctx.module = module
        ctx.mark_non_differentiable(out_len)
        return out, out_len

    @staticmethod
    def backward(ctx, grad_out, _grad_len):
        xc, lens = ctx.saved_tensors
        return ctx.module._run_backward(xc, lens, grad_out), None, None


class AudioToMelSpectrogramPreprocessor(nn.Module):
    """Same arguments as the reference module; ``forward`` is keyword-only like its ``@typecheck``."""

    def __init__(self, sample_rate=16000, window_size=0.02, window_stride=0.01, n_window_size=None,
                 n_window_stride=None, window="hann", normalize="per_feature", n_fft=None, preemph=0.97,
                 features=64, lowfreq=0, highfreq=None, log=True, log_zero_guard_type="add",
                 log_zero_guard_value=2 ** -24, dither=1e-5, pad_to=16, frame_splicing=1, exact_pad=False,
                 pad_value=0, mag_power=2.0, rng=None, nb_augmentation_prob=0.0, nb_max_freq=4000,
                 use_torchaudio: bool = False, mel_norm="slaney", stft_exact_pad=False, stft_conv=False):
        super().__init__()
        self._sample_rate = sample_rate
        if window_size and n_window_size:
            raise ValueError(f"{self} received both window_size and n_window_size. Only one should be specified.")
        if window_stride and n_window_stride:
            raise ValueError(f"{self} received both window_stride and n_window_stride. Only one should be specified.")
        if window_size:
            n_window_size = int(window_size * self._sample_rate)
        if window_stride:
            n_window_stride = int(window_stride * self._sample_rate)
        if use_torchaudio:
            raise NotImplementedError("use_torchaudio=True selects another backend in the reference; not accelerated")
        self.win_length = n_window_size
        self.hop_length = n_window_stride
        self.featurizer = FilterbankFeatures(
            sample_rate=self._sample_rate, n_window_size=n_window_size, n_window_stride=n_window_stride,
            window=window, normalize=normalize, n_fft=n_fft, preemph=preemph, nfilt=features, lowfreq=lowfreq,
            highfreq=highfreq, log=log, log_zero_guard_type=log_zero_guard_type,
            log_zero_guard_value=log_zero_guard_value, dither=dither, pad_to=pad_to, frame_splicing=frame_splicing,
            exact_pad=exact_pad, pad_value=pad_value, mag_power=mag_power, rng=rng,
            nb_augmentation_prob=nb_augmentation_prob, nb_max_freq=nb_max_freq, mel_norm=mel_norm,
            stft_exact_pad=stft_exact_pad, stft_conv=stft_conv)

    def forward(self, *, input_signal, length):
        return self.get_features(input_signal, length)

    def get_features(self, input_signal, length):
        return self.featurizer(input_signal, length)

    @property
    def filter_banks(self):
        return self.featurizer.filter_banks
