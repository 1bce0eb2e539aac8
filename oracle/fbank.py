"""Oracle: ``FilterbankFeatures.forward`` / ``normalize_batch`` (eval mode, no grads).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``roar/collections/asr/parts/preprocessing/features.py``:
``normalize_batch`` ``:25-61``, ``__init__`` ``:196-345``, ``get_seq_len`` ``:368-378``,
``forward`` ``:384-461`` using the same torch calls (eval mode: no dither, no
narrow-band augmentation; ``frame_splicing == 1``).  Defaults are those of
``AudioToMelSpectrogramPreprocessor`` (``audio_preprocessing.py:193-223``).
"""
import math

import numpy as np
import torch

from .melfb import mel_filterbank

CONSTANT = 1e-5


def normalize_batch(x, seq_len, normalize_type):
    if normalize_type == "per_feature":
        x_mean = torch.zeros((seq_len.shape[0], x.shape[1]), dtype=x.dtype)
        x_std = torch.zeros((seq_len.shape[0], x.shape[1]), dtype=x.dtype)
        for i in range(x.shape[0]):
            x_mean[i, :] = x[i, :, : seq_len[i]].mean(dim=1)
            x_std[i, :] = x[i, :, : seq_len[i]].std(dim=1)
        x_std += CONSTANT
        return (x - x_mean.unsqueeze(2)) / x_std.unsqueeze(2)
    if normalize_type == "all_features":
        x_mean = torch.zeros(seq_len.shape, dtype=x.dtype)
        x_std = torch.zeros(seq_len.shape, dtype=x.dtype)
        for i in range(x.shape[0]):
            x_mean[i] = x[i, :, : seq_len[i].item()].mean()
            x_std[i] = x[i, :, : seq_len[i].item()].std()
        x_std += CONSTANT
        return (x - x_mean.view(-1, 1, 1)) / x_std.view(-1, 1, 1)
    if isinstance(normalize_type, dict) and "fixed_mean" in normalize_type and "fixed_std" in normalize_type:
        x_mean = torch.tensor(normalize_type["fixed_mean"], dtype=x.dtype)
        x_std = torch.tensor(normalize_type["fixed_std"], dtype=x.dtype)
        return (x - x_mean.view(x.shape[0], x.shape[1]).unsqueeze(2)) / x_std.view(x.shape[0], x.shape[1]).unsqueeze(2)
    return x


def splice_frames(x, frame_splicing):
    """features.py:83-95, verbatim semantics (the cat of x[:, :, :n] and x[:, :, n:] is x itself)."""
    seq = [x]
    for n in range(1, frame_splicing):
        seq.append(torch.cat([x[:, :, :n], x[:, :, n:]], dim=2))
    return torch.cat(seq, dim=1)


class FilterbankFeaturesOracle:
    def __init__(self, sample_rate=16000, n_window_size=320, n_window_stride=160, window="hann",
                 normalize="per_feature", n_fft=None, preemph=0.97, nfilt=64, lowfreq=0,
                 highfreq=None, log=True, log_zero_guard_type="add", log_zero_guard_value=2 ** -24,
                 pad_to=16, exact_pad=False, pad_value=0, mag_power=2.0, use_grads=False,
                 mel_norm="slaney", frame_splicing=1, max_duration=16.7):
        self.win_length = n_window_size
        self.hop_length = n_window_stride
        self.n_fft = n_fft or 2 ** math.ceil(math.log2(self.win_length))
        self.stft_pad_amount = (self.n_fft - self.hop_length) // 2 if exact_pad else None
        fns = {"hann": torch.hann_window, "hamming": torch.hamming_window,
               "blackman": torch.blackman_window, "bartlett": torch.bartlett_window, "none": None}
        fn = fns.get(window, None)
        self.window = fn(self.win_length, periodic=False) if fn else None
        self.exact_pad = exact_pad
        self.frame_splicing = frame_splicing
        self.normalize = normalize
        self.log = log
        self.preemph = preemph
        self.pad_to = pad_to
        self.pad_value = pad_value
        self.mag_power = mag_power
        self.use_grads = use_grads
        self.log_zero_guard_type = log_zero_guard_type
        self.log_zero_guard_value = log_zero_guard_value
        max_length = self.get_seq_len(torch.tensor(max_duration * sample_rate, dtype=torch.float))
        max_pad = pad_to - (max_length % pad_to) if (pad_to != "max" and pad_to > 0) else 0
        self.max_length = max_length + max_pad
        highfreq = highfreq or sample_rate / 2
        self.fb = torch.tensor(
            mel_filterbank(sample_rate, self.n_fft, nfilt, lowfreq, highfreq, norm=mel_norm),
            dtype=torch.float).unsqueeze(0)

    def _guard(self, x):
        v = self.log_zero_guard_value
        if isinstance(v, str):
            return {"tiny": torch.finfo(torch.float32).tiny, "eps": torch.finfo(torch.float32).eps}[v]
        return v

    def get_seq_len(self, seq_len):
        pad_amount = self.stft_pad_amount * 2 if self.stft_pad_amount is not None else self.n_fft // 2 * 2
        seq_len = torch.floor_divide((seq_len + pad_amount - self.n_fft), self.hop_length) + 1
        return seq_len.to(dtype=torch.long)

    def forward(self, x, seq_len, linear_spec=False, dtype=torch.float32):
        """``dtype=torch.float64`` evaluates the SAME statements on the same float32 input and tables in double
        precision: the exact-arithmetic yardstick against which the float32 spread of the reference's own
        ``torch.stft`` path is measured (tests)."""
        x = torch.as_tensor(np.asarray(x, dtype=np.float32)).clone().to(dtype)
        seq_len = self.get_seq_len(torch.as_tensor(np.asarray(seq_len)))
        if self.stft_pad_amount is not None:
            x = torch.nn.functional.pad(x.unsqueeze(1), (self.stft_pad_amount, self.stft_pad_amount),
                                        "reflect").squeeze(1)
        if self.preemph is not None:
            x = torch.cat((x[:, 0].unsqueeze(1), x[:, 1:] - self.preemph * x[:, :-1]), dim=1)
        x = torch.stft(x, n_fft=self.n_fft, hop_length=self.hop_length, win_length=self.win_length,
                       center=False if self.exact_pad else True,
                       window=self.window.to(dtype=dtype) if self.window is not None else None,
                       return_complex=True)
        guard = 0 if not self.use_grads else CONSTANT
        x = torch.view_as_real(x)
        x = torch.sqrt(x.pow(2).sum(-1) + guard)
        if self.mag_power != 1.0:
            x = x.pow(self.mag_power)
        if linear_spec:
            return x.numpy(), seq_len.numpy()
        x = torch.matmul(self.fb.to(x.dtype), x)
        if self.log:
            if self.log_zero_guard_type == "add":
                x = torch.log(x + self._guard(x))
            elif self.log_zero_guard_type == "clamp":
                x = torch.log(torch.clamp(x, min=self._guard(x)))
            else:
                raise ValueError("log_zero_guard_type was not understood")
        if self.frame_splicing > 1:
            x = splice_frames(x, self.frame_splicing)
        if self.normalize:
            x = normalize_batch(x, seq_len, normalize_type=self.normalize)
        max_len = x.size(-1)
        mask = torch.arange(max_len).repeat(x.size(0), 1) >= seq_len.unsqueeze(1)
        x = x.masked_fill(mask.unsqueeze(1), self.pad_value)
        pad_to = self.pad_to
        if pad_to == "max":
            x = torch.nn.functional.pad(x, (0, int(self.max_length) - x.size(-1)), value=self.pad_value)
        elif pad_to > 0:
            pad_amt = x.size(-1) % pad_to
            if pad_amt != 0:
                x = torch.nn.functional.pad(x, (0, pad_to - pad_amt), value=self.pad_value)
        return x.numpy(), seq_len.numpy()
