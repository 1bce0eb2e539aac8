"""B200 versions of the reference's newer TTS featurizers.

Mirrors ``roar/collections/tts/parts/preprocessing/features.py``:

* ``MelSpectrogramFeaturizer`` (``:166-274``) -- ``AudioToMelSpectrogramPreprocessor`` with
  ``n_fft = win_length``, magnitude mel, ``log(x + 1.0)``, ``mel_norm=None``, no normalisation, no
  pre-emphasis, no dither; saves ``[mel_dim, T]`` (no leading 1).
* ``EnergyFeaturizer`` (``:277-339``) -- ``torch.linalg.norm(mel_spec, axis=0)`` -> ``[T]``.
* ``PitchFeaturizer`` (``:342-469``) -- ``librosa.pyin(audio, fmin, fmax, frame_length=win_length,
  hop_length=hop_length, sr, fill_na=0.0)`` (hop IS passed here, unlike ``TTSDataset``); pitch float32,
  voiced mask **bool**, voiced probability float32.

Same constructor keywords, same file layout (``<feature_dir>/<feature_name>/<rel audio path>.pt``,
``_get_feature_filepath`` ``:84-103``), same ``load`` / ``collate_fn``.  What differs is the unit of
work: the reference's ``save(manifest_entry, ...)`` decodes and processes ONE file on the CPU; here
``compute_batch`` / ``save_batch`` take a list of entries and run the whole batch through the CUDA
library.  ``save`` is kept as a one-entry batch.  Audio must already be at ``sample_rate`` (resampling is
row N3, out of scope).
"""
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from .config import PITCH_FMAX_C7, PITCH_FMIN_C2, SupConfig
from .extractor import SupDataExtractor, split_frames


def get_abs_rel_paths(input_path: Path, base_path: Path) -> Tuple[Path, Path]:
    """``tts_dataset_utils.get_abs_rel_paths`` (``:13-31``)."""
    input_path = Path(input_path)
    if input_path.is_absolute():
        return input_path, input_path.relative_to(base_path)
    return Path(base_path) / input_path, input_path


def get_feature_filepath(manifest_entry: Dict[str, Any], audio_dir: Path, feature_dir: Path, feature_name: str) -> Path:
    """``<audio_dir>/speaker1/audio1.wav`` -> ``<feature_dir>/<feature_name>/speaker1/audio1.pt``."""
    _, rel = get_abs_rel_paths(Path(manifest_entry["audio_filepath"]), Path(audio_dir))
    return Path(feature_dir) / feature_name / rel.with_suffix(".pt")


def stack_tensors(tensors: List[Tensor], max_lens: List[int], pad_value: float = 0.0) -> Tensor:
    """``tts_dataset_utils.stack_tensors`` (``:101-125``): pad the trailing axes, then stack."""
    padded = []
    for t in tensors:
        padding = []
        for i, max_len in enumerate(max_lens, 1):
            padding += [0, max_len - t.shape[-i]]
        padded.append(torch.nn.functional.pad(t, pad=padding, value=pad_value))
    return torch.stack(padded)


def _default_loader(path: Path, sample_rate: int) -> np.ndarray:
    from .extract_sup_data import load_wav
    return load_wav(str(path), sample_rate)


class _BatchFeaturizer:
    """save / load / collate plumbing shared by the three featurizers."""

    feature_names: List[Optional[str]] = []
    sample_rate: int = 22050

    def _audio(self, entries: Sequence[Dict[str, Any]], audio_dir: Path, wavs=None) -> List[np.ndarray]:
        if wavs is not None:
            return [np.ascontiguousarray(w, dtype=np.float32) for w in wavs]
        return [_default_loader(get_abs_rel_paths(Path(e["audio_filepath"]), Path(audio_dir))[0], self.sample_rate)
                for e in entries]

    def compute_batch(self, wavs: Sequence[np.ndarray]) -> Dict[str, List[Tensor]]:
        raise NotImplementedError

    def save_batch(self, manifest_entries: Sequence[Dict[str, Any]], audio_dir: Path, feature_dir: Path,
                   wavs=None) -> None:
        feats = self.compute_batch(self._audio(manifest_entries, audio_dir, wavs))
        for name, tensors in feats.items():
            if name is None:
                continue
            for entry, t in zip(manifest_entries, tensors):
                path = get_feature_filepath(entry, audio_dir, feature_dir, name)
                path.parent.mkdir(exist_ok=True, parents=True)
                torch.save(t.cpu(), path)

    def save(self, manifest_entry: Dict[str, Any], audio_dir: Path, feature_dir: Path) -> None:
        self.save_batch([manifest_entry], audio_dir, feature_dir)

    def load(self, manifest_entry: Dict[str, Any], audio_dir: Path, feature_dir: Path) -> Dict[str, Tensor]:
        out = {}
        for name in self.feature_names:
            if name is not None:
                out[name] = torch.load(get_feature_filepath(manifest_entry, audio_dir, feature_dir, name))
        return out

    def collate_fn(self, train_batch: List[Dict[str, Tensor]]) -> Dict[str, Tensor]:
        out = {}
        for name in self.feature_names:
            if name is None:
                continue
            tensors = [ex[name] for ex in train_batch]
            max_len = max(t.shape[-1] for t in tensors)
            out[name] = stack_tensors(tensors, max_lens=[max_len])
        return out


class MelSpectrogramFeaturizer(_BatchFeaturizer):
    def __init__(self, feature_name: str = "mel_spec", sample_rate: int = 22050, mel_dim: int = 80,
                 win_length: int = 1024, hop_length: int = 256, lowfreq: int = 0, highfreq: int = 8000,
                 log: bool = True, log_zero_guard_type: str = "add", log_zero_guard_value: float = 1.0,
                 mel_norm=None, device=None) -> None:
        if log_zero_guard_type not in ("add", "clamp"):
            raise ValueError(f"log_zero_guard_type must be 'add' or 'clamp', got {log_zero_guard_type!r}")
        self.feature_name = feature_name
        self.feature_names = [feature_name]
        self.sample_rate, self.win_length, self.hop_length = sample_rate, win_length, hop_length
        # FilterbankFeatures: |X| without the 1e-9 floor (features.py:408-410 guard is 0 without grads)
        self.cfg = SupConfig(sample_rate=sample_rate, n_fft=win_length, win_length=win_length, hop_length=hop_length,
                             n_mels=mel_dim, lowfreq=lowfreq, highfreq=highfreq, mel_norm=mel_norm, spec_floor=0.0,
                             mag_power=1.0, log_mode=(log_zero_guard_type if log else None),
                             log_guard=float(log_zero_guard_value), energy_mode="features")
        self._device = device
        self._ex = None

    @property
    def extractor(self) -> SupDataExtractor:
        if self._ex is None:
            self._ex = SupDataExtractor(self.cfg, self._device)
        return self._ex

    def compute_mel_and_energy(self, wavs: Sequence[np.ndarray], want_mel=True, want_energy=False):
        ex = self.extractor
        lm, en, fo = ex.log_mel_energy(ex.pack(list(wavs)), want_log_mel=True, want_energy=want_energy)
        mels = split_frames(lm, fo, self.cfg.n_mels) if want_mel else None
        return mels, (split_frames(en, fo) if want_energy else None)

    def compute_batch(self, wavs):
        mels, _ = self.compute_mel_and_energy(wavs)
        return {self.feature_name: mels}


class EnergyFeaturizer(_BatchFeaturizer):
    def __init__(self, spec_featurizer: MelSpectrogramFeaturizer, feature_name: str = "energy") -> None:
        self.feature_name = feature_name
        self.feature_names = [feature_name]
        self.spec_featurizer = spec_featurizer
        self.sample_rate = spec_featurizer.sample_rate

    def compute_batch(self, wavs):
        _, en = self.spec_featurizer.compute_mel_and_energy(wavs, want_mel=False, want_energy=True)
        return {self.feature_name: en}


class PitchFeaturizer(_BatchFeaturizer):
    def __init__(self, pitch_name: Optional[str] = "pitch", voiced_mask_name: Optional[str] = "voiced_mask",
                 voiced_prob_name: Optional[str] = None, sample_rate: int = 22050, win_length: int = 1024,
                 hop_length: int = 256, pitch_fmin: float = PITCH_FMIN_C2, pitch_fmax: float = PITCH_FMAX_C7,
                 device=None) -> None:
        self.pitch_name, self.voiced_mask_name, self.voiced_prob_name = pitch_name, voiced_mask_name, voiced_prob_name
        self.feature_names = [pitch_name, voiced_mask_name, voiced_prob_name]
        self.sample_rate, self.win_length, self.hop_length = sample_rate, win_length, hop_length
        self.cfg = SupConfig(sample_rate=sample_rate, n_fft=win_length, win_length=win_length, hop_length=hop_length,
                             pitch_fmin=pitch_fmin, pitch_fmax=pitch_fmax, pyin_frame_length=win_length,
                             pyin_hop_length=hop_length)
        self._device = device
        self._ex = None

    @property
    def extractor(self) -> SupDataExtractor:
        if self._ex is None:
            self._ex = SupDataExtractor(self.cfg, self._device)
        return self._ex

    def compute_batch(self, wavs):
        ex = self.extractor
        f0, vf, vp, fo = ex.pyin(ex.pack(list(wavs)))
        out = {}
        if self.pitch_name is not None:
            out[self.pitch_name] = split_frames(f0, fo)
        if self.voiced_mask_name is not None:
            out[self.voiced_mask_name] = [m.to(torch.bool) for m in split_frames(vf, fo)]
        if self.voiced_prob_name is not None:
            out[self.voiced_prob_name] = split_frames(vp, fo)
        return out
