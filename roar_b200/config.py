"""Configuration of one extractor handle: the reference's keyword arguments -> ``roar_sup_config``.

Field names follow ``TTSDataset.__init__`` (``roar/collections/tts/data/dataset.py:71-365``),
``FilterbankFeatures.__init__`` (``roar/collections/asr/parts/preprocessing/features.py:196-345``)
and the ``librosa.pyin`` call at ``dataset.py:696-703``.
"""
import ctypes
from dataclasses import dataclass, fields
from typing import Optional

WINDOWS = {"hann": 0, "hamming": 1, "blackman": 2, "bartlett": 3, "none": 4, None: 4}
LOG_MODES = {None: 0, "clamp": 1, "add": 2}
NORMALIZE = {None: 0, "per_feature": 1, "all_features": 2}

FLOAT32_TINY = 1.1754943508222875e-38  # torch.finfo(torch.float32).tiny
FLOAT32_EPS = 1.1920928955078125e-07
PITCH_FMIN_C2 = 65.40639132514966      # librosa.note_to_hz("C2")
PITCH_FMAX_C7 = 2093.004522404789      # librosa.note_to_hz("C7")


class RoarSupConfig(ctypes.Structure):
    """Binary layout of ``roar_sup_config`` (include/roar_sup.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_int32), ("sample_rate", ctypes.c_int32),
        ("n_fft", ctypes.c_int32), ("win_length", ctypes.c_int32), ("hop_length", ctypes.c_int32),
        ("window", ctypes.c_int32), ("n_mels", ctypes.c_int32), ("mel_norm", ctypes.c_int32),
        ("fmin", ctypes.c_double), ("fmax", ctypes.c_double),
        ("spec_floor", ctypes.c_double), ("mag_power", ctypes.c_double),
        ("log_mode", ctypes.c_int32), ("exact_pad", ctypes.c_int32), ("log_guard", ctypes.c_double),
        ("has_preemph", ctypes.c_int32), ("normalize", ctypes.c_int32),
        ("preemph", ctypes.c_double), ("pad_value", ctypes.c_double),
        ("pad_to", ctypes.c_int32), ("energy_mode", ctypes.c_int32),
        ("pitch_fmin", ctypes.c_double), ("pitch_fmax", ctypes.c_double),
        ("pyin_frame_length", ctypes.c_int32), ("pyin_win_length", ctypes.c_int32),
        ("pyin_hop_length", ctypes.c_int32), ("n_thresholds", ctypes.c_int32),
        ("beta_a", ctypes.c_double), ("beta_b", ctypes.c_double),
        ("boltzmann_parameter", ctypes.c_double), ("resolution", ctypes.c_double),
        ("max_transition_rate", ctypes.c_double), ("switch_prob", ctypes.c_double),
        ("no_trough_prob", ctypes.c_double),
    ]


@dataclass(frozen=True)
class SupConfig:
    """Defaults = ``scripts/dataset_processing/tts/rasa/ds_conf/ds_for_fastpitch_align.yaml:12-27``."""
    sample_rate: int = 22050
    n_fft: int = 1024
    win_length: Optional[int] = None      # None -> n_fft (dataset.py:302)
    hop_length: Optional[int] = None      # None -> n_fft // 4 (dataset.py:304)
    window: Optional[str] = "hann"
    n_mels: int = 80
    lowfreq: float = 0.0
    highfreq: Optional[float] = None      # None -> sample_rate / 2
    mel_norm: Optional[str] = "slaney"
    spec_floor: float = 1e-9              # EPSILON under the sqrt (dataset.py:529)
    mag_power: float = 1.0
    log_mode: Optional[str] = "clamp"
    log_guard: float = FLOAT32_TINY
    exact_pad: bool = False
    preemph: Optional[float] = None
    normalize: Optional[str] = None
    pad_value: float = 0.0
    pad_to: int = 0
    energy_mode: str = "spectrum"        # "spectrum": TTSDataset energy; "features": EnergyFeaturizer (norm over mel axis)
    pitch_fmin: float = PITCH_FMIN_C2
    pitch_fmax: float = PITCH_FMAX_C7
    pyin: bool = True                         # False: mel-only handle (no pYIN tables; FilterbankFeatures)
    pyin_frame_length: Optional[int] = None   # None -> win_length (what TTSDataset passes)
    pyin_win_length: Optional[int] = None
    pyin_hop_length: Optional[int] = None     # None -> frame_length // 4 (TTSDataset does not pass hop)
    n_thresholds: int = 100
    beta_parameters: tuple = (2.0, 18.0)
    boltzmann_parameter: float = 2.0
    resolution: float = 0.1
    max_transition_rate: float = 35.92
    switch_prob: float = 0.01
    no_trough_prob: float = 0.01

    @property
    def win(self) -> int:
        return self.win_length or self.n_fft

    @property
    def hop(self) -> int:
        return self.hop_length or self.n_fft // 4

    @property
    def pyin_frame(self) -> int:
        return self.pyin_frame_length or self.win

    @property
    def pyin_hop(self) -> int:
        return self.pyin_hop_length or self.pyin_frame // 4

    def to_c(self) -> RoarSupConfig:
        if self.window not in WINDOWS:
            raise NotImplementedError(f"Current implementation doesn't support {self.window} window. "
                                      f"Please choose one from {[k for k in WINDOWS if k]}.")
        if self.mel_norm not in ("slaney", None):
            raise ValueError(f"unsupported mel_norm {self.mel_norm!r}")
        c = RoarSupConfig()
        c.struct_size = ctypes.sizeof(RoarSupConfig)
        c.sample_rate = int(self.sample_rate)
        c.n_fft = int(self.n_fft)
        c.win_length = int(self.win)
        c.hop_length = int(self.hop)
        c.window = WINDOWS[self.window]
        c.n_mels = int(self.n_mels)
        c.mel_norm = 1 if self.mel_norm == "slaney" else 0
        c.fmin = float(self.lowfreq)
        c.fmax = float(self.highfreq) if self.highfreq else 0.0
        c.spec_floor = float(self.spec_floor)
        c.mag_power = float(self.mag_power)
        c.log_mode = LOG_MODES[self.log_mode]
        c.exact_pad = int(bool(self.exact_pad))
        c.log_guard = float(self.log_guard)
        c.has_preemph = 0 if self.preemph is None else 1
        c.preemph = 0.0 if self.preemph is None else float(self.preemph)
        c.normalize = NORMALIZE[self.normalize]
        c.pad_value = float(self.pad_value)
        c.pad_to = int(self.pad_to)
        if self.energy_mode not in ("spectrum", "features"):
            raise ValueError(f"unsupported energy_mode {self.energy_mode!r}")
        c.energy_mode = 1 if self.energy_mode == "features" else 0
        c.pitch_fmin = float(self.pitch_fmin)
        c.pitch_fmax = float(self.pitch_fmax)
        c.pyin_frame_length = int(self.pyin_frame) if self.pyin else 0
        c.pyin_win_length = int(self.pyin_win_length or 0)
        c.pyin_hop_length = int(self.pyin_hop_length or 0)
        c.n_thresholds = int(self.n_thresholds)
        c.beta_a, c.beta_b = float(self.beta_parameters[0]), float(self.beta_parameters[1])
        c.boltzmann_parameter = float(self.boltzmann_parameter)
        c.resolution = float(self.resolution)
        c.max_transition_rate = float(self.max_transition_rate)
        c.switch_prob = float(self.switch_prob)
        c.no_trough_prob = float(self.no_trough_prob)
        return c

    def replace(self, **kw) -> "SupConfig":
        d = {f.name: getattr(self, f.name) for f in fields(self)}
        d.update(kw)
        return SupConfig(**d)
