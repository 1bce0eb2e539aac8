"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, exports every symbol
include/roar_sup.h declares, and its host-built tables equal the oracle's.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import melfb, pyin as opyin
from roar_b200 import _lib, build
from roar_b200.config import RoarSupConfig, SupConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "roar_sup.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(roar_(?:sup|fbank)_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in roar_sup.h but not exported"
    for name in _lib.SYMBOLS:
        assert name in declared


def test_config_struct_matches_c_defaults(lib):
    c = RoarSupConfig()
    lib.roar_sup_config_default(ctypes.byref(c))
    assert c.struct_size == ctypes.sizeof(RoarSupConfig)
    py = SupConfig(highfreq=8000.0).to_c()
    for name, _ in RoarSupConfig._fields_:
        a, b = getattr(c, name), getattr(py, name)
        if name == "preemph":
            continue  # unused when has_preemph == 0
        assert a == pytest.approx(b, rel=1e-7), name


def test_create_without_gpu_fails_loudly(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = SupConfig().to_c()
    h = ctypes.c_void_p()
    rc = lib.roar_sup_create(ctypes.byref(c), 0, ctypes.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.roar_sup_last_error()
    from roar_b200.extractor import SupDataExtractor
    with pytest.raises(_lib.RoarSupError):
        SupDataExtractor(SupConfig())


def test_invalid_config_rejected(lib):
    bad = SupConfig(n_fft=1000).to_c()
    out = np.zeros(10, np.float32)
    assert lib.roar_sup_host_window(ctypes.byref(bad), out.ctypes.data_as(ctypes.c_void_p)) == -1
    with pytest.raises(NotImplementedError):
        SupConfig(window="kaiser").to_c()


@pytest.mark.parametrize("kw", [
    dict(sample_rate=22050, n_fft=1024, n_mels=80, lowfreq=0.0, highfreq=8000.0),
    dict(sample_rate=44100, n_fft=2048, n_mels=80, lowfreq=0.0, highfreq=None),
    dict(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, highfreq=None, pyin_frame_length=512),
    dict(sample_rate=22050, n_fft=1024, n_mels=80, highfreq=None, mel_norm=None),
    dict(sample_rate=16000, n_fft=512, n_mels=64, lowfreq=50.0, highfreq=7600.0)])
def test_host_mel_filterbank_equals_oracle(lib, kw):
    cfg = SupConfig(**kw)
    c = cfg.to_c()
    out = np.zeros((cfg.n_mels, cfg.n_fft // 2 + 1), np.float32)
    assert lib.roar_sup_host_mel_filterbank(ctypes.byref(c), out.ctypes.data_as(ctypes.c_void_p)) == 0
    ref = melfb.mel_filterbank(cfg.sample_rate, cfg.n_fft, cfg.n_mels, cfg.lowfreq, cfg.highfreq, cfg.mel_norm)
    assert np.abs(out - ref).max() <= 1e-7 * ref.max()
    assert (out != ref).mean() < 1e-3       # float32-ulp differences at most, and rare
    assert np.array_equal(out != 0, ref != 0)


@pytest.mark.parametrize("window,win,n_fft", [("hann", 1024, 1024), ("hann", 400, 512), ("hamming", 400, 512),
                                              ("blackman", 320, 512), ("bartlett", 401, 512)])
def test_host_window_equals_torch(lib, window, win, n_fft):
    c = SupConfig(n_fft=n_fft, win_length=win, hop_length=128, window=window, pyin_frame_length=n_fft).to_c()
    out = np.zeros(n_fft, np.float32)
    assert lib.roar_sup_host_window(ctypes.byref(c), out.ctypes.data_as(ctypes.c_void_p)) == 0
    fn = dict(hann=torch.hann_window, hamming=torch.hamming_window, blackman=torch.blackman_window,
              bartlett=torch.bartlett_window)[window]
    ref = np.zeros(n_fft, np.float32)
    left = (n_fft - win) // 2
    ref[left:left + win] = fn(win, periodic=False).numpy()
    assert np.abs(out - ref).max() < 3e-7


@pytest.mark.parametrize("sr,frame", [(22050, 1024), (44100, 2048), (22050, 2048)])
def test_host_pyin_tables_equal_oracle(lib, sr, frame):
    cfg = SupConfig(sample_rate=sr, n_fft=frame, hop_length=frame // 4)
    c = cfg.to_c()
    npb, nbps = opyin.n_pitch_bins_for(cfg.pitch_fmin, cfg.pitch_fmax)
    tr, p_init, tw = opyin.hmm_tables(sr, frame // 4, npb, nbps)
    out = np.zeros((2 * npb, 2 * npb), np.float64)
    assert lib.roar_sup_host_pyin_log_transition(ctypes.byref(c), out.ctypes.data_as(ctypes.c_void_p), out.size) == 0
    ref = np.log(tr + np.finfo(np.float64).tiny)
    # bit-exact incl. numpy's pairwise row sums, except where numpy's SIMD log and glibc's log round
    # differently (a handful of entries, 1 ulp)
    assert (out != ref).mean() < 1e-4, (out != ref).sum()
    assert np.abs(out - ref).max() <= 4 * np.finfo(np.float64).eps * 8
    bp = np.zeros(100, np.float64)
    assert lib.roar_sup_host_pyin_beta_probs(ctypes.byref(c), bp.ctypes.data_as(ctypes.c_void_p)) == 0
    _, ref_bp = opyin.beta_threshold_prior()
    assert np.abs(bp - ref_bp).max() < 1e-15
    big = ref_bp > 1e-12
    assert np.abs(bp[big] / ref_bp[big] - 1).max() < 1e-10
