"""TEST INFRASTRUCTURE ONLY: CPU emulation of the CUDA kernels' per-thread code (see hostemu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libhostemu.so")
        src = os.path.join(HERE, "hostemu.cpp")
        deps = [src] + [os.path.join(HERE, "..", "..", "roar_b200", "csrc", f)
                        for f in os.listdir(os.path.join(HERE, "..", "..", "roar_b200", "csrc"))]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
        _LIB = ctypes.CDLL(so)
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def logmel_energy(cfg, audio):
    c = cfg.to_c()
    audio = _f32(audio)
    L = len(audio)
    p = (cfg.n_fft - cfg.hop) // 2 if cfg.exact_pad else cfg.n_fft // 2
    T = (L + 2 * p - cfg.n_fft) // cfg.hop + 1
    lm = np.zeros((cfg.n_mels, T), dtype=np.float32)
    en = np.zeros(T, dtype=np.float32)
    rc = lib().emu_logmel_energy(ctypes.byref(c), audio.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(L),
                                 lm.ctypes.data_as(ctypes.c_void_p), en.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return lm, en


def pyin(cfg, audio, n_lags):
    c = cfg.to_c()
    audio = _f32(audio)
    L = len(audio)
    T = 1 + L // cfg.pyin_hop
    f0 = np.zeros(T, np.float32)
    vf = np.zeros(T, np.float32)
    vp = np.zeros(T, np.float32)
    cm = np.zeros((T, n_lags), np.float64)
    st = np.zeros(T, np.int32)
    rc = lib().emu_pyin(ctypes.byref(c), audio.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(L),
                        f0.ctypes.data_as(ctypes.c_void_p), vf.ctypes.data_as(ctypes.c_void_p),
                        vp.ctypes.data_as(ctypes.c_void_p), cm.ctypes.data_as(ctypes.c_void_p),
                        st.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return f0, vf, vp, cm, st


def prior(N, M, scaling=1.0):
    out = np.zeros((M, N), np.float32)
    rc = lib().emu_prior(ctypes.c_int32(N), ctypes.c_int32(M), ctypes.c_double(scaling),
                         out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def prior_interp(text_len, mel_len, round_mel=50, round_text=10):
    out = np.zeros((mel_len, text_len), np.float32)
    rc = lib().emu_prior_interp(ctypes.c_int32(text_len), ctypes.c_int32(mel_len), ctypes.c_int32(round_mel),
                                ctypes.c_int32(round_text), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def uniform_row_check(cfg, n_rand=64):
    c = cfg.to_c()
    out = np.zeros(4, np.float64)
    rc = lib().emu_uniform_row_check(ctypes.byref(c), ctypes.c_int32(n_rand), out.ctypes.data_as(ctypes.c_void_p))
    return rc, out


def trim(audio, top_db=60.0, ref_value=0.0, frame_length=2048, hop_length=512):
    audio = _f32(audio)
    out = np.zeros(2, np.int64)
    rc = lib().emu_trim(audio.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(audio)), ctypes.c_double(top_db),
                        ctypes.c_double(ref_value), ctypes.c_int32(frame_length), ctypes.c_int32(hop_length),
                        out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return int(out[0]), int(out[1])


def fbank_backward(cfg, x, length, grad_out):
    c = cfg.to_c()
    x = _f32(x)
    go = _f32(grad_out)
    gx = np.zeros(len(x), np.float32)
    rc = lib().emu_fbank_backward(ctypes.byref(c), x.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(x)),
                                  ctypes.c_int64(int(length)), go.ctypes.data_as(ctypes.c_void_p),
                                  gx.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return gx
