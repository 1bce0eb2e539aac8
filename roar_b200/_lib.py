"""ctypes binding of libroar_sup.so (include/roar_sup.h).  No fallback: if the library is missing or
fails to load, every compute entry point raises."""
import ctypes
import os

from .config import RoarSupConfig

HERE = os.path.dirname(os.path.abspath(__file__))
# ROAR_SUP_LIB selects another build of the same library (kernel A/B measurements, diagnostic builds); the
# default is the in-tree libroar_sup.so next to this file
LIB_PATH = os.environ.get("ROAR_SUP_LIB") or os.path.join(HERE, "libroar_sup.so")

# every symbol include/roar_sup.h declares
SYMBOLS = [
    "roar_sup_config_default", "roar_sup_abi_version", "roar_sup_last_error", "roar_sup_create",
    "roar_sup_destroy", "roar_sup_num_frames", "roar_sup_pyin_num_frames", "roar_sup_pyin_geometry",
    "roar_sup_host_mel_filterbank", "roar_sup_host_window", "roar_sup_host_pyin_log_transition",
    "roar_sup_host_pyin_beta_probs", "roar_sup_workspace_bytes", "roar_sup_logmel_energy",
    "roar_sup_pyin", "roar_sup_align_prior", "roar_sup_align_prior_interp", "roar_sup_pitch_partials_init", "roar_sup_pitch_partials",
    "roar_sup_pitch_partials_grouped", "roar_fbank_out_frames", "roar_fbank_forward",
    "roar_sup_set_profiling", "roar_sup_profile_read", "roar_sup_trim", "roar_fbank_backward",
    "roar_fbank_workspace_bytes", "roar_sup_pcm16_to_f32",
    "roar_sup_wav_probe_batch", "roar_sup_wav_read_batch", "roar_sup_pt_write_batch", "roar_sup_debug_counters",
    "roar_sup_upload", "roar_sup_resample",
]
ABI_VERSION = 2
N_KERNEL_IDS = 12      # ROAR_K_COUNT

_lib = None


class RoarSupError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RoarSupError(
            f"{LIB_PATH} is missing: build it with `python -m roar_b200.build` "
            "(nvcc, sm_100a). roar_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    cfgp = ctypes.POINTER(RoarSupConfig)
    lib.roar_sup_config_default.argtypes = [cfgp]
    lib.roar_sup_config_default.restype = None
    lib.roar_sup_abi_version.restype = ctypes.c_int
    lib.roar_sup_last_error.restype = ctypes.c_char_p
    lib.roar_sup_create.argtypes = [cfgp, ctypes.c_int, ctypes.POINTER(vp)]
    lib.roar_sup_destroy.argtypes = [vp]
    lib.roar_sup_destroy.restype = None
    lib.roar_sup_num_frames.argtypes = [vp, i64]
    lib.roar_sup_num_frames.restype = i64
    lib.roar_sup_pyin_num_frames.argtypes = [vp, i64]
    lib.roar_sup_pyin_num_frames.restype = i64
    lib.roar_sup_pyin_geometry.argtypes = [vp, ctypes.POINTER(i32 * 8)]
    lib.roar_sup_host_mel_filterbank.argtypes = [cfgp, vp]
    lib.roar_sup_host_window.argtypes = [cfgp, vp]
    lib.roar_sup_host_pyin_log_transition.argtypes = [cfgp, vp, i64]
    lib.roar_sup_host_pyin_beta_probs.argtypes = [cfgp, vp]
    lib.roar_sup_workspace_bytes.argtypes = [vp, i32, i64, i64]
    lib.roar_sup_workspace_bytes.restype = ctypes.c_size_t
    lib.roar_sup_logmel_energy.argtypes = [vp, vp, vp, vp, i32, vp, i64, vp, vp, vp, ctypes.c_size_t, vp]
    lib.roar_sup_pyin.argtypes = [vp, vp, vp, vp, i32, vp, i64, i32, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.roar_sup_align_prior.argtypes = [vp, vp, vp, i32, vp, i32, f64, vp, vp]
    lib.roar_sup_align_prior_interp.argtypes = [vp, vp, vp, i32, vp, i32, i32, i32, vp, vp]
    lib.roar_sup_trim.argtypes = [vp, vp, vp, vp, i32, i32, f64, f64, i32, i32, vp, vp, vp]
    lib.roar_sup_pitch_partials_init.argtypes = [vp, vp, i32, vp]
    lib.roar_sup_pitch_partials.argtypes = [vp, vp, i64, vp, vp]
    lib.roar_sup_pitch_partials_grouped.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.roar_fbank_out_frames.argtypes = [vp, i64]
    lib.roar_fbank_out_frames.restype = i64
    lib.roar_fbank_forward.argtypes = [vp, vp, vp, i32, i64, vp, vp, vp, ctypes.c_size_t, vp]
    lib.roar_fbank_backward.argtypes = [vp, vp, vp, i32, i64, vp, vp, vp, ctypes.c_size_t, vp]
    lib.roar_fbank_workspace_bytes.argtypes = [vp, i32]
    lib.roar_fbank_workspace_bytes.restype = ctypes.c_size_t
    lib.roar_sup_pcm16_to_f32.argtypes = [vp, vp, i64, vp, vp]
    lib.roar_sup_wav_probe_batch.argtypes = [vp, i32, vp, i32]
    lib.roar_sup_wav_read_batch.argtypes = [vp, vp, i32, vp, vp, i32, i32, vp, vp, i32]
    lib.roar_sup_pt_write_batch.argtypes = [vp, i32, vp, vp, vp, vp, i32]
    lib.roar_sup_debug_counters.argtypes = [vp, i32, ctypes.c_int]
    lib.roar_sup_upload.argtypes = [vp, vp, vp, ctypes.c_size_t, vp]
    lib.roar_sup_resample.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.roar_sup_set_profiling.argtypes = [vp, ctypes.c_int]
    lib.roar_sup_profile_read.argtypes = [vp, vp, vp, ctypes.c_int]
    if lib.roar_sup_abi_version() != ABI_VERSION:
        raise RoarSupError("libroar_sup.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RoarSupError(f"libroar_sup error {rc}: {load().roar_sup_last_error().decode()}")
