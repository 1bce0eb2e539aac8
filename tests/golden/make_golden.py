"""Generate the golden fixtures under tests/golden/ from the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What comes from the reference itself (imported by file path, source untouched):
  * ``prior_*.npz``       ``beta_binomial_prior_distribution`` / ``BetaBinomialInterpolator``
                          (``roar/collections/tts/parts/utils/tts_dataset_utils.py:69-149``)
  * ``fbank_*.npz``       ``FilterbankFeatures.forward``
                          (``roar/collections/asr/parts/preprocessing/features.py:196-461``).
                          The module imports ``librosa`` (absent here) only for
                          ``librosa.filters.mel``; a stub module supplies the restated
                          ``oracle.melfb.mel_filterbank`` for that one call, everything else
                          (STFT, pre-emphasis, log guard, normalize_batch, masking, padding) is
                          the reference's code executing.
What cannot come from the reference (needs librosa / hydra / lightning / soundfile):
  * ``spec_*.npz``        the statements of ``TTSDataset.get_spec/get_log_mel`` and the energy
                          line (``dataset.py:324-333,524-537,751-753``) executed via
                          ``oracle.spec`` (same torch calls), so it pins torch.stft behaviour.
  * ``pyin_*.npz``        ``oracle.pyin`` with the *dense* Viterbi -- parity unpinned, kept to
                          detect drift of the oracle itself and to give the GPU tests
                          committed vectors.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import melfb, pyin as opyin, spec as ospec  # noqa: E402
from roar_b200 import synth  # noqa: E402


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_prior_module():
    return load_by_path("ref_tts_dataset_utils",
                        f"{REF}/roar/collections/tts/parts/utils/tts_dataset_utils.py")


def reference_features_module():
    """Import the reference's features.py with stubs for what this image lacks."""
    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {"__init__": lambda self, *a, **kw: None})

    lib = types.ModuleType("librosa")
    lib.filters = types.ModuleType("librosa.filters")
    lib.filters.mel = lambda *, sr, n_fft, n_mels, fmin, fmax, norm="slaney": melfb.mel_filterbank(
        sr, n_fft, n_mels, fmin, fmax, norm)
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = lib.filters
    for name in ["roar", "roar.collections", "roar.collections.asr", "roar.collections.asr.parts",
                 "roar.collections.asr.parts.preprocessing",
                 "roar.collections.asr.parts.preprocessing.perturb",
                 "roar.collections.asr.parts.preprocessing.segment", "roar.utils"]:
        sys.modules[name] = _Any(name)
    logmod = types.ModuleType("roar.utils.logging")
    for fn in ("info", "debug", "warning", "error"):
        setattr(logmod, fn, lambda *a, **k: None)
    sys.modules["roar.utils"].logging = logmod
    sys.modules["roar.utils.logging"] = logmod
    return load_by_path("ref_features", f"{REF}/roar/collections/asr/parts/preprocessing/features.py")


def main():
    # ---------------------------------------------------------------- prior (reference code)
    ref = reference_prior_module()
    cases = [(7, 13), (3, 5), (2, 9), (100, 560), (120, 801), (57, 222), (200, 1500), (1, 4)]
    out = {}
    for n, m in cases:
        out[f"p_{n}_{m}"] = ref.beta_binomial_prior_distribution(n, m)
    interp = ref.BetaBinomialInterpolator()
    for w, h in [(560, 100), (333, 47), (49, 9)]:
        out[f"i_{w}_{h}"] = interp(w, h)
    np.savez_compressed(f"{HERE}/prior_ref.npz", **out)
    print("prior_ref.npz", {k: v.shape for k, v in out.items()})

    # ---------------------------------------------------------------- fbank (reference code)
    feats = reference_features_module()
    man = synth.corpus_manifest("C5", n_utts=4)
    wavs = [synth.synth_utterance(5, u.utt_id, min(u.n_samples, 8000 * (2 + i)), 16000, u.speaker)
            for i, u in enumerate(man)]
    lens = np.array([len(w) for w in wavs], dtype=np.int64)
    x = np.zeros((len(wavs), lens.max()), dtype=np.float32)
    for i, w in enumerate(wavs):
        x[i, : len(w)] = w
    variants = {
        "asr_default": dict(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=80,
                            n_fft=512, dither=0.0),
        "tts_fastpitch": dict(sample_rate=22050, n_window_size=1024, n_window_stride=256, nfilt=80,
                              n_fft=1024, lowfreq=0, highfreq=8000, normalize=None, preemph=None,
                              log=True, log_zero_guard_type="add", log_zero_guard_value=1.0,
                              mag_power=1.0, pad_to=1, pad_value=0.0, dither=0.0),
        "clamp_allfeat_exactpad": dict(sample_rate=16000, n_window_size=400, n_window_stride=160,
                                       nfilt=64, n_fft=512, dither=0.0, exact_pad=True,
                                       normalize="all_features", log_zero_guard_type="clamp",
                                       log_zero_guard_value="tiny", pad_to=8, pad_value=-1.0,
                                       mel_norm=None),
    }
    out = {"x": x, "lens": lens}
    for name, kw in variants.items():
        m = feats.FilterbankFeatures(**kw)
        m.eval()
        y, yl = m.forward(torch.from_numpy(x.copy()), torch.from_numpy(lens))
        out[f"{name}__feat"] = y.numpy()
        out[f"{name}__len"] = yl.numpy()
        print(name, y.shape, yl.tolist())
    np.savez_compressed(f"{HERE}/fbank_ref.npz", **out)

    # ---------------------------------------------------------------- spec / pyin (oracle)
    man = synth.corpus_manifest("C1", n_utts=3)
    out = {}
    fb = melfb.mel_filterbank(22050, 1024, 80, 0.0, 8000.0)
    for i, u in enumerate(man):
        n = min(u.n_samples, 22050 * 2)
        y = synth.synth_utterance(1234, u.utt_id, n, 22050, u.speaker)
        lm, en = ospec.log_mel_energy(y, fb=fb)
        f0, vf, vp = opyin.pyin(y, 65.40639132514966, 2093.004522404789, sr=22050,
                                frame_length=1024, fill_na=0.0, dense_viterbi=True)
        out[f"audio{i}"] = y
        out[f"logmel{i}"] = lm
        out[f"energy{i}"] = en
        out[f"f0_{i}"] = f0
        out[f"vflag{i}"] = vf
        out[f"vprob{i}"] = vp
    out["fb"] = fb
    np.savez_compressed(f"{HERE}/supdata_oracle.npz", **out)
    print("supdata_oracle.npz written")


if __name__ == "__main__":
    main()
