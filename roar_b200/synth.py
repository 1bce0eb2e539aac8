"""Seeded synthetic speech-like corpora.

There is no dataset in the build or on the GPU box, so every test, the smoke run and
``bench.py`` use waveforms that are a pure function of ``(corpus_seed, utt_id)``:

* ``synth_utterance`` (NumPy, float64 maths -> float32) -- used by tests/goldens; identical
  on every machine.
* ``synth_corpus_device`` (torch, on the GPU) -- used by ``bench.py`` to build the
  24 h / 1000 h manifests directly in HBM; same signal model, vectorised.

Signal model: glottal-like harmonic source ``sum_k a_k sin(k*phi(t))`` with ``a_k ~ 1/k`` and
a 2-formant tilt, f0 = speaker mean * 2^(vibrato + slow drift), syllable-like gating
(voiced 80-400 ms, unvoiced 30-150 ms of noise or silence, ~65 % voiced), white noise at
-50 dBFS, peak 0.3-0.9, occasional octave jumps and clipped segments as stressors.
``text_len = 2 + round(dur * U(11, 17))`` mirrors ``pad_with_space`` tokenisation
(reference ``tts_tokenizers.py:157,259``).
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

# manifests of BASELINE.json "configs" 
CORPORA = {
    "C1": dict(n_utts=100, sr=22050, dur=("uniform", 1.0, 10.0), seed=1234, n_speakers=4),
    "C2": dict(n_utts=13100, sr=22050, dur=("normal", 6.57, 2.2, 1.1, 10.1), seed=2, n_speakers=1),
    "C3": dict(n_utts=720000, sr=22050, dur=("lognormal", 4.5, 0.5, 0.5, 20.0), seed=3, n_speakers=400),
    "C4": dict(n_utts=20000, sr=44100, dur=("uniform", 10.0, 30.0), seed=4, n_speakers=10),
    "C5": dict(n_utts=256, sr=16000, dur=("uniform", 2.0, 16.7), seed=5, n_speakers=32),
}


@dataclass
class Utterance:
    utt_id: int
    n_samples: int
    text_len: int
    speaker: int
    duration: float


def corpus_manifest(name: str, n_utts: Optional[int] = None) -> List[Utterance]:
    """Durations / text lengths / speakers of corpus ``name`` (no audio)."""
    spec = CORPORA[name]
    n = spec["n_utts"] if n_utts is None else n_utts
    sr = spec["sr"]
    rng = np.random.default_rng([spec["seed"], 0xD0])
    kind = spec["dur"][0]
    if kind == "uniform":
        dur = rng.uniform(spec["dur"][1], spec["dur"][2], size=n)
    elif kind == "normal":
        dur = np.clip(rng.normal(spec["dur"][1], spec["dur"][2], size=n), spec["dur"][3], spec["dur"][4])
    else:
        dur = np.clip(np.exp(rng.normal(np.log(spec["dur"][1]), spec["dur"][2], size=n)),
                      spec["dur"][3], spec["dur"][4])
    rate = rng.uniform(11.0, 17.0, size=n)
    spk = rng.integers(0, spec["n_speakers"], size=n)
    out = []
    for i in range(n):
        ns = int(round(dur[i] * sr))
        out.append(Utterance(i, ns, 2 + int(round(dur[i] * rate[i])), int(spk[i]), ns / sr))
    return out


def speaker_mean_f0(seed: int, speaker: int) -> float:
    rng = np.random.default_rng([seed, 0x5B, speaker])
    if rng.random() < 0.5:
        return float(np.exp(rng.uniform(np.log(90.0), np.log(150.0))))
    return float(np.exp(rng.uniform(np.log(160.0), np.log(280.0))))


def synth_utterance(seed: int, utt_id: int, n_samples: int, sr: int, speaker: int = 0) -> np.ndarray:
    """float32 ``[n_samples]`` in [-1, 1]; pure function of its arguments."""
    rng = np.random.default_rng([seed, 0xA0, utt_id])
    n = n_samples
    t = np.arange(n, dtype=np.float64) / sr
    fbar = speaker_mean_f0(seed, speaker)
    # f0 contour in semitones: vibrato + slow drift (sum of 3 slow sinusoids, +-3 st)
    st = 0.3 * np.sin(2 * np.pi * 5.0 * t + rng.uniform(0, 2 * np.pi))
    for _ in range(3):
        st += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * rng.uniform(0.1, 0.6) * t + rng.uniform(0, 2 * np.pi))
    # syllable gating
    gate = np.zeros(n)
    noise_gate = np.zeros(n)
    octave = np.zeros(n)
    pos = int(rng.uniform(0.0, 0.1) * sr)
    ramp = max(2, int(0.005 * sr))
    while pos < n:
        v = int(rng.uniform(0.08, 0.40) * sr)
        e = min(n, pos + v)
        seg = np.ones(e - pos)
        r = min(ramp, (e - pos) // 2)
        if r > 0:
            w = 0.5 * (1 - np.cos(np.pi * np.arange(r) / r))
            seg[:r] = w
            seg[-r:] = w[::-1]
        gate[pos:e] = seg * rng.uniform(0.5, 1.0)
        if rng.random() < 0.04:
            octave[pos:e] = 12.0 if rng.random() < 0.5 else -12.0
        pos = e
        u = int(rng.uniform(0.03, 0.15) * sr)
        e = min(n, pos + u)
        if rng.random() < 0.5:
            noise_gate[pos:e] = rng.uniform(0.02, 0.15)
        pos = e
    f0 = fbar * 2.0 ** ((st + octave) / 12.0)
    phase = 2 * np.pi * np.cumsum(f0) / sr
    # harmonic amplitudes: 1/k with a 2-formant tilt
    f1, f2 = rng.uniform(400, 900), rng.uniform(1200, 2600)
    kmax = int(min(40, np.floor(0.45 * sr / (fbar * 2.0 ** (4.0 / 12.0)))))
    y = np.zeros(n)
    for k in range(1, max(2, kmax + 1)):
        fk = k * fbar
        amp = (1.0 / k) * (1.0 + 2.0 * np.exp(-0.5 * ((fk - f1) / 150.0) ** 2)
                           + 1.5 * np.exp(-0.5 * ((fk - f2) / 250.0) ** 2))
        y += amp * np.sin(k * phase + rng.uniform(0, 2 * np.pi))
    y *= gate
    # unvoiced: first-order low-passed noise
    nz = rng.standard_normal(n)
    nz[1:] = 0.6 * nz[1:] + 0.4 * nz[:-1]
    peak = np.max(np.abs(y))
    if peak > 0:
        y /= peak
    y += noise_gate * nz
    peak = np.max(np.abs(y))
    if peak > 0:
        y *= rng.uniform(0.3, 0.9) / peak
    y += (10 ** (-50 / 20)) * rng.standard_normal(n)
    if rng.random() < 0.05:  # a clipped stretch
        a = int(rng.uniform(0, 0.8) * n)
        b = min(n, a + int(0.3 * sr))
        y[a:b] = np.clip(y[a:b] * 4.0, -0.95, 0.95)
    return np.clip(y, -1.0, 1.0).astype(np.float32)


def synth_corpus(name: str, n_utts: Optional[int] = None):
    """-> (manifest, list of float32 waveforms) on the CPU."""
    spec = CORPORA[name]
    man = corpus_manifest(name, n_utts)
    return man, [synth_utterance(spec["seed"], u.utt_id, u.n_samples, spec["sr"], u.speaker) for u in man]


def synth_corpus_device(name: str, device, n_utts: Optional[int] = None, align: int = 4,
                        batch: int = 256):
    """Build the packed ragged corpus directly on ``device`` with torch.

    -> (manifest, audio float32 ``[total]``, sample_off int64 ``[n]`` (multiples of ``align``),
    sample_len int32 ``[n]``).  Same signal model as ``synth_utterance`` (vectorised, so not
    sample-identical to it); deterministic for a given torch build and device type.
    """
    import torch

    spec = CORPORA[name]
    sr = spec["sr"]
    man = corpus_manifest(name, n_utts)
    n = len(man)
    lens = np.array([u.n_samples for u in man], dtype=np.int64)
    padded = (lens + align - 1) // align * align
    offs = np.zeros(n, dtype=np.int64)
    offs[1:] = np.cumsum(padded)[:-1]
    total = int(padded.sum())
    audio = torch.zeros(total, dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(spec["seed"] * 7919 + 17)
    fbar_spk = np.array([speaker_mean_f0(spec["seed"], s) for s in range(spec["n_speakers"])])
    order = np.argsort(-lens, kind="stable")
    for s in range(0, n, batch):
        idx = order[s : s + batch]
        B = len(idx)
        Lmax = int(lens[idx].max())
        L = torch.as_tensor(lens[idx], device=device)
        t = torch.arange(Lmax, device=device, dtype=torch.float64)[None, :] / sr
        fbar = torch.as_tensor(fbar_spk[[man[i].speaker for i in idx]], device=device)[:, None]

        def U(lo, hi, *shape):
            return lo + (hi - lo) * torch.rand(*shape, generator=g, device=device, dtype=torch.float64)

        st = 0.3 * torch.sin(2 * np.pi * 5.0 * t + U(0, 2 * np.pi, B, 1))
        for _ in range(3):
            st = st + U(0.3, 1.0, B, 1) * torch.sin(2 * np.pi * U(0.1, 0.6, B, 1) * t + U(0, 2 * np.pi, B, 1))
        # gating: syllable clock of random period per utterance, duty ~65 %, smoothed edges
        per = U(0.18, 0.55, B, 1)
        ph0 = U(0, 1, B, 1)
        duty = U(0.55, 0.8, B, 1)
        cyc = (t / per + ph0)
        frac = cyc - torch.floor(cyc)
        edge = 0.005 / per
        gate = torch.clamp(frac / edge, 0, 1) * torch.clamp((duty - frac) / edge, 0, 1)
        syl = torch.floor(cyc)
        # per-syllable pseudo-random level / noise choice (hash of syllable index)
        h = torch.frac(torch.sin(syl * 12.9898 + ph0 * 78.233) * 43758.5453).abs()
        gate = gate * (0.5 + 0.5 * h)
        octave = torch.where(h < 0.04, torch.where(h < 0.02, 12.0, -12.0), 0.0)
        noise_gate = torch.where((frac > duty) & (h > 0.5), 0.02 + 0.13 * h, 0.0)
        f0 = fbar * torch.pow(2.0, (st + octave) / 12.0)
        phase = 2 * np.pi * torch.cumsum(f0, dim=1) / sr
        f1 = U(400, 900, B, 1)
        f2 = U(1200, 2600, B, 1)
        y = torch.zeros(B, Lmax, device=device, dtype=torch.float32)
        kmax = int(min(40, np.floor(0.45 * sr / (float(fbar.max()) * 2.0 ** (4.0 / 12.0)))))
        for k in range(1, max(2, kmax + 1)):
            fk = k * fbar
            amp = (1.0 / k) * (1.0 + 2.0 * torch.exp(-0.5 * ((fk - f1) / 150.0) ** 2)
                               + 1.5 * torch.exp(-0.5 * ((fk - f2) / 250.0) ** 2))
            y += (amp * torch.sin(k * phase + U(0, 2 * np.pi, B, 1))).float()
        y = y * gate.float()
        valid = (torch.arange(Lmax, device=device)[None, :] < L[:, None])
        y = y * valid
        peak = y.abs().amax(dim=1, keepdim=True).clamp_min(1e-9)
        y = y / peak
        nz = torch.randn(B, Lmax, generator=g, device=device, dtype=torch.float32)
        nz[:, 1:] = 0.6 * nz[:, 1:] + 0.4 * nz[:, :-1]
        y = y + noise_gate.float() * nz
        peak = y.abs().amax(dim=1, keepdim=True).clamp_min(1e-9)
        y = y * (U(0.3, 0.9, B, 1).float() / peak)
        y = y + (10 ** (-50 / 20)) * torch.randn(B, Lmax, generator=g, device=device, dtype=torch.float32)
        y = torch.clamp(y, -1.0, 1.0) * valid
        for b, i in enumerate(idx):
            audio[offs[i] : offs[i] + lens[i]] = y[b, : lens[i]]
    return (man, audio, torch.as_tensor(offs, device=device),
            torch.as_tensor(lens.astype(np.int32), device=device))


# ------------------------------------------------------------------------------------------------------
# Stateless device synthesis: utterance i of a corpus is a pure function of (corpus seed, utt_id), whatever
# subset / batch it is generated in -- so every rank of a sharded run (bench.py --workload C3, SURVEY.md
# section 8e) regenerates exactly its own utterances of ONE seeded manifest and the all-reduced pitch
# statistics do not depend on the number of GPUs.  Per-utterance scalars come from a NumPy generator seeded by
# (seed, 0xB0, utt_id); per-sample noise from a counter-based integer hash of (utt_id, stream, sample index).
def _mix32(x):
    """32-bit integer finaliser on int64 tensors holding uint32 values."""
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    return x ^ (x >> 16)


def _hash_normal(torch, utt_key, stream: int, L: int, device):
    """Standard-normal float32 ``[B, L]``: Box-Muller over two hashed uniforms per sample."""
    t = torch.arange(L, device=device, dtype=torch.int64)[None, :]
    k = _mix32((utt_key[:, None] + stream * 0x9E3779B1) & 0xFFFFFFFF)
    h1 = _mix32(k ^ _mix32(2 * t))
    h2 = _mix32(k ^ _mix32(2 * t + 1))
    u1 = (h1.to(torch.float32) + 0.5) * (1.0 / 4294967296.0)
    u2 = (h2.to(torch.float32) + 0.5) * (1.0 / 4294967296.0)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)


def utterance_params(seed: int, utt_id: int) -> np.ndarray:
    """The 64 per-utterance uniforms of the stateless generator."""
    return np.random.default_rng([seed, 0xB0, utt_id]).random(64)


def synth_corpus_device_stateless(name: str, device, indices=None, n_utts: Optional[int] = None, align: int = 4,
                                  batch: int = 256, pcm16: bool = True):
    """Packed ragged corpus on ``device`` for utterances ``indices`` (default: all) of manifest ``name``.

    -> (manifest subset, audio float32 ``[total]``, sample_off int64 ``[n]``, sample_len int32 ``[n]``).  With
    ``pcm16`` the samples are the exact ``x / 2**15`` values of a 16-bit quantisation (what a wav corpus holds:
    ``audio * 32768`` is then integral and can travel host -> device as int16).  Same signal model as
    ``synth_utterance``."""
    import torch

    spec = CORPORA[name]
    sr = spec["sr"]
    full = corpus_manifest(name, n_utts)
    man = full if indices is None else [full[int(i)] for i in indices]
    n = len(man)
    lens = np.array([u.n_samples for u in man], dtype=np.int64)
    padded = (lens + align - 1) // align * align
    offs = np.zeros(n, dtype=np.int64)
    if n > 1:
        offs[1:] = np.cumsum(padded)[:-1]
    total = int(padded.sum())
    audio = torch.zeros(total, dtype=torch.float32, device=device)
    fbar_spk = {}
    order = np.argsort(-lens, kind="stable")
    two_pi = 2 * np.pi
    for s in range(0, n, batch):
        idx = order[s: s + batch]
        B = len(idx)
        Lmax = int(lens[idx].max())
        L = torch.as_tensor(lens[idx], device=device)
        P = np.stack([utterance_params(spec["seed"], man[i].utt_id) for i in idx])          # [B, 64]
        for i in idx:
            sp = man[i].speaker
            if sp not in fbar_spk:
                fbar_spk[sp] = speaker_mean_f0(spec["seed"], sp)
        fbar_h = np.array([fbar_spk[man[i].speaker] for i in idx])
        col = lambda j, lo=0.0, hi=1.0: torch.as_tensor(lo + (hi - lo) * P[:, j], device=device)[:, None]  # noqa: E731
        t = torch.arange(Lmax, device=device, dtype=torch.float64)[None, :] / sr
        fbar = torch.as_tensor(fbar_h, device=device)[:, None]
        st = 0.3 * torch.sin(two_pi * 5.0 * t + col(0, 0, two_pi))
        for q in range(3):
            st = st + col(1 + 3 * q, 0.3, 1.0) * torch.sin(two_pi * col(2 + 3 * q, 0.1, 0.6) * t + col(3 + 3 * q, 0, two_pi))
        per, ph0, duty = col(10, 0.18, 0.55), col(11), col(12, 0.55, 0.8)
        cyc = t / per + ph0
        frac = cyc - torch.floor(cyc)
        edge = 0.005 / per
        gate = torch.clamp(frac / edge, 0, 1) * torch.clamp((duty - frac) / edge, 0, 1)
        syl = torch.floor(cyc)
        h = torch.frac(torch.sin(syl * 12.9898 + ph0 * 78.233) * 43758.5453).abs()
        gate = gate * (0.5 + 0.5 * h)
        octave = torch.where(h < 0.04, torch.where(h < 0.02, 12.0, -12.0), 0.0)
        noise_gate = torch.where((frac > duty) & (h > 0.5), 0.02 + 0.13 * h, 0.0)
        f0 = fbar * torch.pow(2.0, (st + octave) / 12.0)
        phase = two_pi * torch.cumsum(f0, dim=1) / sr
        f1, f2 = col(13, 400, 900), col(14, 1200, 2600)
        kmax_h = np.minimum(40, np.floor(0.45 * sr / (fbar_h * 2.0 ** (4.0 / 12.0)))).astype(np.int64)
        kmax = torch.as_tensor(kmax_h, device=device)[:, None]
        y = torch.zeros(B, Lmax, device=device, dtype=torch.float32)
        for k in range(1, max(2, int(kmax_h.max()) + 1)):
            fk = k * fbar
            amp = (1.0 / k) * (1.0 + 2.0 * torch.exp(-0.5 * ((fk - f1) / 150.0) ** 2)
                               + 1.5 * torch.exp(-0.5 * ((fk - f2) / 250.0) ** 2))
            if k > 1:      # harmonic k only below 0.45 * sr for this utterance's speaker
                amp = torch.where(kmax >= k, amp, torch.zeros_like(amp))
            y += (amp * torch.sin(k * phase + col(15 + (k % 45), 0, two_pi))).float()
        y = y * gate.float()
        valid = (torch.arange(Lmax, device=device)[None, :] < L[:, None])
        y = y * valid
        peak = y.abs().amax(dim=1, keepdim=True).clamp_min(1e-9)
        y = y / peak
        key = torch.as_tensor([(man[i].utt_id * 2654435761 + spec["seed"] * 40503) & 0xFFFFFFFF for i in idx],
                              device=device, dtype=torch.int64)
        nz = _hash_normal(torch, key, 1, Lmax, device)
        nz[:, 1:] = 0.6 * nz[:, 1:] + 0.4 * nz[:, :-1]
        y = y + noise_gate.float() * nz
        peak = (y * valid).abs().amax(dim=1, keepdim=True).clamp_min(1e-9)
        y = y * (col(61, 0.3, 0.9).float() / peak)
        y = y + (10 ** (-50 / 20)) * _hash_normal(torch, key, 2, Lmax, device)
        y = torch.clamp(y, -1.0, 1.0) * valid
        if pcm16:
            y = torch.clamp(torch.round(y * 32767.0), -32768, 32767) * (1.0 / 32768.0)
        for b, i in enumerate(idx):
            audio[offs[i]: offs[i] + lens[i]] = y[b, : lens[i]]
    return (man, audio, torch.as_tensor(offs, device=device),
            torch.as_tensor(lens.astype(np.int32), device=device))
