"""CPU tests that pin the oracle (oracle/) to the committed golden vectors and to
independent cross-checks.  No GPU, no /root/reference at run time."""
import os

import numpy as np
import pytest
import scipy.stats
import torch

from oracle import fbank, melfb, prior, pyin, spec, stats
from roar_b200 import synth

FMIN, FMAX = 65.40639132514966, 2093.004522404789


# ------------------------------------------------------------------ mel filterbank
@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", [
    (22050, 1024, 80, 0.0, 8000.0), (44100, 2048, 80, 0.0, None),
    (22050, 2048, 80, 0.0, 8000.0), (16000, 512, 80, 0.0, None)])
def test_melfb_matches_torchaudio(sr, n_fft, n_mels, fmin, fmax):
    import torchaudio
    fb = melfb.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    ta = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fmax or sr / 2, n_mels, sr,
                                               norm="slaney", mel_scale="slaney").T.numpy()
    assert fb.dtype == np.float32 and fb.shape == (n_mels, n_fft // 2 + 1)
    assert np.abs(fb - ta).max() <= 5e-6 * fb.max()


def test_melfb_sparsity_survey_numbers():
    fb = melfb.mel_filterbank(22050, 1024, 80, 0.0, 8000.0)
    assert np.count_nonzero(fb) == 727
    assert np.nonzero(fb.any(axis=0))[0].max() == 371
    fb = melfb.mel_filterbank(16000, 512, 80, 0.0, None)
    assert np.count_nonzero(fb) == 500


# ------------------------------------------------------------------ spec / log-mel / energy
def test_spec_matches_float64_framing():
    y = synth.synth_utterance(1234, 7, 30000, 22050, 1)
    a = spec.get_spec(y).numpy()
    b = spec.spec_f64(y)
    assert a.shape == b.shape == (513, 1 + 30000 // 256)
    assert np.abs(a - b).max() <= 3e-6 * b.max()


def test_spec_win_shorter_than_nfft():
    y = synth.synth_utterance(5, 1, 16000, 16000, 0)
    a = spec.get_spec(y, 512, 160, 400).numpy()
    b = spec.spec_f64(y, 512, 160, 400)
    assert np.abs(a - b).max() <= 3e-6 * b.max()


def test_supdata_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "supdata_oracle.npz"))
    fb = melfb.mel_filterbank(22050, 1024, 80, 0.0, 8000.0)
    assert np.array_equal(fb, g["fb"])
    for i in range(3):
        y = g[f"audio{i}"]
        u = synth.corpus_manifest("C1", 3)[i]
        assert np.array_equal(y, synth.synth_utterance(1234, u.utt_id, len(y), 22050, u.speaker))
        lm, en = spec.log_mel_energy(y, fb=fb)
        assert lm.shape == (1, 80, 1 + len(y) // 256)
        np.testing.assert_allclose(lm, g[f"logmel{i}"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(en, g[f"energy{i}"], rtol=1e-5)


# ------------------------------------------------------------------ prior
def test_prior_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "prior_ref.npz"))
    for k in g.files:
        kind, a, b = k.split("_")
        a, b = int(a), int(b)
        if kind == "p":
            got = prior.beta_binomial_prior_distribution(a, b)
            assert got.shape == (b, a) and got.dtype == np.float32
        else:
            got = prior.BetaBinomialInterpolator()(a, b)
        assert np.array_equal(got, g[k]), k


def test_prior_float32_noise_vs_float64():
    p32 = prior.beta_binomial_prior_distribution(100, 560)
    p64 = prior.prior_f64(100, 560)
    big = p64 > 1e-6
    assert np.abs(p32[big] / p64[big] - 1).max() < 5e-3
    assert abs(p64.sum(axis=1) - 1).max() < 1e-9
    assert (p32.argmax(1) == p64.argmax(1)).mean() > 0.995


# ------------------------------------------------------------------ fbank
@pytest.mark.parametrize("name,kw", [
    ("asr_default", dict(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=80, n_fft=512)),
    ("tts_fastpitch", dict(sample_rate=22050, n_window_size=1024, n_window_stride=256, nfilt=80,
                           n_fft=1024, lowfreq=0, highfreq=8000, normalize=None, preemph=None,
                           log_zero_guard_type="add", log_zero_guard_value=1.0, mag_power=1.0,
                           pad_to=1, pad_value=0.0)),
    ("clamp_allfeat_exactpad", dict(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=64,
                                    n_fft=512, exact_pad=True, normalize="all_features",
                                    log_zero_guard_type="clamp", log_zero_guard_value="tiny",
                                    pad_to=8, pad_value=-1.0, mel_norm=None))])
def test_fbank_oracle_matches_reference_golden(golden_dir, name, kw):
    g = np.load(os.path.join(golden_dir, "fbank_ref.npz"))
    feat, ln = fbank.FilterbankFeaturesOracle(**kw).forward(g["x"], g["lens"])
    assert np.array_equal(ln, g[f"{name}__len"])
    assert np.array_equal(feat, g[f"{name}__feat"])


# ------------------------------------------------------------------ pyin
def test_pyin_closed_forms():
    th, bp = pyin.beta_threshold_prior()
    x = th
    cdf = 1 - (1 - x) ** 19 - 19 * x * (1 - x) ** 18
    assert np.abs(np.diff(cdf) - bp).max() < 1e-14 and abs(bp.sum() - 1) < 1e-12
    k = np.arange(5)
    lam, N = 2.0, 5
    pmf = (1 - np.exp(-lam)) * np.exp(-lam * k) / (1 - np.exp(-lam * N))
    assert np.abs(scipy.stats.boltzmann.pmf(k, lam, N) - pmf).max() < 1e-15
    assert pyin.n_pitch_bins_for(FMIN, FMAX) == (601, 10)
    assert pyin.periods(22050, FMIN, FMAX, 1024, 512) == (10, 338)
    assert pyin.periods(44100, FMIN, FMAX, 2048, 1024) == (21, 675)
    tr, p0, tw = pyin.hmm_tables(22050, 256, 601, 10)
    assert tw == 51 and tr.shape == (1202, 1202) and abs(tr.sum(1) - 1).max() < 1e-12
    assert np.count_nonzero(tr[300, :601]) == 51 and tr[0, 26] == 0 and tr[0, 25] > 0


def test_pyin_banded_viterbi_equals_dense_and_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "supdata_oracle.npz"))
    for i in range(2):
        y = g[f"audio{i}"]
        f0, vf, vp = pyin.pyin(y, FMIN, FMAX, sr=22050, frame_length=1024, fill_na=0.0)
        assert len(f0) == 1 + len(y) // 256
        assert np.array_equal(f0, g[f"f0_{i}"])       # golden was made with the dense DP
        assert np.array_equal(vf, g[f"vflag{i}"])
        np.testing.assert_allclose(vp, g[f"vprob{i}"], rtol=0, atol=1e-12)


def test_pyin_tracks_known_pitch():
    sr = 22050
    t = np.arange(sr) / sr
    y = (0.5 * np.sin(2 * np.pi * 220.0 * t) + 0.2 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    f0, vf, vp = pyin.pyin(y, FMIN, FMAX, sr=sr, frame_length=1024, fill_na=0.0)
    mid = slice(5, -5)
    assert vf[mid].all()
    assert np.abs(1200 * np.log2(f0[mid] / 220.0)).max() < 10.0
    z = np.zeros(8000, dtype=np.float32)
    f0, vf, vp = pyin.pyin(z, FMIN, FMAX, sr=sr, frame_length=1024, fill_na=0.0)
    assert not vf.any() and (f0 == 0).all() and (vp == 0).all()


# ------------------------------------------------------------------ stats
def test_pitch_stats():
    rng = np.random.default_rng(0)
    ps = [np.where(rng.random(500) < 0.6, rng.uniform(80, 300, 500), 0).astype(np.float32) for _ in range(20)]
    a = stats.pitch_stats(ps)
    b = stats.pitch_stats_f64(ps)
    assert abs(a["mean"] - b["mean"]) < 1e-3 and abs(a["std"] - b["std"]) < 1e-3
    assert a["min"] == b["min"] and a["max"] == b["max"]
    p = stats.normalize_pitch(ps[0], a["mean"], a["std"])
    assert (p[ps[0] == 0] == 0).all()


def test_float32_gammaln_of_integers_is_correctly_rounded():
    """What the prior kernel relies on: torch's float32 gammaln of an integer argument equals the float64
    value rounded once (so a float64 log-factorial table reproduces the reference's float32 terms)."""
    import torch
    from scipy.special import gammaln
    x = np.arange(1, 20000, dtype=np.float64)
    t = torch.special.gammaln(torch.tensor(x, dtype=torch.float32)).numpy()
    eq = t == gammaln(x).astype(np.float32)
    assert eq[:10000].all()            # every argument a 100 s utterance can produce
    assert eq.mean() > 0.9998          # two 1-ulp exceptions below 20 000
