"""The kernels' per-thread code (roar_b200/csrc/*.cuh), executed on the CPU by tests/hostemu, against
the oracle.  This pins kernel LOGIC without a GPU; the `-m gpu` tests pin the real launches."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hostemu  # noqa: E402
from oracle import fbank as ofbank, prior as oprior, pyin as opyin, spec as ospec  # noqa: E402
from roar_b200 import synth  # noqa: E402
from roar_b200.config import SupConfig  # noqa: E402

FMIN, FMAX = 65.40639132514966, 2093.004522404789


def _wav(i, n, sr=22050, seed=1234):
    u = synth.corpus_manifest("C1", i + 1)[i]
    return synth.synth_utterance(seed, u.utt_id, n, sr, u.speaker)


@pytest.mark.parametrize("n", [513, 4096, 4097, 30000])
def test_emu_logmel_energy(n):
    cfg = SupConfig(highfreq=8000.0)
    y = _wav(2, n)
    lm, en = hostemu.logmel_energy(cfg, y)
    olm, oen = ospec.log_mel_energy(y)
    assert (np.abs(lm - olm[0]) / np.maximum(1, np.abs(olm[0]))).max() <= 1e-4
    np.testing.assert_allclose(en, oen, rtol=1e-4)


@pytest.mark.parametrize("kw", [dict(sample_rate=44100, n_fft=2048, hop_length=512, highfreq=None),
                                dict(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, highfreq=None, pyin_frame_length=512),
                                dict(sample_rate=22050, n_fft=256, hop_length=64, n_mels=40, highfreq=None)])
def test_emu_logmel_other_geometries(kw):
    cfg = SupConfig(**kw)
    y = _wav(1, 20000, sr=kw["sample_rate"])
    lm, en = hostemu.logmel_energy(cfg, y)
    olm, oen = ospec.log_mel_energy(y, sr=cfg.sample_rate, n_fft=cfg.n_fft, hop_length=cfg.hop,
                                    win_length=cfg.win, n_mels=cfg.n_mels, fmin=0.0, fmax=cfg.highfreq)
    assert lm.shape == olm[0].shape
    assert (np.abs(lm - olm[0]) / np.maximum(1, np.abs(olm[0]))).max() <= 1e-4
    np.testing.assert_allclose(en, oen, rtol=1e-4)


def test_emu_fbank_front_matches_oracle_pre_normalisation():
    """FilterbankFeatures front end (pre-emphasis, power spectrum, log add) through the K1 code."""
    y = _wav(0, 16000, sr=16000, seed=5)
    cfg = SupConfig(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, highfreq=None, spec_floor=0.0,
                    mag_power=2.0, log_mode="add", log_guard=2 ** -24, preemph=0.97, pyin_frame_length=512)
    lm, _ = hostemu.logmel_energy(cfg, y)
    ref, _ = ofbank.FilterbankFeaturesOracle(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=80,
                                             n_fft=512, normalize=None, pad_to=0).forward(y[None, :], [len(y)])
    assert (np.abs(lm - ref[0]) / np.maximum(1, np.abs(ref[0]))).max() <= 1e-4


def test_emu_fbank_exact_pad_preemph_order():
    """exact_pad pads first and pre-emphasises the padded signal (features.py:387-400)."""
    from roar_b200.config import FLOAT32_TINY
    y = _wav(1, 12000, sr=16000, seed=5)
    cfg = SupConfig(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=64, highfreq=None,
                    mel_norm=None, spec_floor=0.0, mag_power=2.0, log_mode="clamp", log_guard=FLOAT32_TINY,
                    exact_pad=True, preemph=0.97, pyin_frame_length=512)
    lm, _ = hostemu.logmel_energy(cfg, y)
    ref, rl = ofbank.FilterbankFeaturesOracle(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=64,
                                              n_fft=512, exact_pad=True, normalize=None,
                                              log_zero_guard_type="clamp", log_zero_guard_value="tiny", pad_to=0,
                                              mel_norm=None).forward(y[None, :], [len(y)])
    assert lm.shape == ref[0].shape == (64, 12000 // 160)
    assert (np.abs(lm - ref[0]) / np.maximum(1, np.abs(ref[0]))).max() <= 1e-4


def cm_lags(cfg):
    """number of CMND lags of a configuration (librosa: min_period .. max_period)"""
    frame = cfg.pyin_frame
    win = cfg.pyin_win_length or frame // 2
    min_period = int(np.floor(cfg.sample_rate / cfg.pitch_fmax))
    max_period = min(int(np.ceil(cfg.sample_rate / cfg.pitch_fmin)), frame - win - 1)
    return max_period - min_period + 1


@pytest.mark.parametrize("i,n", [(0, 44100), (1, 30000), (2, 9000)])
def test_emu_pyin_equals_oracle(i, n):
    cfg = SupConfig(highfreq=8000.0)
    y = _wav(i, n)
    f0, vf, vp, cm, st = hostemu.pyin(cfg, y, 329)
    of0, ovf, ovp, info = opyin.pyin(y, FMIN, FMAX, sr=22050, frame_length=1024, fill_na=0.0, return_internals=True)
    assert np.abs(cm - info["yin_frames"].T).max() < 1e-9
    assert np.array_equal(st, info["states"])
    assert np.array_equal(vf.astype(bool), ovf)
    assert np.array_equal(f0, of0.astype(np.float32))
    assert np.abs(vp - ovp).max() < 1e-6


def test_emu_pyin_44k_and_edge_inputs():
    cfg = SupConfig(sample_rate=44100, n_fft=2048, hop_length=512, highfreq=None)
    y = _wav(1, 40000, sr=44100)
    f0, vf, vp, cm, st = hostemu.pyin(cfg, y, 655)
    of0, ovf, ovp, info = opyin.pyin(y, FMIN, FMAX, sr=44100, frame_length=2048, fill_na=0.0, return_internals=True)
    assert np.array_equal(st, info["states"]) and np.abs(vp - ovp).max() < 1e-6
    cfg = SupConfig(highfreq=8000.0)
    for y in (np.zeros(3000, np.float32), np.ones(1, np.float32), _wav(0, 300)):
        f0, vf, vp, cm, st = hostemu.pyin(cfg, y, 329)
        of0, ovf, ovp = opyin.pyin(y, FMIN, FMAX, sr=22050, frame_length=1024, fill_na=0.0)
        assert np.array_equal(vf.astype(bool), ovf) and np.array_equal(f0, of0.astype(np.float32))


def test_emu_viterbi_fast_path_equals_generic(monkeypatch):
    """The pruned Viterbi (live lists, dominance skipping) must decode the same state sequence as the
    kernel that scans every in-band source -- including octave jumps, clipping and pure noise."""
    cfg = SupConfig(highfreq=8000.0)
    rng = np.random.default_rng(7)
    t = np.arange(22050) / 22050.0
    jump = np.where(t < 0.5, np.sin(2 * np.pi * 110 * t), np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    wavs = [_wav(3, 30000), (0.05 * rng.standard_normal(20000)).astype(np.float32), jump,
            np.clip(3.0 * _wav(4, 20000), -0.9, 0.9).astype(np.float32)]
    for y in wavs:
        monkeypatch.delenv("ROAR_SUP_VITERBI", raising=False)
        fast = hostemu.pyin(cfg, y, 329)
        monkeypatch.setenv("ROAR_SUP_VITERBI", "generic")
        gen = hostemu.pyin(cfg, y, 329)
        assert np.array_equal(fast[4], gen[4])
        assert np.array_equal(fast[0], gen[0]) and np.array_equal(fast[1], gen[1])


def test_emu_viterbi_round2_rules_are_exact(monkeypatch):
    """Dead-on-arrival dense steps, the twin filter of the voiced live list and the flat-segment shortcut
    (k_viterbi.cuh rules 7, 8, 10) only
    skip work that cannot reach the decoded path: switching either off, or both, or taking the every-source
    recursion gives the same states.  PCM-quantised audio with clipped voiced probability (many dense steps)."""
    cfg = SupConfig(highfreq=8000.0)
    man = synth.corpus_manifest("C2", 3)
    for u in man[:2]:
        y = synth.synth_utterance(synth.CORPORA["C2"]["seed"], u.utt_id, min(u.n_samples, 40000), 22050, u.speaker)
        y = (np.clip(np.round(y * 32767), -32768, 32767) / 32768.0).astype(np.float32)
        for k in ("ROAR_SUP_VITERBI", "ROAR_EMU_NO_DOA", "ROAR_EMU_NO_TWIN", "ROAR_EMU_NO_FLAT"):
            monkeypatch.delenv(k, raising=False)
        base = hostemu.pyin(cfg, y, 329)
        assert (base[2] == 1.0).any()           # frames whose voiced probability clipped to 1: dense steps exist
        for envs in (("ROAR_EMU_NO_DOA",), ("ROAR_EMU_NO_TWIN",), ("ROAR_EMU_NO_FLAT",),
                     ("ROAR_EMU_NO_DOA", "ROAR_EMU_NO_TWIN", "ROAR_EMU_NO_FLAT")):
            for k in envs:
                monkeypatch.setenv(k, "1")
            alt = hostemu.pyin(cfg, y, 329)
            for k in envs:
                monkeypatch.delenv(k)
            assert np.array_equal(base[4], alt[4]), envs
        monkeypatch.setenv("ROAR_SUP_VITERBI", "generic")
        gen = hostemu.pyin(cfg, y, 329)
        monkeypatch.delenv("ROAR_SUP_VITERBI")
        assert np.array_equal(base[4], gen[4])


def test_uniform_interior_row_is_exact():
    """Interior transition rows differ by a last-place unit in a few entries (rounding of the row sum);
    the kernel's uniform band scan uses ONE row for all of them once vmax <= uniform_vmax.  Brute force:
    the float64 sums are identical for every interior row, band offset and binade below that bound."""
    for cfg in (SupConfig(highfreq=8000.0), SupConfig(sample_rate=44100, n_fft=2048, hop_length=512)):
        rc, out = hostemu.uniform_row_check(cfg, 64)
        assert rc == 0
        mismatches, differing, bound, compared = out
        assert differing > 0 and compared > 1000      # the premise: rows do differ
        assert mismatches == 0
        assert 1.0 <= bound <= 2.0 ** 14               # valid after a handful of frames


def test_emu_prior(golden_dir):
    g = np.load(os.path.join(golden_dir, "prior_ref.npz"))
    for k in ("p_7_13", "p_100_560", "p_1_4", "p_2_9"):
        _, n, m = k.split("_")
        got = hostemu.prior(int(n), int(m))
        ref64 = oprior.prior_f64(int(n), int(m))
        assert np.abs(got - ref64).max() < 2e-3          # the reference's float32 noise, reproduced
        assert np.abs(got - g[k]).max() < 1e-7           # golden = the reference's own output
        assert np.array_equal(got.argmax(1), g[k].argmax(1))
    got = hostemu.prior(50, 300, 0.5)
    assert np.abs(got - oprior.prior_f64(50, 300, 0.5)).max() < 1e-6


def test_emu_featurizer_mel_and_energy():
    """MelSpectrogramFeaturizer / EnergyFeaturizer arithmetic (features.py:166-302): magnitude mel,
    log(x + 1), mel_norm=None, energy = L2 norm of the log-mel over the mel axis."""
    import torch
    y = _wav(2, 20000)
    cfg = SupConfig(highfreq=8000.0, mel_norm=None, spec_floor=0.0, log_mode="add", log_guard=1.0, energy_mode="features")
    lm, en = hostemu.logmel_energy(cfg, y)
    ref, _ = ofbank.FilterbankFeaturesOracle(sample_rate=22050, n_window_size=1024, n_window_stride=256, nfilt=80,
                                             n_fft=1024, lowfreq=0, highfreq=8000, mag_power=1.0, normalize=None,
                                             log_zero_guard_type="add", log_zero_guard_value=1.0, pad_to=1,
                                             mel_norm=None, preemph=None).forward(y[None, :], [len(y)])
    ref = np.asarray(ref[0])
    assert lm.shape == ref.shape
    assert (np.abs(lm - ref) / np.maximum(1, np.abs(ref))).max() <= 1e-4
    ref_en = torch.linalg.norm(torch.from_numpy(ref), axis=0).numpy()
    np.testing.assert_allclose(en, ref_en, rtol=1e-4)


def test_emu_prior_interpolator():
    """BetaBinomialInterpolator (tts_dataset_utils.py:69-92): rounded-size prior + ndimage.zoom(order=1),
    including scipy's artefact of zeroing a last row/column whose coordinate rounds past the input."""
    it = oprior.BetaBinomialInterpolator()
    zero_cases = 0
    for w, h in [(560, 100), (75, 15), (1, 1), (2, 1), (99, 2), (812, 141), (333, 47), (50, 10), (149, 31)]:
        ref = it(w, h)
        got = hostemu.prior_interp(h, w)
        assert got.shape == ref.shape == (w, h)
        assert np.abs(got - ref).max() < 2e-4
        big = ref > 1e-30          # ignore float32-underflow noise of the reference's gammaln
        assert np.array_equal(got[big] == 0, ref[big] == 0)
        assert np.array_equal((got == 0).all(axis=1), (ref == 0).all(axis=1))
        zero_cases += int((ref == 0).all(axis=1).any() or (ref == 0).all(axis=0).any())
    assert zero_cases >= 1     # (75, 15) has the zeroed last row / column


@pytest.mark.parametrize("kw,okw", [
    (dict(pyin_win_length=600), dict(win_length=600)),                       # W not a multiple of hop: one block per frame
    (dict(pyin_hop_length=128), dict(hop_length=128)),                       # W = 4 * hop: four shared blocks per frame
    (dict(pyin_win_length=384, pyin_hop_length=128), dict(win_length=384, hop_length=128)),
])
def test_emu_pyin_block_geometries(kw, okw):
    """The block decomposition of the autocorrelation for non-default window / hop combinations."""
    cfg = SupConfig(highfreq=8000.0, **kw)
    y = _wav(5, 16000)
    f0, vf, vp, cm, st = hostemu.pyin(cfg, y, cm_lags(cfg))
    of0, ovf, ovp, info = opyin.pyin(y, FMIN, FMAX, sr=22050, frame_length=1024, fill_na=0.0, return_internals=True, **okw)
    assert cm.shape == info["yin_frames"].T.shape
    assert np.abs(cm - info["yin_frames"].T).max() < 1e-9
    assert np.array_equal(st, info["states"]) and np.abs(vp - ovp).max() < 1e-6


def test_emu_trim_equals_oracle():
    """librosa.effects.trim as AudioSegment calls it (segment.py:76-88): leading / trailing silence, both
    thresholds, a fixed reference, digital silence (everything is 'loud' relative to itself)."""
    from oracle import trim as otrim
    rng = np.random.default_rng(3)
    for i in range(12):
        w = _wav(i % 6, 20000 + 777 * i)
        y = np.concatenate([1e-5 * rng.standard_normal(int(rng.integers(0, 9000))).astype(np.float32), w,
                            np.zeros(int(rng.integers(0, 9000)), np.float32)])
        for top_db in (60, 25):
            _, se = otrim.trim(y, top_db=top_db)
            assert hostemu.trim(y, top_db) == se
        _, se = otrim.trim(y, top_db=40, ref=0.5, frame_length=1024, hop_length=256)
        assert hostemu.trim(y, 40, 0.5, 1024, 256) == se
    z = np.zeros(5000, np.float32)
    assert hostemu.trim(z) == otrim.trim(z)[1] == (0, 5000)


def _torch_fbank_grad(x, length, go, n_fft=1024, hop=256, n_mels=80, guard=1e-5, log="clamp", lg=1e-5, power=1.0,
                      pad_value=-11.52, sr=22050):
    """torch autograd through the reference statements of FilterbankFeatures.forward with use_grads=True
    (features.py:403-452): stft, sqrt(re^2 + im^2 + 1e-5), pow, fb @, log, mask."""
    import torch
    from oracle import melfb
    xt = torch.tensor(x, dtype=torch.float32, requires_grad=True)
    win = torch.hann_window(n_fft, periodic=False)
    X = torch.stft(xt[None], n_fft, hop_length=hop, win_length=n_fft, center=True, window=win, return_complex=True,
                   pad_mode="reflect")
    mag = torch.sqrt(torch.view_as_real(X).pow(2).sum(-1) + guard)
    if power != 1.0:
        mag = mag.pow(power)
    fb = torch.tensor(melfb.mel_filterbank(sr, n_fft, n_mels, 0.0, None))[None]
    mel = torch.matmul(fb, mag)
    out = torch.log(torch.clamp(mel, min=lg)) if log == "clamp" else torch.log(mel + lg)
    seq = (length + 2 * (n_fft // 2) - n_fft) // hop + 1
    out = out.masked_fill((torch.arange(out.shape[-1]) >= seq)[None, None], pad_value)
    (out * torch.tensor(go)[None]).sum().backward()
    return xt.grad.numpy()


@pytest.mark.parametrize("length,kw,tkw", [
    (9000, {}, {}),
    (6000, {}, {}),                                                              # masked tail frames
    (9000, dict(log_mode="add", log_guard=1.0), dict(log="add", lg=1.0)),
    (9000, dict(mag_power=2.0), dict(power=2.0)),
])
def test_emu_fbank_backward_matches_torch_autograd(length, kw, tkw):
    """K1b: d loss / d audio through the preprocessor (use_grads=True) against torch autograd."""
    rng = np.random.default_rng(0)
    x = _wav(3, 9000)
    go = rng.standard_normal((80, 1 + len(x) // 256)).astype(np.float32)
    cfg = SupConfig(highfreq=None, spec_floor=1e-5, log_mode=kw.get("log_mode", "clamp"),
                    log_guard=kw.get("log_guard", 1e-5), mag_power=kw.get("mag_power", 1.0))
    ref = _torch_fbank_grad(x, length, go, **tkw)
    got = hostemu.fbank_backward(cfg, x, length, go)
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()
