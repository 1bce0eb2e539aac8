"""GPU parity at the sizes and shapes BASELINE.json's `configs` name, through the C ABI:

  C1  all 100 full-length 22.05 kHz utterances, every sup-data type, all four north-star gates
  C2  the first 500 utterances of the LJSpeech-shaped bench manifest (16-bit PCM ingest): pYIN bit-equal
  C3  a slice of the multispeaker manifest (0.5-20 s log-normal, 400 speakers): per-speaker statistics
  C4  44.1 kHz / n_fft 2048 / hop 512 / fmax None, utterances of 10-30 s (T up to ~2 580): log-mel, energy, pYIN, prior
  C5  the Conformer preprocessor at batch 256 through AudioToMelSpectrogramPreprocessor(input_signal=, length=)

The CPU side is the oracle (process pool, banded Viterbi = bit-identical to the dense one).  Gates are
BASELINE.json's: log-mel / energy 1e-4 relative, f0 within 1 cent on >= 99.9 % of voiced frames, voiced flags
and prior arg-max exact on >= 99.9 % of frames, pitch statistics 1e-5.  A summary of what was measured is
written to gpurun_out/parity_configs.json.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    yield
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_configs.json"), "w") as f:
            json.dump(REPORT, f, indent=1)
    except OSError:
        pass


def _rel(a, b):
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def _compare_all(out, refs, text_lens, n_mels=80):
    """CUDA outputs of one packed batch vs per-utterance oracle dicts -> measured quantities."""
    from roar_b200.extractor import split_frames
    fo = out["frame_off"]
    lms = split_frames(out["log_mel"], fo, n_mels) if out.get("log_mel") is not None else None
    f0a = out["pitch"].cpu().numpy()
    vfa = out["voiced_mask"].cpu().numpy()
    vpa = out["p_voiced"].cpu().numpy()
    ena = out["energy"].cpu().numpy() if out.get("energy") is not None else None
    m = dict(frames=0, voiced=0, flag_ok=0, cent_ok=0, f0_bit_equal=0, argmax_ok=0, prior_rows=0,
             log_mel_max=0.0, energy_max=0.0, p_voiced_max=0.0)
    for i, r in enumerate(refs):
        a, b = int(fo[i]), int(fo[i + 1])
        assert len(r["pitch"]) == b - a
        m["frames"] += b - a
        if lms is not None and "log_mel" in r:
            ref_lm = r["log_mel"][0] if r["log_mel"].ndim == 3 else r["log_mel"]
            assert tuple(lms[i].shape) == ref_lm.shape
            m["log_mel_max"] = max(m["log_mel_max"], float(_rel(lms[i].cpu().numpy(), ref_lm).max()))
        if ena is not None and "energy" in r:
            m["energy_max"] = max(m["energy_max"], float((np.abs(ena[a:b] - r["energy"]) / np.maximum(1e-12, np.abs(r["energy"]))).max()))
        f0, vf, vp = f0a[a:b], vfa[a:b], vpa[a:b]
        m["flag_ok"] += int((vf == r["voiced_mask"]).sum())
        m["f0_bit_equal"] += int((f0 == r["pitch"]).sum())
        both = (vf != 0) & (r["voiced_mask"] != 0)
        m["voiced"] += int(both.sum())
        m["cent_ok"] += int((np.abs(1200 * np.log2(f0[both] / r["pitch"][both])) <= 1.0).sum())
        m["p_voiced_max"] = max(m["p_voiced_max"], float(np.abs(vp - r["p_voiced"]).max()))
        if out.get("align_prior_matrix") is not None and "align_prior_matrix" in r:
            oo = out["prior_off"]
            pr = out["align_prior_matrix"][oo[i]:oo[i + 1]].view(b - a, text_lens[i]).cpu().numpy()
            m["argmax_ok"] += int((pr.argmax(1) == r["align_prior_matrix"].argmax(1)).sum())
            m["prior_rows"] += b - a
    return m


def _assert_gates(m, need_prior=True):
    assert m["log_mel_max"] <= 1e-4, m
    assert m["energy_max"] <= 1e-4, m
    assert m["flag_ok"] / m["frames"] >= 0.999, m
    assert m["voiced"] > 0 and m["cent_ok"] / m["voiced"] >= 0.999, m
    assert m["p_voiced_max"] <= 1e-5, m
    if need_prior:
        assert m["prior_rows"] == m["frames"] and m["argmax_ok"] / m["prior_rows"] >= 0.999, m


def test_config1_all_100_utterances_all_gates():
    from oracle import pool, stats as ostats
    from roar_b200 import synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor, finalize_pitch_stats
    man = synth.corpus_manifest("C1")
    assert len(man) == 100
    tasks = pool.tasks_for("C1", man)
    refs = pool.run(tasks)
    wavs = [pool.synth(t) for t in tasks]
    ex = SupDataExtractor(SupConfig(highfreq=8000.0))
    st = ex.new_pitch_partials(1)
    tl = [u.text_len for u in man]
    out = ex.extract(ex.pack(wavs), text_lens=tl, stats=st)
    torch.cuda.synchronize()
    m = _compare_all(out, refs, tl)
    _assert_gates(m)
    got = finalize_pitch_stats(st)
    ref = ostats.pitch_stats_f64([r["pitch"] for r in refs])
    assert got["count"] == ref["count"]
    assert abs(got["pitch_mean"] / ref["mean"] - 1) <= 1e-5 and abs(got["pitch_std"] / ref["std"] - 1) <= 1e-5
    assert got["pitch_min"] == ref["min"] and got["pitch_max"] == ref["max"]
    m["audio_seconds"] = float(sum(len(w) for w in wavs)) / 22050
    REPORT["C1"] = m


def test_config2_first_500_utterances_pcm16_ingest():
    """The bench manifest's head, ingested as 16-bit PCM (roar_sup_pcm16_to_f32): pYIN outputs must be bit-equal
    to the oracle run on the same quantised waveform, log-mel within 1e-4."""
    from oracle import pool
    from roar_b200 import synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor, pack_layout
    man = synth.corpus_manifest("C2", 500)
    ex = SupDataExtractor(SupConfig(highfreq=8000.0))
    tasks = pool.tasks_for("C2", man, what="pyin+logmel", pcm16=True, fb=ex.mel_filterbank())
    refs = pool.run(tasks)
    lens = np.array([u.n_samples for u in man], dtype=np.int64)
    offs, total = pack_layout(lens)
    host = torch.zeros(total, dtype=torch.int16, pin_memory=True)
    for t, o, n in zip(tasks, offs, lens):
        host.numpy()[o:o + n] = pool.quantize_pcm16(synth.synth_utterance(t["seed"], t["utt_id"], t["n_samples"], t["sr"], t["speaker"]))
    batch = ex.pack_pcm16_from_host_buffer(host, offs, lens)
    out = ex.extract(batch, types=("log_mel", "pitch", "voiced_mask", "p_voiced"))
    torch.cuda.synchronize()
    m = _compare_all(out, refs, None)
    assert m["frames"] > 250000
    assert m["flag_ok"] == m["frames"] and m["f0_bit_equal"] == m["frames"], m
    assert m["p_voiced_max"] <= 1e-6 and m["log_mel_max"] <= 1e-4, m
    REPORT["C2_500_pcm16"] = m


def test_config3_slice_per_speaker_stats():
    """Multispeaker manifest slice (log-normal 0.5-20 s, 400 speakers): parity + `default` and per-speaker
    statistics (compute_speaker_stats.py:105-132) against oracle.stats."""
    from oracle import pool, stats as ostats
    from roar_b200 import synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor
    man = synth.corpus_manifest("C3", 400)
    durs = np.array([u.duration for u in man])
    assert durs.min() < 1.5 and durs.max() > 12.0           # the ragged mix the config names
    ex = SupDataExtractor(SupConfig(highfreq=8000.0))
    tasks = pool.tasks_for("C3", man, what="pyin+logmel", fb=ex.mel_filterbank())
    refs = pool.run(tasks)
    wavs = [pool.synth(t) for t in tasks]
    out = ex.extract(ex.pack(wavs), types=("log_mel", "pitch", "voiced_mask", "p_voiced"))
    m = _compare_all(out, refs, None)
    assert m["flag_ok"] / m["frames"] >= 0.999 and m["cent_ok"] / m["voiced"] >= 0.999 and m["log_mel_max"] <= 1e-4, m
    spk = np.array([u.speaker for u in man])
    ids = {s: k + 1 for k, s in enumerate(sorted(set(spk.tolist())))}
    groups = np.array([ids[s] for s in spk], dtype=np.int32)
    acc = ex.new_pitch_partials(len(ids) + 1)
    ex.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], groups, len(ids) + 1, acc)
    ex.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], np.zeros(len(man), np.int32), len(ids) + 1, acc)
    acc = acc.cpu().numpy()
    from roar_b200.extract_sup_data import stats_from_partials
    ref_all = ostats.pitch_stats_f64([r["pitch"] for r in refs])
    got_all = stats_from_partials(acc[0])
    assert abs(got_all["pitch_mean"] / ref_all["mean"] - 1) <= 1e-5 and abs(got_all["pitch_std"] / ref_all["std"] - 1) <= 1e-5
    checked = 0
    for s, g in ids.items():
        ps = [r["pitch"] for r, q in zip(refs, spk) if q == s]
        if sum(int((p != 0).sum()) for p in ps) < 2:
            continue
        r = ostats.pitch_stats_f64(ps)
        got = stats_from_partials(acc[g])
        assert abs(got["pitch_mean"] / r["mean"] - 1) <= 1e-5 and abs(got["pitch_std"] / r["std"] - 1) <= 1e-5, s
        assert got["pitch_min"] == r["min"] and got["pitch_max"] == r["max"]
        checked += 1
    assert checked >= 100
    m["speakers_checked"] = checked
    REPORT["C3_slice"] = m


def test_config4_long_44k_all_types():
    """HiFiTTS shape: 44.1 kHz, n_fft = win 2048, hop 512, 80 mels, fmax None; six utterances of 10-30 s incl. the
    two longest of the manifest head (T ~ 2 500 frames: the long Viterbi chain)."""
    from oracle import pool
    from roar_b200 import synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor
    head = synth.corpus_manifest("C4", 40)
    order = np.argsort([-u.n_samples for u in head])
    man = [head[i] for i in list(order[:2]) + [0, 1, 2, int(order[-1])]]
    assert max(u.duration for u in man) > 28.0 and min(u.duration for u in man) >= 10.0
    cfg = dict(n_fft=2048, hop_length=512, win_length=2048, fmax=None)
    tasks = pool.tasks_for("C4", man, cfg=cfg)
    refs = pool.run(tasks)
    wavs = [pool.synth(t) for t in tasks]
    ex = SupDataExtractor(SupConfig(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, highfreq=None))
    tl = [u.text_len for u in man]
    out = ex.extract(ex.pack(wavs), text_lens=tl)
    m = _compare_all(out, refs, tl)
    assert m["frames"] > 8000
    _assert_gates(m)
    assert m["f0_bit_equal"] == m["frames"], m
    REPORT["C4_long"] = m


def test_config5_preprocessor_batch256():
    """AudioToMelSpectrogramPreprocessor(input_signal=, length=) (audio_preprocessing.py:77-82) at the config's
    batch 256 x U(2, 16.7) s.  The reference statements in float32 (torch.stft on the CPU) are the target; the
    same statements in float64 measure how far the reference's OWN float32 result sits from exact arithmetic.
    Power spectrum + log(x + 2^-24): in (near-)silent frames the float32 FFT rounding of either implementation is
    comparable to the 6e-8 guard, so there the 1e-4 tolerance is widened by the reference's own measured error in
    that frame -- and nowhere else."""
    from oracle import fbank as ofbank
    from roar_b200 import synth
    from roar_b200.features import AudioToMelSpectrogramPreprocessor
    B = 256
    man = synth.corpus_manifest("C5", B)
    wavs = [synth.synth_utterance(5, u.utt_id, u.n_samples, 16000, u.speaker) for u in man]
    lens = np.array([len(w) for w in wavs], dtype=np.int64)
    assert lens.min() >= 2 * 16000 and lens.max() <= int(16.7 * 16000) + 1
    x = np.zeros((B, int(lens.max())), dtype=np.float32)
    for i, w in enumerate(wavs):
        x[i, :len(w)] = w
    pre = AudioToMelSpectrogramPreprocessor(sample_rate=16000, window_size=0.025, window_stride=0.01, features=80,
                                            n_fft=512, dither=0.0).cuda().eval()
    with pytest.raises(TypeError):
        pre(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda())       # kwargs-only, like @typecheck
    got, got_len = pre(input_signal=torch.from_numpy(x).cuda(), length=torch.from_numpy(lens).cuda())
    g = got.cpu().numpy()
    orc = ofbank.FilterbankFeaturesOracle(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=80, n_fft=512)
    worst = dict(max_vs_ref=0.0, max_ref_vs_f64=0.0, max_gpu_vs_f64=0.0, above_1e4=0, elements=0, outside_widened=0)
    q_gpu, q_ref = [], []
    for lo in range(0, B, 32):                      # the oracle in slices: bounded host memory
        sl = slice(lo, lo + 32)
        ref, ref_len = orc.forward(x[sl], lens[sl])
        tru, _ = orc.forward(x[sl], lens[sl], dtype=torch.float64)
        ref, tru = np.asarray(ref), np.asarray(tru)
        assert np.array_equal(got_len[sl].cpu().numpy(), np.asarray(ref_len))
        gg = g[sl][:, :, :ref.shape[2]]
        assert g.shape[2] >= ref.shape[2] and (g[sl][:, :, ref.shape[2]:] == 0).all()    # pad_to 16 of the batch max
        scale = np.maximum(1.0, np.abs(ref))
        d = np.abs(gg - ref)
        e_ref = np.abs(ref - tru)
        frame_noise = e_ref.max(axis=1, keepdims=True)       # the reference's own float32 error in that frame
        worst["outside_widened"] += int((d > 1e-4 * scale + 4.0 * frame_noise).sum())
        worst["above_1e4"] += int((d > 1e-4 * scale).sum())
        worst["elements"] += d.size
        worst["max_vs_ref"] = max(worst["max_vs_ref"], float((d / scale).max()))
        worst["max_ref_vs_f64"] = max(worst["max_ref_vs_f64"], float((e_ref / scale).max()))
        worst["max_gpu_vs_f64"] = max(worst["max_gpu_vs_f64"], float((np.abs(gg - tru) / scale).max()))
        q_gpu.append(np.quantile(np.abs(gg - tru) / scale, 0.999))
        q_ref.append(np.quantile(e_ref / scale, 0.999))
    worst["q999_gpu_vs_f64"], worst["q999_ref_vs_f64"] = float(np.max(q_gpu)), float(np.max(q_ref))
    REPORT["C5_batch256"] = worst
    assert tuple(g.shape[:2]) == (B, 80)
    assert worst["outside_widened"] == 0, worst
    assert worst["above_1e4"] <= 1e-4 * worst["elements"], worst
    # as close to exact arithmetic as the reference itself is
    assert worst["q999_gpu_vs_f64"] <= max(2.0 * worst["q999_ref_vs_f64"], 2e-5), worst
    assert worst["max_gpu_vs_f64"] <= max(3.0 * worst["max_ref_vs_f64"], 1e-4), worst


@pytest.mark.parametrize("B", [128, 512])
def test_fbank_large_batches_forward_and_backward(B):
    """the fbank workspace must hold for ordinary ASR / mel-loss batch sizes (B = 512 forward,
    B >= 96 backward used to overflow the workspace sized by roar_sup_workspace_bytes)."""
    from oracle import fbank as ofbank
    from roar_b200.features import FilterbankFeatures
    rng = np.random.default_rng(B)
    L = 4000
    lens = rng.integers(1200, L + 1, size=B)
    x = (0.1 * rng.standard_normal((B, L))).astype(np.float32)
    for i in range(B):
        x[i, lens[i]:] = 0
    m = FilterbankFeatures(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=40, n_fft=512, dither=0.0,
                           normalize=None, preemph=None, log_zero_guard_type="clamp", log_zero_guard_value=1e-5,
                           mag_power=1.0, pad_to=0, use_grads=True).cuda()
    xg = torch.tensor(x, device="cuda", requires_grad=True)
    out, out_len = m(xg, torch.tensor(lens, device="cuda"))
    out.sum().backward()
    assert torch.isfinite(xg.grad).all() and float(xg.grad.abs().sum()) > 0
    orc = ofbank.FilterbankFeaturesOracle(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=40, n_fft=512,
                                          normalize=None, preemph=None, log_zero_guard_type="clamp",
                                          log_zero_guard_value=1e-5, mag_power=1.0, pad_to=0, use_grads=True)
    ref, ref_len = orc.forward(x[:16], lens[:16])
    assert np.array_equal(out_len.cpu().numpy()[:16], np.asarray(ref_len))
    assert _rel(out.detach().cpu().numpy()[:16], np.asarray(ref)).max() <= 1e-4


def test_viterbi_pruned_equals_exhaustive_on_500_utterances(monkeypatch):
    """The pruned Viterbi kernel against the kernel that visits every in-band source, on 500 utterances of the
    bench manifest (~280 k frames) and three long 44.1 kHz ones: identical f0 / flags on every frame."""
    from roar_b200 import synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor
    for corpus, n, cfg in (("C2", 500, SupConfig(highfreq=8000.0)),
                           ("C4", 3, SupConfig(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, highfreq=None))):
        man, audio, offs, lens = synth.synth_corpus_device(corpus, "cuda", n_utts=n)
        monkeypatch.delenv("ROAR_SUP_VITERBI", raising=False)
        fast = SupDataExtractor(cfg)
        monkeypatch.setenv("ROAR_SUP_VITERBI", "generic")
        gen = SupDataExtractor(cfg)
        monkeypatch.delenv("ROAR_SUP_VITERBI", raising=False)
        o, l = offs.cpu().numpy(), lens.cpu().numpy().astype(np.int64)
        a = fast.pyin(fast.batch_from_device(audio, o, l))
        b = gen.pyin(gen.batch_from_device(audio, o, l))
        assert a[0].numel() > (250000 if corpus == "C2" else 4000)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), corpus


def test_pcm16_ingest_exact_and_unaligned():
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor
    ex = SupDataExtractor(SupConfig(highfreq=8000.0))
    rng = np.random.default_rng(1)
    for n in (1, 7, 8, 4099, 100003):
        pcm = rng.integers(-32768, 32768, size=n + 3, dtype=np.int16)
        pcm[:2] = [-32768, 32767]
        for shift in (0, 1, 3):       # 16-byte aligned and unaligned device pointers
            d = torch.from_numpy(pcm).cuda()[shift:shift + n]
            got = ex.pcm16_to_f32(d).cpu().numpy()
            assert np.array_equal(got, pcm[shift:shift + n].astype(np.float32) / np.float32(32768.0))


def test_cli_streaming_cache_roundtrip_through_reference_collate(tmp_path, capsys):
    """Row a15 end to end: 16-bit wav files -> streaming CLI (several small batches, two in flight) -> `.pt`
    cache -> the reference's load + collate logic (tests/ref_contract.py, pinned to the reference's own
    `_collate_fn` output) -> the batch tuple a FastPitch training step unpacks, compared with the same tuple
    built from the CPU oracle.  Also: `mel_filepath` override (dataset.py:646-649), `ignore_file` pruning."""
    import pickle
    import sys
    from scipy.io import wavfile
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_contract as RC
    from oracle import pool, prior as oprior, stats as ostats
    from roar_b200 import extract_sup_data as X, synth
    man = synth.corpus_manifest("C1", 9)
    tasks = pool.tasks_for("C1", man, pcm16=True)
    for t in tasks:
        t["n_samples"] = min(t["n_samples"], 22050 * 3 + 1000 * t["utt_id"])
    rows, audio = [], []
    for k, t in enumerate(tasks):
        y = synth.synth_utterance(t["seed"], t["utt_id"], t["n_samples"], t["sr"], t["speaker"])
        p = tmp_path / "wavs" / f"spk{t['speaker']}" / f"utt{k}.wav"
        p.parent.mkdir(parents=True, exist_ok=True)
        wavfile.write(p, 22050, pool.quantize_pcm16(y))
        audio.append(pool.synth(t))
        rows.append({"audio_filepath": str(p), "duration": t["n_samples"] / 22050, "text": "x" * (t["text_len"] - 2),
                     "speaker": t["speaker"]})
    # utterance 7 carries its own mel file; utterance 8 is on the ignore list
    own_mel = tmp_path / "own_mel.pt"
    torch.save(torch.full((1, 80, 5), 7.0), own_mel)
    rows[7]["mel_filepath"] = str(own_mel)
    ig = tmp_path / "ignore.pkl"
    with open(ig, "wb") as f:
        pickle.dump([rows[8]["audio_filepath"]], f)
    mf = tmp_path / "train.json"
    mf.write_text("\n".join(json.dumps(r) for r in rows) + "\n")
    sup = tmp_path / "sup"
    order = RC.TYPE_TABLE["orders"]["all"]
    res = X.main([f"manifest_filepath={mf}", f"sup_data_path={sup}", f"sup_data_types=[{','.join(order)}]",
                  f"dataset.ignore_file={ig}", "batch_audio_seconds=7", "pipeline_depth=2"])
    capsys.readouterr()
    assert res["run"]["utterances"] == 8 and res["run"]["batches"] >= 3
    base = X.get_base_dir([r["audio_filepath"] for r in rows[:8]])
    uids = [X.rel_audio_id(r["audio_filepath"], base) for r in rows]
    assert not (sup / "log_mel" / f"{uids[7]}.pt").exists() and (sup / "pitch" / f"{uids[7]}.pt").exists()
    assert not (sup / "pitch" / f"{uids[8]}.pt").exists()
    assert not [f for f in os.listdir(sup / "pitch") if ".tmp" in f]
    refs = pool.run(tasks[:7])
    folders = {t: sup / t for t in ("log_mel", "pitch", "voiced_mask", "p_voiced", "energy")}
    ref_stats = ostats.pitch_stats_f64([r["pitch"] for r in pool.run(tasks[:8])])
    assert abs(res["pitch_mean"] / ref_stats["mean"] - 1) <= 1e-5 and abs(res["pitch_std"] / ref_stats["std"] - 1) <= 1e-5
    table = json.load(open(sup / "pitch_stats.json"))
    mean, std = table["default"]["pitch_mean"], table["default"]["pitch_std"]
    got_items, ref_items = [], []
    for k in range(7):
        tokens = np.arange(1, tasks[k]["text_len"] + 1)
        T = 1 + len(audio[k]) // 256
        got_items.append(RC.getitem_from_cache(audio[k], tokens, uids[k], folders, order, speaker_id=tasks[k]["speaker"],
                                               prior_fn=lambda n, T=T: oprior.beta_binomial_prior_distribution(n, T),
                                               pitch_mean=mean, pitch_std=std))
        r = refs[k]
        pitch = torch.from_numpy(r["pitch"].copy())
        pitch -= mean
        pitch[pitch == -mean] = 0.0
        pitch /= std
        lm = torch.from_numpy(r["log_mel"]).squeeze(0)
        ref_items.append((torch.from_numpy(audio[k]), torch.tensor(len(audio[k])).long(), torch.from_numpy(tokens).long(),
                          torch.tensor(len(tokens)).long(), lm, torch.tensor(lm.shape[1]).long(), None,
                          torch.from_numpy(r["align_prior_matrix"]), pitch, torch.tensor(len(pitch)).long(),
                          torch.from_numpy(r["energy"]), torch.tensor(len(r["energy"])).long(),
                          torch.tensor(tasks[k]["speaker"]).long(), torch.from_numpy(r["voiced_mask"]),
                          torch.from_numpy(r["p_voiced"]), None, None, None))
    for sup_order in RC.TYPE_TABLE["orders"].values():
        got = RC.collate(got_items, sup_order)
        ref = RC.collate(ref_items, sup_order)
        names = []
        for n in RC.MAIN + list(sup_order):
            names += [n] + ([f"{n}_lens"] if n in RC.WITH_LENS else [])
        assert len(got) == len(ref) == len(names)
        for n, a, b in zip(names, got, ref):
            assert a.shape == b.shape and a.dtype == b.dtype, n
            if n in ("log_mel", "energy"):
                assert float((torch.abs(a - b) / torch.clamp(torch.abs(b), min=1.0 if n == "log_mel" else 1e-12)).max()) <= 1e-4, n
            elif n == "p_voiced":
                assert float(torch.abs(a - b).max()) <= 1e-5
            else:
                assert torch.equal(a, b), n          # audio, text, lens, prior, normalised pitch, voiced mask, speaker
    REPORT["cli_roundtrip"] = {"utterances": 7, "orders": list(RC.TYPE_TABLE["orders"]), "run": res["run"]}


def test_fbank_frame_splicing_fixed_stats_pad_to_max():
    """FilterbankFeatures options finished with the reference's tensor statements after the kernel
    (features.py:434-460): frame_splicing > 1, a fixed mean / std table, pad_to="max"."""
    from oracle import fbank as ofbank
    from roar_b200.features import FilterbankFeatures
    rng = np.random.default_rng(3)
    B, L = 3, 12000
    lens = np.array([12000, 7000, 9001])
    x = (0.2 * rng.standard_normal((B, L))).astype(np.float32)
    for i in range(B):
        x[i, lens[i]:] = 0
    base = dict(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=16, n_fft=512)
    fixed = {"fixed_mean": rng.standard_normal((B, 16)).tolist(), "fixed_std": (1 + rng.random((B, 16))).tolist()}
    cases = [dict(frame_splicing=3, normalize="per_feature", pad_to=16),
             dict(frame_splicing=2, normalize="all_features", pad_to="max", max_duration=1.0, pad_value=-2.0),
             dict(normalize=fixed, pad_to=8),
             dict(normalize=None, pad_to="max", max_duration=1.0)]
    for kw in cases:
        m = FilterbankFeatures(dither=0.0, **base, **kw).cuda().eval()
        got, got_len = m(torch.from_numpy(x).cuda(), torch.from_numpy(lens).cuda())
        ref, ref_len = ofbank.FilterbankFeaturesOracle(**base, **kw).forward(x, lens)
        assert tuple(got.shape) == ref.shape, (kw, tuple(got.shape), ref.shape)
        assert np.array_equal(got_len.cpu().numpy(), ref_len)
        err = _rel(got.cpu().numpy(), ref)
        assert np.quantile(err, 0.999) <= 1e-4 and err.max() <= 2e-3, (kw, float(err.max()))


def test_resampler_vs_scipy_and_cli_mixed_rates(tmp_path, capsys):
    """Row N3: roar_sup_resample against scipy.signal.resample_poly (the kernel's arithmetic; parity with the
    reference's soxr_hq is unpinned, oracle/resample.py), a pass-band tone check, and the CLI on a corpus that mixes
    22.05 kHz files with 44.1 kHz / 16 kHz ones (resampled on the GPU, then identical to feeding the resampled audio)."""
    from scipy.io import wavfile
    from oracle import pool, resample as oresample, spec as ospec
    from roar_b200 import extract_sup_data as X, synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor
    from roar_b200.resample import resample_batch
    ex = SupDataExtractor(SupConfig(highfreq=8000.0))
    rng = np.random.default_rng(4)
    for a, b in ((44100, 22050), (48000, 22050), (16000, 22050), (22050, 16000)):
        wavs = [(0.3 * rng.standard_normal(n)).astype(np.float32) for n in (1, 257, 10007, 44100)]
        out = resample_batch(ex, ex.pack(wavs), a, b)
        for i, w in enumerate(wavs):
            ref = oresample.resample(w, a, b)
            got = out.audio[int(out.offs_host[i]):int(out.offs_host[i] + out.lens_host[i])].cpu().numpy()
            assert got.shape == ref.shape, (a, b, got.shape, ref.shape)
            assert np.abs(got - ref).max() <= 2e-6, (a, b, float(np.abs(got - ref).max()))
    t = np.arange(44100) / 44100.0
    tone = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    y = resample_batch(ex, ex.pack([tone]), 44100, 22050).audio.cpu().numpy()[:22050]
    assert oresample.band_limited_check(y, 22050, 1000.0, 0.5)
    # CLI: mixed sample rates in one manifest
    rows, expect = [], []
    for k, sr in enumerate((22050, 44100, 16000, 22050)):
        yk = synth.synth_utterance(77, k, int(1.5 * sr) + 100 * k, sr, k)
        p = tmp_path / "w" / f"u{k}.wav"
        p.parent.mkdir(exist_ok=True)
        wavfile.write(p, sr, pool.quantize_pcm16(yk))
        y16 = pool.quantize_pcm16(yk).astype(np.float32) / np.float32(32768.0)
        expect.append(y16 if sr == 22050 else oresample.resample(y16, sr, 22050))
        rows.append(json.dumps({"audio_filepath": str(p), "duration": len(yk) / sr, "text": "abc"}))
    mf = tmp_path / "m.json"
    mf.write_text("\n".join(rows) + "\n")
    sup = tmp_path / "sup"
    X.main([f"manifest_filepath={mf}", f"sup_data_path={sup}", "sup_data_types=[log_mel,energy]"])
    capsys.readouterr()
    fb = ex.mel_filterbank()
    for k, yk in enumerate(expect):
        lm = torch.load(sup / "log_mel" / f"u{k}.pt").numpy()
        ref = ospec.get_log_mel(yk, fb).numpy()
        assert lm.shape == ref.shape, (k, lm.shape, ref.shape)
        assert _rel(lm, ref).max() <= 2e-4, (k, float(_rel(lm, ref).max()))     # resampler rounding (2e-6 abs) on top of the 1e-4 gate


def test_cli_packed_cache_equals_pt_cache(tmp_path, capsys):
    """`cache_format=packed`: the packed shards + index hold exactly the tensors of the per-file layout."""
    from scipy.io import wavfile
    from oracle import pool
    from roar_b200 import extract_sup_data as X, synth
    from roar_b200.dataset_utils import PackedCache
    rows = []
    for k in range(5):
        y = synth.synth_utterance(99, k, 22050 + 3000 * k, 22050, k)
        p = tmp_path / "w" / f"u{k}.wav"
        p.parent.mkdir(exist_ok=True)
        wavfile.write(p, 22050, pool.quantize_pcm16(y))
        rows.append(json.dumps({"audio_filepath": str(p), "duration": len(y) / 22050, "text": "abc"}))
    mf = tmp_path / "m.json"
    mf.write_text("\n".join(rows) + "\n")
    types = "sup_data_types=[log_mel,pitch,voiced_mask,p_voiced,energy]"
    a = X.main([f"manifest_filepath={mf}", f"sup_data_path={tmp_path / 'pt'}", types, "batch_audio_seconds=3"])
    b = X.main([f"manifest_filepath={mf}", f"sup_data_path={tmp_path / 'pk'}", types, "batch_audio_seconds=3", "cache_format=packed"])
    capsys.readouterr()
    assert a["pitch_mean"] == b["pitch_mean"] and b["run"]["batches"] >= 2
    cache = PackedCache(tmp_path / "pk" / "packed")
    for k in range(5):
        for t in ("log_mel", "pitch", "voiced_mask", "p_voiced", "energy"):
            ref = torch.load(tmp_path / "pt" / t / f"u{k}.pt")
            got = cache.load(t, f"u{k}")
            assert got.dtype == ref.dtype and got.shape == ref.shape and torch.equal(got, ref), (t, k)
    # resumed packed run: nothing left to do
    c = X.main([f"manifest_filepath={mf}", f"sup_data_path={tmp_path / 'pk'}", types, "cache_format=packed"])
    capsys.readouterr()
    assert c is None or c.get("run", {}).get("utterances", 0) == 0
