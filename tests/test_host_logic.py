"""CPU tests of the host-side logic of the extraction driver (no GPU): config overrides, manifest
handling, cache naming, sharding, statistics and the world_size-2 exchange step on gloo."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from roar_b200 import extract_sup_data as X
from roar_b200.config import SupConfig
from roar_b200.extractor import finalize_pitch_stats, pack_layout


def test_overrides_and_interpolation(tmp_path):
    cfg = X.load_config(["manifest_filepath=a.json", "sup_data_path=out", "sup_data_types=[pitch,energy,log_mel]",
                         "dataset.sample_rate=44100", "dataset.n_fft=2048", "dataset.hop_length=512",
                         "dataset.win_length=2048", "dataset.highfreq=null", "dataloader_params.num_workers=4"])
    assert cfg["sup_data_types"] == ["pitch", "energy", "log_mel"]
    assert cfg["dataset"]["sample_rate"] == 44100 and cfg["dataset"]["highfreq"] is None
    s = X.sup_config_from(cfg)
    assert (s.sample_rate, s.n_fft, s.hop, s.win, s.highfreq) == (44100, 2048, 512, 2048, None)
    y = tmp_path / "ds.yaml"
    y.write_text("name: ds_for_fastpitch_align\nmanifest_filepath: m.json\nsup_data_path: sd\n"
                 "sup_data_types: [align_prior_matrix, pitch]\n"
                 "dataset:\n  manifest_filepath: ${manifest_filepath}\n  sup_data_path: ${sup_data_path}\n"
                 "  sample_rate: 22050\n  n_fft: 1024\n  hop_length: 256\n  win_length: 1024\n  n_mels: 80\n"
                 "  highfreq: 8000\n")
    cfg = X.load_config([f"--config-path={tmp_path}", "--config-name=ds", "manifest_filepath=zz.json"])
    assert cfg["dataset"]["manifest_filepath"] == "zz.json" and cfg["dataset"]["sup_data_path"] == "sd"


def test_default_config_matches_reference_yaml_values():
    s = X.sup_config_from(X.load_config([]))
    assert (s.sample_rate, s.n_fft, s.win, s.hop, s.n_mels, s.lowfreq, s.highfreq) == (22050, 1024, 1024, 256, 80, 0.0, 8000.0)
    assert s.pitch_fmin == pytest.approx(65.40639132514966) and s.pitch_fmax == pytest.approx(2093.004522404789)
    assert s.pyin_frame == 1024 and s.pyin_hop == 256   # pyin ignores hop_length: frame_length // 4


def test_ids_base_dir_and_manifest(tmp_path):
    paths = [tmp_path / "spk1" / "a" / "u1.wav", tmp_path / "spk1" / "u2.wav", tmp_path / "spk2" / "u3.wav"]
    base = X.get_base_dir([str(p) for p in paths])
    assert base == tmp_path
    assert X.rel_audio_id(str(paths[0]), base) == "spk1_a_u1"
    assert X.rel_audio_id(str(paths[2]), base) == "spk2_u3"
    m = tmp_path / "m.json"
    with open(m, "w") as f:
        for p, d in zip(paths, [0.05, 3.0, 30.0]):
            f.write(json.dumps({"audio_filepath": str(p), "duration": d, "text": "abc"}) + "\n")
    items = X.read_manifest(str(m), min_duration=0.1, max_duration=20.0)
    assert [Path_stem(i) for i in items] == ["u2"]
    assert X.text_length({"text": "hello"}) == 7 and X.text_length({"text_tokens": [1, 2, 3]}) == 3
    # TTSDataset.filter_files (dataset.py:367-406): the ignore file lists full audio_filepath strings ...
    import pickle
    ig = tmp_path / "ignore.pkl"
    with open(ig, "wb") as f:
        pickle.dump([str(paths[1]), "u3"], f)          # a bare stem does NOT match
    items = X.read_manifest(str(m), ignore_file=str(ig))
    assert [Path_stem(i) for i in items] == ["u1", "u3"]
    # ... several manifests are concatenated, and the duration filter is off unless EVERY entry has a duration
    m2 = tmp_path / "m2.json"
    m2.write_text(json.dumps({"audio_filepath": str(tmp_path / "spk3" / "u4.wav"), "text": "x"}) + "\n")
    items = X.read_manifest([str(m), str(m2)], min_duration=0.1, max_duration=20.0)
    assert [Path_stem(i) for i in items] == ["u1", "u2", "u3", "u4"]


def Path_stem(item):
    return os.path.splitext(os.path.basename(item["audio_filepath"]))[0]


def test_sharding_partitions_and_balances():
    rng = np.random.default_rng(0)
    durs = rng.uniform(0.5, 20.0, size=1001)
    for world in (1, 2, 4, 8):
        shards = [X.shard_indices(durs, world, r) for r in range(world)]
        allidx = np.concatenate(shards)
        assert len(allidx) == len(durs) and len(set(allidx.tolist())) == len(durs)
        tot = np.array([durs[s].sum() for s in shards])
        # longest-processing-time greedy: every rank within one (short) utterance of the mean
        assert tot.max() - tot.min() <= durs.min() + 1e-9 or tot.max() / tot.min() < 1.001
        assert all(np.array_equal(s, np.sort(s)) for s in shards)
    # a skewed corpus (one very long utterance) still balances as well as it can
    skew = np.array([100.0] + [1.0] * 300)
    tot = np.array([skew[X.shard_indices(skew, 4, r)].sum() for r in range(4)])
    assert tot.max() == 100.0 and tot.min() == 100.0
    b = X.make_batches(np.array([10, 50, 20, 40, 30]), 60)
    assert sorted(np.concatenate(b).tolist()) == [0, 1, 2, 3, 4]
    assert all(sum([10, 50, 20, 40, 30][i] for i in g) <= 60 or len(g) == 1 for g in b)


def test_pack_layout_alignment():
    offs, total = pack_layout(np.array([5, 8, 1, 13]))
    assert offs.tolist() == [0, 8, 16, 20] and total == 36 and all(o % 4 == 0 for o in offs)


def test_partials_and_stats_match_oracle():
    from oracle import stats as ostats
    rng = np.random.default_rng(1)
    ps = [np.where(rng.random(300) < 0.6, rng.uniform(80, 300, 300), 0).astype(np.float32) for _ in range(9)]
    t = X.empty_partials(1)
    for p in ps:
        t[0] = X.merge_partials(t[0], torch.from_numpy(X.partials_from_pitch(p)))
    got = X.stats_from_partials(t[0])
    ref = ostats.pitch_stats_f64(ps)
    assert got["pitch_mean"] == pytest.approx(ref["mean"], rel=1e-12)
    assert got["pitch_std"] == pytest.approx(ref["std"], rel=1e-10)
    assert got["pitch_min"] == ref["min"] and got["pitch_max"] == ref["max"]
    assert finalize_pitch_stats(t)["pitch_std"] == pytest.approx(ref["std"], rel=1e-10)
    ref32 = ostats.pitch_stats(ps)
    assert got["pitch_mean"] == pytest.approx(ref32["mean"], rel=1e-5)
    assert got["pitch_std"] == pytest.approx(ref32["std"], rel=1e-5)
    assert X.stats_from_partials(X.empty_partials(1)[0]) is None


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    durs = rng.uniform(1, 10, 40)
    pitches = [np.where(rng.random(200) < 0.7, rng.uniform(80, 300, 200), 0).astype(np.float32) for _ in durs]
    mine = X.shard_indices(durs, world, rank)
    t = X.empty_partials(3)
    for i in mine:
        part = torch.from_numpy(X.partials_from_pitch(pitches[i]))
        t[0] = X.merge_partials(t[0], part)
        t[1 + i % 2] = X.merge_partials(t[1 + i % 2], part)
    red = X.allreduce_partials(t)
    q.put((rank, red.numpy().tolist()))
    dist.destroy_process_group()


def test_allreduce_partials_world2_gloo():
    import torch.multiprocessing as mp
    from oracle import stats as ostats
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]          # every rank ends with the global statistics
    rng = np.random.default_rng(7)
    durs = rng.uniform(1, 10, 40)
    pitches = [np.where(rng.random(200) < 0.7, rng.uniform(80, 300, 200), 0).astype(np.float32) for _ in durs]
    ref = ostats.pitch_stats_f64(pitches)
    got = X.stats_from_partials(res[0][0])
    assert got["pitch_mean"] == pytest.approx(ref["mean"], rel=1e-12)
    assert got["pitch_std"] == pytest.approx(ref["std"], rel=1e-10)
    assert got["pitch_min"] == ref["min"] and got["pitch_max"] == ref["max"]
    even = ostats.pitch_stats_f64([p for i, p in enumerate(pitches) if i % 2 == 0])
    assert X.stats_from_partials(res[0][1])["pitch_mean"] == pytest.approx(even["mean"], rel=1e-12)


def test_wav_decode_int16_and_float(tmp_path):
    from scipy.io import wavfile
    x = (np.sin(np.arange(2000) / 7.0) * 0.5).astype(np.float32)
    wavfile.write(tmp_path / "f.wav", 22050, x)
    wavfile.write(tmp_path / "i.wav", 22050, (x * 32767).astype(np.int16))
    a = X.load_wav(str(tmp_path / "f.wav"), 22050)
    b = X.load_wav(str(tmp_path / "i.wav"), 22050)
    assert a.dtype == np.float32 and np.array_equal(a, x) and np.abs(b - x).max() < 1e-4
    assert np.array_equal(b, (x * 32767).astype(np.int16).astype(np.float32) / 32768)     # x / 2^15, exact
    with pytest.raises(ValueError):
        X.load_wav(str(tmp_path / "f.wav"), 16000)


def test_native_wav_decoder_formats_offset_duration(tmp_path):
    """roar_sup_wav_* against scipy's decoder and AudioSegment's scaling rules (segment.py:140-153, 218-224):
    PCM 8/16/24/32, float 32/64, stereo (average / one channel), offset + duration, raw int16 for the GPU path."""
    import struct
    import torch
    from scipy.io import wavfile
    from roar_b200 import host_io as H
    rng = np.random.default_rng(0)
    y = (0.5 * rng.standard_normal(30001)).clip(-1, 1)
    i16 = (y * 32767).astype(np.int16)
    i32 = (y * 2147483000).astype(np.int32)
    u8 = ((y * 127) + 128).astype(np.uint8)
    st = np.stack([i16, (i16 // 2).astype(np.int16)], axis=1)
    cases = {"f32": (y.astype(np.float32), y.astype(np.float32)),
             "i16": (i16, i16.astype(np.float32) / 32768),
             "i32": (i32, (i32.astype(np.float64) / 2 ** 31).astype(np.float32)),
             "u8": (u8, (u8.astype(np.float32) - 128) / 128),
             "f64": (y, y.astype(np.float32)),
             "st": (st, (st.astype(np.float32) / 32768).mean(axis=1))}
    for k, (data, ref) in cases.items():
        wavfile.write(tmp_path / f"{k}.wav", 22050, data)
        got, sr = H.load_wav(str(tmp_path / f"{k}.wav"))
        assert sr == 22050 and got.dtype == np.float32 and np.array_equal(got, ref), k
    got, _ = H.load_wav(str(tmp_path / "st.wav"), channel=1)
    assert np.array_equal(got, st[:, 1].astype(np.float32) / 32768)
    got, _ = H.load_wav(str(tmp_path / "i16.wav"), offset=0.5, duration=0.25)
    assert np.array_equal(got, cases["i16"][1][11025:11025 + 5512])
    i24 = (y[:1000] * 8388607).astype(np.int32)
    raw = b"".join(struct.pack("<i", int(v))[:3] for v in i24)
    hdr = (b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 22050, 66150, 3, 24)
           + b"data" + struct.pack("<I", len(raw)))
    (tmp_path / "i24.wav").write_bytes(hdr + raw)
    got, _ = H.load_wav(str(tmp_path / "i24.wav"))
    assert np.array_equal(got, (i24.astype(np.float64) / 2 ** 23).astype(np.float32))
    info = H.wav_probe([str(tmp_path / "i16.wav"), str(tmp_path / "f32.wav")])
    assert H.is_pcm16_mono(info).tolist() == [True, False] and info["n_frames"].tolist() == [30001, 30001]
    dst = torch.zeros(30001 + 8, dtype=torch.int16)
    H.wav_read([str(tmp_path / "i16.wav")], info[:1], np.array([0]), info["n_frames"][:1], dst, np.array([8]))
    assert np.array_equal(dst.numpy()[8:], i16) and not dst.numpy()[:8].any()
    with pytest.raises(Exception, match="nope.wav"):
        H.wav_probe([str(tmp_path / "nope.wav")])
    (tmp_path / "junk.wav").write_bytes(b"not a wav file at all")
    with pytest.raises(Exception, match="RIFF"):
        H.wav_probe([str(tmp_path / "junk.wav")])


def test_featurizer_paths_config_and_collate(tmp_path):
    """Newer featurizer layout (tts/parts/preprocessing/features.py:84-160): file naming, YAML with
    ``_target_`` / ``${...}`` in the reference's shape, padded collate."""
    import torch
    import yaml
    from roar_b200 import compute_features as cf
    from roar_b200 import featurizers as F
    audio_dir, feature_dir = tmp_path / "audio", tmp_path / "feat"
    e_rel = {"audio_filepath": "speaker1/audio1.wav"}
    e_abs = {"audio_filepath": str(audio_dir / "speaker2" / "a.flac")}
    assert F.get_feature_filepath(e_rel, audio_dir, feature_dir, "pitch") == feature_dir / "pitch" / "speaker1" / "audio1.pt"
    assert F.get_feature_filepath(e_abs, audio_dir, feature_dir, "mel") == feature_dir / "mel" / "speaker2" / "a.pt"
    cfg = {
        "sample_rate": 22050, "win_length": 1024, "hop_length": 256,
        "mel_feature": {"_target_": "roar.collections.tts.parts.preprocessing.features.MelSpectrogramFeaturizer",
                        "sample_rate": "${sample_rate}", "win_length": "${win_length}", "hop_length": "${hop_length}",
                        "mel_dim": 80, "lowfreq": 0, "highfreq": None},
        "pitch_feature": {"_target_": "roar.collections.tts.parts.preprocessing.features.PitchFeaturizer",
                          "sample_rate": "${sample_rate}", "win_length": "${win_length}", "hop_length": "${hop_length}",
                          "pitch_fmin": 60, "pitch_fmax": 640},
        "energy_feature": {"_target_": "roar.collections.tts.parts.preprocessing.features.EnergyFeaturizer",
                           "spec_featurizer": "${mel_feature}"},
        "featurizers": {"pitch": "${pitch_feature}", "energy": "${energy_feature}"},
    }
    p = tmp_path / "feature.yaml"
    p.write_text(yaml.safe_dump(cfg, sort_keys=False))
    fz = cf.load_featurizers(p)
    assert list(fz) == ["pitch", "energy"]
    assert isinstance(fz["pitch"], F.PitchFeaturizer) and fz["pitch"].cfg.pitch_fmax == 640.0
    assert fz["pitch"].cfg.pyin_hop_length == 256 and fz["pitch"].voiced_prob_name is None
    assert isinstance(fz["energy"].spec_featurizer, F.MelSpectrogramFeaturizer)
    mel_cfg = fz["energy"].spec_featurizer.cfg
    assert (mel_cfg.log_mode, mel_cfg.log_guard, mel_cfg.mel_norm, mel_cfg.spec_floor, mel_cfg.energy_mode) == \
        ("add", 1.0, None, 0.0, "features")
    with pytest.raises(ValueError):
        cf._resolve({"featurizers": {"x": {"_target_": "a.b.UnknownFeaturizer"}}})
    batch = [{"pitch": torch.ones(5), "voiced_mask": torch.ones(5, dtype=torch.bool)},
             {"pitch": torch.ones(3), "voiced_mask": torch.zeros(3, dtype=torch.bool)}]
    out = fz["pitch"].collate_fn(batch)
    assert out["pitch"].shape == (2, 5) and out["pitch"][1, 3:].sum() == 0 and out["voiced_mask"].dtype == torch.bool
    assert [len(b) for b in cf.make_batches([10, 50, 30, 20], 60)] == [1, 3]


def test_pitch_normalisation_and_stats_selection(tmp_path):
    import json
    import torch
    from oracle import stats as ostats
    from roar_b200 import dataset_utils as du
    pitch = np.array([0.0, 110.0, 0.0, 220.5, 98.7], dtype=np.float32)
    ref = ostats.normalize_pitch(pitch, 150.25, 42.5)
    got = du.normalize_pitch(torch.from_numpy(pitch.copy()), 150.25, 42.5).numpy()
    assert np.array_equal(got, ref) and got[0] == 0 and got[2] == 0
    stats = {"default": {"pitch_mean": 100.0, "pitch_std": 10.0}, "7": {"pitch_mean": 200.0, "pitch_std": 20.0}}
    p = tmp_path / "pitch_stats.json"
    p.write_text(json.dumps(stats))
    st = du.load_pitch_stats(p)
    assert du.select_pitch_stats({"speaker_id": 7}, pitch_stats=st) == (200.0, 20.0)
    assert du.select_pitch_stats({"speaker_id": 9}, pitch_stats=st) == (100.0, 10.0)
    assert du.select_pitch_stats({"speaker_id": 7}, 1.0, 2.0, st) == (1.0, 2.0)
    with pytest.raises(ValueError):
        du.select_pitch_stats({}, pitch_stats={"3": stats["7"]})
    with pytest.raises(ValueError):
        du.select_pitch_stats({})


def test_native_pt_writer_matches_torch_save(tmp_path):
    """roar_sup_pt_write_batch writes exactly the tensors the reference saves with torch.save: CPU float32,
    log_mel [1, n_mels, T], compact storage; readable by torch.load (weights_only or not) and a valid zip."""
    import zipfile
    import torch
    from roar_b200 import host_io as H
    n_mels, T = 4, [3, 5, 2, 70000]
    fo = np.concatenate([[0], np.cumsum(T)])
    base = torch.randn(int((n_mels + 1) * fo[-1]))
    offs, shapes, paths = [], [], []
    for i in range(len(T)):
        offs.append(n_mels * fo[i]); shapes.append((1, n_mels, T[i])); paths.append(str(tmp_path / "log_mel" / f"u{i}.pt"))
        offs.append(n_mels * fo[-1] + fo[i]); shapes.append((T[i],)); paths.append(str(tmp_path / "pitch" / f"u{i}.pt"))
    (tmp_path / "log_mel").mkdir(); (tmp_path / "pitch").mkdir()
    for threads in (1, 4):
        H.pt_write_batch(base, np.array(offs), shapes, paths, threads)
        for o, sh, p in zip(offs, shapes, paths):
            for wo in (True, False):
                t = torch.load(p, weights_only=wo)
                assert t.dtype == torch.float32 and tuple(t.shape) == sh and not t.is_cuda and t.is_contiguous()
                assert torch.equal(t.flatten(), base[o:o + t.numel()])
                assert t.untyped_storage().nbytes() == t.numel() * 4      # compact storage, not the whole batch
            assert zipfile.ZipFile(p).testzip() is None
        assert not [f for f in os.listdir(tmp_path / "pitch") if ".tmp" in f]
    # same payload as torch.save: identical tensors after a round trip through either writer
    torch.save(base[:12].view(1, 4, 3).clone(), tmp_path / "ref.pt")
    assert torch.equal(torch.load(tmp_path / "ref.pt"), torch.load(paths[0]))
    with pytest.raises(Exception):
        H.pt_write_batch(base, np.array([0]), [(3,)], [str(tmp_path / "missing_dir" / "x.pt")], 1)


def test_packed_cache_roundtrip_and_resample_plan(tmp_path):
    """Row N1 packed format (writer + reader) and the resampler's host-side plan against scipy's output length."""
    import torch
    from scipy.signal import resample_poly
    from roar_b200 import resample as R
    from roar_b200.dataset_utils import PackedCache, write_packed_batch
    base = torch.arange(1000, dtype=torch.float32)
    idx = []
    write_packed_batch(tmp_path, "r0_x_b000000", base, [0, 500, 900], [(1, 4, 10), (7,), (3,)],
                       ["log_mel/a", "pitch/a", "pitch/b"], idx)
    (tmp_path / "index_r0.jsonl").write_text("".join(idx))
    c = PackedCache(tmp_path)
    assert tuple(c.load("log_mel", "a").shape) == (1, 4, 10) and torch.equal(c.load("log_mel", "a").flatten(), base[:40])
    assert torch.equal(c.load("pitch", "a"), base[500:507]) and torch.equal(c.load("pitch", "b"), base[900:903])
    assert "pitch/b" in c and "energy/a" not in c and PackedCache.index_ids(tmp_path) == {"a", "b"}
    x = np.random.default_rng(0).standard_normal(10007).astype(np.float32)
    for a, b in ((44100, 22050), (48000, 22050), (16000, 22050), (22050, 16000)):
        up, down, taps, n_pre_pad, n_pre_remove = R.plan(a, b)
        y = resample_poly(x, up, down)
        assert len(y) == int(R.out_len(len(x), up, down)) and taps.dtype == np.float32 and len(taps) == 20 * max(up, down) + 1
        # the kernel's formula, evaluated here for a few outputs
        for n in (0, 1, len(y) // 2, len(y) - 1):
            c0 = (n + n_pre_remove) * down - n_pre_pad
            k_hi = min(c0 // up, len(x) - 1) if c0 >= 0 else -1
            k_lo = max(0, -(-(c0 - (len(taps) - 1)) // up))
            ks = np.arange(k_lo, k_hi + 1)
            assert abs(float((taps[c0 - ks * up].astype(np.float64) * x[ks]).sum()) - y[n]) <= 2e-6
    assert R.plan(22050, 22050)[:2] == (1, 1)
