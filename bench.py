#!/usr/bin/env python
"""bench.py -- sup-data extraction throughput (audio-seconds processed per second) on B200.

    python bench.py --gpus N --steps K --warmup W [--workload C2|C3|C4|C5] [--impl reference]

A "step" is one pass of the hot path over the workload (BASELINE.json `configs`):

  C2 (default, the headline)  LJSpeech-shaped 24 h manifest, 22.05 kHz, n_fft 1024 / hop 256 / 80 mels: log-mel +
      energy, pYIN f0 / voiced flag / voiced probability, beta-binomial prior, pitch-stat partials.  With N GPUs every
      rank processes its own 24 h shard (weak scaling).
  C3  multispeaker manifest (log-normal 0.5-20 s, 400 speakers), ONE seeded manifest dealt over the N ranks by the
      CLI's partitioner `shard_indices` (strong scaling: total work fixed), global + per-speaker pitch statistics.
  C4  HiFiTTS-shaped: 44.1 kHz, n_fft = win 2048, hop 512, 80 mels, fmax None, utterances of 10-30 s.
  C5  Conformer ASR preprocessor: `roar_fbank_forward`, 16 kHz, win 400 / hop 160 / n_fft 512, 80 mels, batch 256.
In every case the only collective is the all-reduce of the pitch-stat partials (SUM / MIN / MAX); C5 has none.

  value     audio-s/s with the audio already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the public host API with the input in pinned HOST memory as 16-bit PCM (what a wav
            corpus holds; float32 for C5): per step the input is copied host -> device, converted on the GPU, and every
            tensor the reference caches on disk plus the pitch statistics is copied back
  roofline  the dominant kernel (chain) against the CUDA-core FP peak and the measured HBM peak (SURVEY.md 8d:
            frac = max of the two), algorithmic flops / bytes only
  cpu_baseline  the reference-style CPU path (oracle port) on this box's host cores over a bounded sample of the same
            workload (rank 0, N = 1 only); `--impl reference` times that path alone with the same JSON shape
  cli_e2e   (C2, N = 1) disk -> `.pt`: `python -m roar_b200.extract_sup_data` run in-process on 16-bit wav files in tmpfs
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KERNEL_NAMES = ["tile_offsets", "stft_mel", "pyin_cmnd", "pyin_probs", "len_sort", "viterbi", "backtrack",
                "prior", "stats", "fbank_norm", "pyin_energy", "pcm16"]
METRIC = "audio-seconds processed/sec (sup-data extraction)"
SUP_TYPES = ["log_mel", "align_prior_matrix", "pitch", "voiced_mask", "p_voiced", "energy"]
CACHED = ("log_mel", "pitch", "voiced_mask", "p_voiced", "energy")

WORKLOADS = {
    "C2": dict(corpus="C2", n_utts=13100, scaling="weak", sup=dict(highfreq=8000.0),
               desc="C2: LJSpeech-shaped 24 h synthetic manifest, 22.05 kHz, n_fft 1024/hop 256/80 mels"),
    "C3": dict(corpus="C3", n_utts=72000, scaling="strong", sup=dict(highfreq=8000.0),
               desc="C3: multispeaker synthetic manifest (log-normal 0.5-20 s, 400 speakers), 22.05 kHz, n_fft 1024/"
                    "hop 256/80 mels, one manifest sharded over the ranks with global + per-speaker pitch statistics"),
    "C4": dict(corpus="C4", n_utts=4000, scaling="weak",
               sup=dict(sample_rate=44100, n_fft=2048, win_length=2048, hop_length=512, highfreq=None),
               desc="C4: HiFiTTS-shaped 44.1 kHz, n_fft 2048/hop 512/80 mels, fmax=None, utterances of 10-30 s"),
    "C5": dict(corpus="C5", n_utts=256, scaling="weak", sup=None,
               desc="C5: Conformer ASR preprocessor log-mel only, 16 kHz, n_fft 512/hop 160/80 mels, batch 256"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- shared description
def config_dict(wname, n_utts, world):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm."""
    from roar_b200 import synth
    w = WORKLOADS[wname]
    man = synth.corpus_manifest(w["corpus"], n_utts)
    hours = sum(u.duration for u in man) / 3600
    if wname == "C5":
        return {"workload": w["desc"], "batch": n_utts, "audio_hours_per_step_per_gpu": hours, "input": "float32 [B, Lmax]",
                "sup_data_types": ["log_mel"], "l2": "inputs larger than L2 (%.0f MB per batch)" % (max(u.n_samples for u in man) * n_utts * 4 / 1e6),
                "parallelism": f"one batch per rank x{world}, no collective"}
    per = "per_gpu" if w["scaling"] == "weak" else "total"
    return {"workload": w["desc"], f"utterances_{per}": n_utts, f"audio_hours_{per}": hours,
            "sup_data_types": SUP_TYPES, "input": "16-bit PCM (float32 x/2^15 values resident in HBM for `value`; int16 over PCIe for `e2e`)",
            "l2": "inputs larger than L2 (%.1f GB audio per pass)" % (sum(u.n_samples for u in man) * 4 / 1e9 / (world if w["scaling"] == "strong" else 1)),
            "parallelism": (f"utterance shards x{world} (every rank its own manifest), one all-reduce of pitch partials" if w["scaling"] == "weak"
                            else f"one manifest dealt over {world} rank(s) by shard_indices (LPT on durations), one all-reduce of pitch partials")}


def oracle_kwargs(wname):
    sup = WORKLOADS[wname]["sup"] or {}
    kw = {}
    if "n_fft" in sup:
        kw.update(n_fft=sup["n_fft"], hop_length=sup["hop_length"], win_length=sup["win_length"])
    if "highfreq" in sup:
        kw["fmax"] = sup["highfreq"]
    return kw


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_run(wname, steps, warmup, per_step=None):
    """The reference-style CPU path, `steps` timed bounded samples of the workload's manifest.
    -> (audio_seconds, wall_seconds, cores, per_step, description)"""
    cores = os.cpu_count() or 1
    if wname == "C5":
        return cpu_reference_fbank(steps, warmup, per_step or max(4, min(32, cores)))
    from oracle import extract as oextract
    w = WORKLOADS[wname]
    per_step = per_step or cores
    kw = oracle_kwargs(wname)
    for _ in range(warmup):
        oextract.timed_cpu_extraction(w["corpus"], 0, min(per_step, cores), cores, **kw)
    audio = wall = 0.0
    for s in range(steps):
        a, t, _ = oextract.timed_cpu_extraction(w["corpus"], s * per_step, per_step, cores, **kw)
        audio += a
        wall += t
    return audio, wall, cores, per_step, ("oracle port of TTSDataset.__getitem__ (3 STFTs, dense-Viterbi pyin, float32 "
                                          "prior), one process per core")


def cpu_reference_fbank(steps, warmup, rows):
    """FilterbankFeatures.forward on the CPU (the oracle's restatement: the same torch calls) on `rows` rows of the
    C5 batch per step, all host cores as torch intra-op threads."""
    import torch
    from oracle import fbank as ofbank
    from roar_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    man = synth.corpus_manifest("C5", rows)
    wavs = [synth.synth_utterance(5, u.utt_id, u.n_samples, 16000, u.speaker) for u in man]
    lens = np.array([len(x) for x in wavs], dtype=np.int64)
    x = np.zeros((rows, int(lens.max())), dtype=np.float32)
    for i, wv in enumerate(wavs):
        x[i, :len(wv)] = wv
    orc = ofbank.FilterbankFeaturesOracle(sample_rate=16000, n_window_size=400, n_window_stride=160, nfilt=80, n_fft=512)
    for _ in range(warmup):
        orc.forward(x, lens)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.forward(x, lens)
    wall = time.perf_counter() - t0
    return float(lens.sum()) / 16000 * steps, wall, cores, rows, "oracle restatement of FilterbankFeatures.forward (torch.stft on the CPU, all cores as intra-op threads)"


def bind_cores(local_rank, world):
    """Give every rank its own slice of the host cores (staging threads, launch thread): on a box where all GPUs sit on
    one NUMA node the ranks otherwise migrate over the same cores."""
    if world <= 1 or not hasattr(os, "sched_getaffinity"):
        return None
    cores = sorted(os.sched_getaffinity(0))
    per = len(cores) // world
    if per < 1:
        return None
    mine = cores[local_rank * per:(local_rank + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


# --------------------------------------------------------------------------------------------- C5: fbank forward
def run_fbank(args, rank, world, local_rank, cpu_baseline):
    import torch
    import torch.distributed as dist
    from roar_b200 import synth
    from roar_b200.features import AudioToMelSpectrogramPreprocessor

    dev = torch.device("cuda", local_rank)
    B = args.n_utts or WORKLOADS["C5"]["n_utts"]
    man = synth.corpus_manifest("C5", B)
    lens_h = np.array([u.n_samples for u in man], dtype=np.int64)
    Lmax = int(lens_h.max())
    g = torch.Generator(device=dev); g.manual_seed(5 + 1000 * rank)
    x = 0.1 * torch.randn(B, Lmax, generator=g, device=dev)
    t = torch.arange(Lmax, device=dev)[None, :]
    x = x * torch.sin(t * (2 * np.pi * 180.0 / 16000)) * (t < torch.as_tensor(lens_h, device=dev)[:, None])
    lens = torch.as_tensor(lens_h, device=dev)
    pre = AudioToMelSpectrogramPreprocessor(sample_rate=16000, window_size=0.025, window_stride=0.01, features=80,
                                            n_fft=512, dither=0.0).to(dev).eval()
    audio_s = float(lens_h.sum()) / 16000
    h = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup)):
        out, out_len = pre(input_signal=x, length=lens)
    h = pre.featurizer._h
    lib = pre.featurizer._lib
    barrier()
    lib.roar_sup_set_profiling(h, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out, out_len = pre(input_signal=x, length=lens)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    from roar_b200 import _lib
    kms = (ctypes.c_double * _lib.N_KERNEL_IDS)()
    kcnt = (ctypes.c_int64 * _lib.N_KERNEL_IDS)()
    lib.roar_sup_profile_read(h, kms, kcnt, 1)
    lib.roar_sup_set_profiling(h, 0)
    # e2e: pinned host float32 in, features + lengths back
    hx = torch.empty(B, Lmax, dtype=torch.float32, pin_memory=True); hx.copy_(x)
    hl = torch.from_numpy(lens_h).pin_memory()
    ho = torch.empty(tuple(out.shape), dtype=torch.float32, pin_memory=True)
    hol = torch.empty(B, dtype=torch.int64, pin_memory=True)

    def step_e2e():
        dx = hx.to(dev, non_blocking=True)
        dl = hl.to(dev, non_blocking=True)
        o, ol = pre(input_signal=dx, length=dl)
        ho.copy_(o, non_blocking=True)
        hol.copy_(ol, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    wall = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    e2e_s = float(wall.item())
    if rank != 0:
        return None
    hbm_peak, peak_src = measured_peaks()
    T = int(out.shape[2])
    frames = int((1 + lens_h // 160).sum())
    per = {k: kms[i] / max(1, args.steps) for i, k in enumerate(KERNEL_NAMES)}
    alg_b = 4 * B * Lmax + 4 * 80 * B * T          # the dense [B, Lmax] input read once, [B, 80, Tpad] written once
    flops = B * (1 + Lmax // 160) * (2.5 * 512 * 9 + 3 * 257 + 2 * 500 + 80 + 3 * 80)
    dom_ms = per["stft_mel"] + per["fbank_norm"]
    fp_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    ach_gbs = alg_b / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    ach_tf = flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    hb, fpf = ach_gbs / hbm_peak, ach_tf / fp_peak
    roofline = {"bound": "hbm" if hb >= fpf else "fp", "kernel": "k_stft_mel + k_fbank_normalize",
                "achieved": ach_gbs if hb >= fpf else ach_tf, "peak": hbm_peak if hb >= fpf else fp_peak,
                "unit": "GB/s" if hb >= fpf else "TFLOP/s", "frac": max(hb, fpf), "traffic": None,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hb, "peak_source": peak_src},
                "fp": {"achieved": ach_tf, "peak": fp_peak, "unit": "TFLOP/s", "frac": fpf,
                       "peak_source": "theoretical FP32 CUDA-core peak 148 SM x 128 lanes x 2 x 1.965 GHz (no FP32 entry in MEASURED_PEAKS.json)"},
                "algorithmic_bytes_per_launch": int(alg_b), "algorithmic_flops_per_launch": float(flops), "ms_per_launch": dom_ms}
    value = world * audio_s * args.steps / (total_ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict("C5", B, world),
            "x_realtime_per_gpu": value / world, "gpu_launches": int(sum(kcnt)), "clocks": clocks,
            "kernels_ms_per_step": {k: round(v, 4) for k, v in per.items() if v > 0}, "roofline": roofline,
            "e2e": {"value": world * audio_s * args.steps / e2e_s, "unit": "audio-s/s", "h2d_bytes_per_step": B * Lmax * 4 + B * 8,
                    "d2h_bytes_per_step": int(out.numel()) * 4 + B * 8, "ms_per_step": 1e3 * e2e_s / args.steps},
            "frames_per_step": frames}
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    return line


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--n-utts", type=int, default=0, help="utterances of the workload's manifest (0 = its default: "
                    "C2 13100 per rank, C3 72000 in total = 100 h, C4 4000 per rank = 22 h, C5 batch 256)")
    ap.add_argument("--chunk-utts", type=int, default=0, help="utterances per device-resident call (0 = C2: all, C3: 13000, C4: 2000)")
    ap.add_argument("--e2e-chunk-utts", type=int, default=0, help="largest streamed host chunk (utterances; 0 = C2/C3: 2200, C4: 500)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-streams", type=int, default=2, choices=(1, 2),
                    help="extractor objects / compute streams the streamed chunks alternate over (2: the occupancy "
                         "tail of one chunk's kernels is filled by the next chunk's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cli-utts", type=int, default=8000, help="C2, N = 1: utterances of the disk -> .pt CLI leg; 0 = skip")
    ap.add_argument("--cli-dir", default="/dev/shm", help="where the CLI leg puts its wav files and cache (tmpfs)")
    ap.add_argument("--no-affinity", action="store_true", help="do not give each rank its own slice of the host cores")
    ap.add_argument("--e2e-f32", action="store_true", help="diagnostic: ship float32 audio host -> device (4 bytes per sample) instead of 16-bit PCM")
    ap.add_argument("--e2e-run-ahead", type=int, default=0,
                    help="e2e: how many chunks beyond one per lane the launching thread may queue (0 = a chunk is issued "
                         "once the previous chunk of its lane has finished: measured 20 %% faster than queueing the whole "
                         "step up front, which lets the two lanes fall into lock-step on the same kernels); -1 = unbounded")
    ap.add_argument("--trace-e2e", action="store_true", help="diagnostic: per-step wall times of the e2e leg on stderr")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wname = args.workload
    W = WORKLOADS[wname]
    n_manifest = args.n_utts or W["n_utts"]

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        audio, wall, cores, per_step, how = cpu_reference_run(wname, args.steps, min(args.warmup, 1), None)
        v = audio / wall
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": W["scaling"], "vs_baseline": None, "dtype": "f32" if wname == "C5" else "f32+f64", "data": "synthetic",
            "config": config_dict(wname, n_manifest, args.gpus),
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} x {per_step} utterances of the workload's manifest per step, {how}"},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ CPU baseline first (before CUDA init)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = max(8, min(4 * cores, 128)) if wname in ("C2", "C3") else max(4, min(cores, 32))
        audio, wall, cores, per_step, how = cpu_reference_run(wname, 1, 0, n)
        cpu_baseline = {"value": audio / wall, "unit": "audio-s/s", "cores": cores, "kind": "port",
                        "sample": f"first {per_step} utterances ({audio:.0f} audio-s) of the workload's manifest, {how}"}

    import torch
    import torch.distributed as dist

    bound = None if args.no_affinity else bind_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if wname == "C5":
        line = run_fbank(args, rank, world, local_rank, cpu_baseline)
        if line is not None:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    from roar_b200 import _lib, synth
    from roar_b200.config import SupConfig
    from roar_b200.extract_sup_data import merge_partials, shard_indices
    from roar_b200.extractor import SupDataExtractor, finalize_pitch_stats

    scfg = SupConfig(**W["sup"])
    SR, HOP, N_MELS = scfg.sample_rate, scfg.hop, scfg.n_mels
    ex = SupDataExtractor(scfg, dev)
    if W["scaling"] == "strong":
        # ONE seeded manifest; the CLI's partitioner deals it over the ranks; each rank synthesises only its shard
        full = synth.corpus_manifest(W["corpus"], n_manifest)
        mine = shard_indices([u.duration for u in full], world, rank)
        man, audio, offs, lens = synth.synth_corpus_device_stateless(W["corpus"], dev, indices=mine, n_utts=n_manifest)
        total_audio_s = sum(u.n_samples for u in full) / SR
    else:
        # every rank its own shard (same manifest shape, different seed); the waveform is quantised to 16 bits
        key = f"{W['corpus']}_rank"
        synth.CORPORA[key] = dict(synth.CORPORA[W["corpus"]], seed=synth.CORPORA[W["corpus"]]["seed"] + 1000 * rank)
        man, audio, offs, lens = synth.synth_corpus_device(key, dev, n_utts=n_manifest)
        audio = torch.clamp(torch.round(audio * 32767.0), -32768, 32767) * (1.0 / 32768.0)
        total_audio_s = None
    offs_h = offs.cpu().numpy()
    lens_h = lens.cpu().numpy().astype(np.int64)
    text_lens = np.array([u.text_len for u in man], dtype=np.int32)
    speakers = np.array([u.speaker for u in man], dtype=np.int32)
    n_spk = synth.CORPORA[W["corpus"]]["n_speakers"]
    per_speaker = wname == "C3"
    audio_s = float(lens_h.sum()) / SR
    if total_audio_s is None:
        total_audio_s = world * audio_s
    n_utts = len(man)
    chunk = args.chunk_utts or {"C2": n_utts, "C3": 13000, "C4": 2000}[wname]
    chunk = max(1, min(chunk, n_utts))
    bounds = [(a, min(a + chunk, n_utts)) for a in range(0, n_utts, chunk)]

    def chunk_batch(e, a, b, src=None, lo=None):
        lo = int(offs_h[a]) if lo is None else lo
        hi = int(offs_h[b - 1] + (lens_h[b - 1] + 3) // 4 * 4)
        return e.batch_from_device(audio[lo:hi] if src is None else src, offs_h[a:b] - lo, lens_h[a:b])

    batches = [chunk_batch(ex, a, b) for a, b in bounds]
    n_groups = 1 + (n_spk if per_speaker else 0)
    stats = ex.new_pitch_partials(n_groups)

    def reduce_stats(st):
        """The one exchange step: SUM on (sum, sumsq, count), MIN / MAX on the extremes."""
        if world == 1:
            return st
        red = st.clone()
        sums, mn, mx = red[:, :3].contiguous(), red[:, 3].contiguous(), red[:, 4].contiguous()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        red[:, :3], red[:, 3], red[:, 4] = sums, mn, mx
        return red

    def accumulate(e, out, a, b, st):
        if per_speaker:          # group 0 = all, 1 + s = speaker s (compute_speaker_stats.py:105-132)
            e.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], 1 + speakers[a:b], n_groups, st)
            e.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], np.zeros(b - a, np.int32), n_groups, st)
        else:
            e.pitch_partials(out["pitch"], st)

    def step_device():
        ex.lib.roar_sup_pitch_partials_init(ex._h, ctypes.c_void_p(stats.data_ptr()), n_groups,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        for (a, b), bt in zip(bounds, batches):
            out = ex.extract(bt, text_lens=text_lens[a:b])
            accumulate(ex, out, a, b, stats)
            del out
        return reduce_stats(stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ex.kernel_launches
    ex.lib.roar_sup_set_profiling(ex._h, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        final = step_device()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = ex.kernel_launches - launches0
    kms = (ctypes.c_double * _lib.N_KERNEL_IDS)()
    kcnt = (ctypes.c_int64 * _lib.N_KERNEL_IDS)()
    ex.lib.roar_sup_profile_read(ex._h, kms, kcnt, 1)
    ex.lib.roar_sup_set_profiling(ex._h, 0)
    pstats = finalize_pitch_stats(final)

    # ------------------------------------------------------------------ e2e: host buffers (16-bit PCM), copies inside
    e2e = None
    if not args.no_e2e:
        total_samples = int(audio.numel())
        if args.e2e_f32:
            host_pcm = torch.empty(total_samples, dtype=torch.float32, pin_memory=True)
            host_pcm.copy_(audio)
        else:
            host_pcm = torch.empty(total_samples, dtype=torch.int16, pin_memory=True)
            host_pcm.copy_((audio * 32768.0).to(torch.int16))         # exact: the audio is x / 2^15
        frames_all = int((1 + lens_h // HOP).sum())
        host_out = {k: torch.empty(frames_all * (N_MELS if k == "log_mel" else 1), dtype=torch.float32, pin_memory=True)
                    for k in CACHED}
        host_stats = torch.empty(n_groups, 5, dtype=torch.float64, pin_memory=True)
        # streamed host chunks: small first and last chunks (the first H2D and the last D2H copy are the only
        # ones nothing hides behind), large ones in between (short occupancy tails in the kernels)
        ec = args.e2e_chunk_utts or {"C2": 2200, "C3": 2200, "C4": 500}[wname]
        ramp = [max(64, ec // 8), max(64, ec // 4), max(64, ec // 2)]
        if n_utts <= 2 * sum(ramp):
            sizes = [n_utts]
        else:
            mid = n_utts - 2 * sum(ramp)
            k = (mid + ec - 1) // ec
            sizes = ramp + [mid // k + (1 if i < mid % k else 0) for i in range(k)] + ramp[::-1]
        eb, a = [], 0
        for sz in sizes:
            eb.append((a, a + sz))
            a += sz
        assert a == n_utts
        copy_stream = torch.cuda.Stream(dev)
        out_stream = torch.cuda.Stream(dev)
        frame_cum = np.concatenate([[0], np.cumsum(1 + lens_h // HOP)])
        # chunks alternate over `--e2e-streams` extractor objects, each with its own compute stream, handle and
        # workspace: consecutive chunks are independent, so the tail of one chunk's kernels overlaps the head
        # of the next chunk's
        n_lane = args.e2e_streams
        lanes = [ex] + [SupDataExtractor(scfg, dev) for _ in range(n_lane - 1)]
        lane_streams = [torch.cuda.Stream(dev) for _ in range(n_lane)]

        def step_e2e():
            t_step = time.perf_counter()
            main = torch.cuda.current_stream()
            sts = []
            for e, cs in zip(lanes, lane_streams):
                cs.wait_stream(main)
                with torch.cuda.stream(cs):
                    sts.append(e.new_pitch_partials(n_groups))
            staged = {}

            def stage(i):
                a, b = eb[i]
                lo = int(offs_h[a])
                hi = int(offs_h[b - 1] + (lens_h[b - 1] + 3) // 4 * 4)
                with torch.cuda.stream(copy_stream):
                    d = host_pcm[lo:hi].to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                staged[i] = (d, ev, lo)

            stage(0)
            pending = []
            dones = []
            for i, (a, b) in enumerate(eb):
                back = i - n_lane - args.e2e_run_ahead
                if args.e2e_run_ahead >= 0 and back >= 0:
                    dones[back].synchronize()      # bounded run-ahead: keeps the lanes staggered (see --e2e-run-ahead)
                d, ev, lo = staged.pop(i)
                if i + 1 < len(eb):
                    stage(i + 1)
                e, cs = lanes[i % n_lane], lane_streams[i % n_lane]
                with torch.cuda.stream(cs):
                    cs.wait_event(ev)
                    d.record_stream(cs)
                    f32 = d if args.e2e_f32 else e.pcm16_to_f32(d)      # the decoder's int -> float step, on the GPU
                    bt = chunk_batch(e, a, b, src=f32, lo=lo)
                    out = e.extract(bt, text_lens=text_lens[a:b])
                    accumulate(e, out, a, b, sts[i % n_lane])
                    done = torch.cuda.Event()
                    done.record(cs)
                    dones.append(done)
                f_lo, f_hi = int(frame_cum[a]), int(frame_cum[b])
                with torch.cuda.stream(out_stream):
                    out_stream.wait_event(done)
                    for k in host_out:
                        m = N_MELS if k == "log_mel" else 1
                        host_out[k][m * f_lo:m * f_hi].copy_(out[k], non_blocking=True)
                pending.append((out, d, f32, bt))     # keep device buffers alive until the copies are done
            if args.trace_e2e:
                sys.stderr.write(f"[rank {rank}] e2e issue done after {1e3 * (time.perf_counter() - t_step):.1f} ms\n")
            for cs in lane_streams:
                main.wait_stream(cs)
            st = sts[0]
            for o in sts[1:]:
                st = merge_partials(st, o)
            st = reduce_stats(st)                     # the same SUM / MIN / MAX exchange as the device leg
            host_stats.copy_(st, non_blocking=True)
            out_stream.synchronize()
            main.synchronize()
            del pending

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ts = time.perf_counter()
            step_e2e()
            if args.trace_e2e:
                sys.stderr.write(f"[rank {rank}] e2e step {1e3 * (time.perf_counter() - ts):.1f} ms\n")
        barrier()
        wall = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e_s = float(wall.item())
        e2e = {"value": total_audio_s * args.steps / e2e_s, "unit": "audio-s/s",
               "h2d_bytes_per_step": total_samples * (4 if args.e2e_f32 else 2),
               "d2h_bytes_per_step": frames_all * 4 * (N_MELS + 4) + 40 * n_groups,
               "bytes_note": "per rank; input travels as 16-bit PCM and is converted on the GPU (roar_sup_pcm16_to_f32)",
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "pitch_mean": finalize_pitch_stats(host_stats)["pitch_mean"],
               "core_affinity": (f"{len(bound)} cores per rank" if bound else "unbound")}
        del host_pcm, host_out, lanes

    # ------------------------------------------------------------------ disk -> .pt through the CLI (rows a1-a3, N1)
    cli_e2e = None
    if rank == 0 and world == 1 and wname == "C2" and args.cli_utts > 0 and os.path.isdir(args.cli_dir):
        cli_e2e = cli_leg(args, audio, offs_h, lens_h, SR)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel (chain)
    hbm_peak, peak_src = measured_peaks()
    fp_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    fp_src = "theoretical FP32 CUDA-core peak 148 SM x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no FP32 entry)"
    per = {k: (kms[i] / max(1, args.steps)) for i, k in enumerate(KERNEL_NAMES)}       # ms per step
    frames = int((1 + lens_h // HOP).sum())
    samples = int(lens_h.sum())
    prior_elems = int(((1 + lens_h // HOP) * text_lens.astype(np.int64)).sum())
    # algorithmic HBM bytes per step of each kernel (inputs it must read + final outputs it must write;
    # scratch excluded) -- DESIGN.md section 4
    alg = {
        "stft_mel": 4 * samples + 4 * (N_MELS + 1) * frames,
        "pyin_energy": 4 * samples,
        "pyin_cmnd": 4 * samples,
        "pyin_probs": 4 * frames,
        "viterbi": 0,
        "backtrack": 8 * frames,
        "prior": 4 * prior_elems,
        "stats": 4 * frames,
    }
    # algorithmic flops per frame, cheapest standard algorithm
    F, n_fft = scfg.pyin_frame, scfg.n_fft
    nnz = int((ex.mel_filterbank() != 0).sum())
    fl = {
        "stft_mel": 2.5 * n_fft * np.log2(n_fft) + 3 * (n_fft // 2 + 1) + 2 * (n_fft // 2 + 1) + 2 * nnz + N_MELS,
        "pyin_cmnd": 3 * 2.5 * F * np.log2(F) + 6 * (F // 2 + 1) + 4 * F,          # via three real FFTs, not the direct form
        "pyin_probs": 10000.0 * (ex.max_period - ex.min_period + 1) / 329.0,
        "viterbi": 2.0 * (2 * ex.n_pitch_bins) * (2 * ex.transition_width + 1),       # banded max-plus step
    }
    chain = ("pyin_energy", "pyin_cmnd", "pyin_probs", "viterbi", "backtrack")
    dom = max(("stft_mel",) + chain + ("prior",), key=lambda k: per[k])
    n_launch = max(1, len(bounds))
    if dom in chain:
        # the pYIN chain is one logical kernel split at scratch hand-offs: audio in (4*hop B/frame), f0 / flag / prob out
        dom_alg = 4 * samples + 12 * frames
        dom_fl = (fl["pyin_cmnd"] + fl["pyin_probs"] + fl["viterbi"]) * frames
        dom_ms = sum(per[k] for k in chain)
        dom_name = "pyin chain (pyin_energy + pyin_cmnd + pyin_probs + viterbi + backtrack); slowest member: " + dom
        members = list(chain)
    else:
        dom_alg, dom_fl, dom_ms, dom_name, members = alg[dom], fl.get(dom, 0.0) * frames, per[dom], dom, [dom]
    ach_gbs = dom_alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    ach_tf = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    # measured DRAM traffic of the same kernels (one ncu --set full capture of this workload, committed under
    # profiles/): scratch hand-offs between the chain's kernels make it larger than the algorithmic bytes
    traffic = traffic_src = None
    if wname == "C2" and n_utts == 13100 and len(bounds) == 1:
        for tname in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if not os.path.exists(tpath):
                continue
            with open(tpath) as f:
                tk = json.load(f)["kernels"]
            names = {"pyin_energy": "k_pyin_energy", "pyin_cmnd": "k_pyin_cmnd", "pyin_probs": "k_pyin_probs",
                     "viterbi": "k_pyin_viterbi51", "backtrack": "k_pyin_backtrack", "stft_mel": "k_stft_mel", "prior": "k_align_prior"}
            vals = [tk.get(names[m], {}) for m in members]
            if all(v.get("dram_read_bytes") is not None and v.get("dram_write_bytes") is not None for v in vals):
                traffic = int(sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in vals))
                traffic_src = "profiles/" + tname
                break
    hb, fpf = ach_gbs / hbm_peak, ach_tf / fp_peak
    roofline = {"bound": "fp" if fpf >= hb else "hbm", "kernel": dom_name,
                "achieved": ach_tf if fpf >= hb else ach_gbs, "peak": fp_peak if fpf >= hb else hbm_peak,
                "unit": "TFLOP/s" if fpf >= hb else "GB/s", "frac": max(hb, fpf), "traffic": traffic, "traffic_source": traffic_src,
                "fp": {"achieved": ach_tf, "peak": fp_peak, "unit": "TFLOP/s", "frac": fpf, "peak_source": fp_src},
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hb, "peak_source": peak_src},
                "algorithmic_bytes_per_launch": dom_alg // n_launch, "algorithmic_flops_per_launch": dom_fl / n_launch,
                "ms_per_launch": dom_ms / n_launch,
                "note": "SURVEY.md 8d: frac = max(algorithmic bytes / HBM peak, algorithmic flops / FP peak); flops are the "
                        "cheapest standard algorithm's (FFT-form CMND, banded Viterbi), the kernels execute more (direct "
                        "float64 autocorrelation) -- the chain is CUDA-core issue / FP64-pipe bound, not HBM bound"}
    fp64_file = os.path.join(ROOT, "profiles", "r2_fp64_peak.json")
    if os.path.exists(fp64_file):      # measured FP64 peak of this GPU model (micro-benchmark): what bounds K2a's direct autocorrelation
        with open(fp64_file) as f:
            fp64 = json.load(f)
        # executed DFMA of the direct block autocorrelation: one block of `hop` samples per frame plus one per 15-frame
        # tile, (max_period + 1) lags rounded up to 11 per lane
        lags_exec = -(-(ex.max_period + 1) // 11) * 11
        executed = (2.0 * (frames * 16.0 / 15.0 + 0.5 * n_utts) * lags_exec * scfg.pyin_hop
                    if scfg.pyin_frame == 4 * scfg.pyin_hop else None)
        roofline["fp64"] = {"peak_tflops_measured": fp64["dfma_tflops"], "peak_source": "profiles/r2_fp64_peak.json (" + fp64["source"] + ")",
                            "k_pyin_cmnd_executed_tflops": (executed / (per.get("pyin_cmnd", 0.0) * 1e-3) / 1e12) if executed and per.get("pyin_cmnd") else None,
                            "note": "executed (not algorithmic) float64 multiply-adds of the direct block autocorrelation over the kernel time"}
    roofline_kernels = {}
    for k, b in alg.items():
        if per.get(k, 0) > 0:
            gbs = b / (per[k] * 1e-3) / 1e9
            tf = fl.get(k, 0.0) * frames / (per[k] * 1e-3) / 1e12
            roofline_kernels[k] = {"ms_per_step": round(per[k], 4), "algorithmic_bytes": int(b), "achieved_gbs": round(gbs, 2),
                                   "hbm_frac": round(gbs / hbm_peak, 5), "achieved_tflops": round(tf, 3), "fp_frac": round(tf / fp_peak, 5)}
    path_bytes = 4 * samples + 4 * (N_MELS + 4) * frames + 4 * prior_elems
    path_flops = sum(fl.values()) * frames
    step_ms = total_ms / args.steps
    value = total_audio_s * args.steps / (total_ms * 1e-3)
    cfg = config_dict(wname, n_manifest, world)
    line = {
        "metric": METRIC, "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": W["scaling"], "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": cfg,
        "x_realtime_per_gpu": value / world,
        "gpu_launches": launches, "clocks": clocks,
        "kernels_ms_per_step": {k: round(v, 4) for k, v in per.items() if v > 0},
        "kernels_note": "CUDA-event brackets on each kernel's launching stream inside the timed region (rank 0); the spectral "
                        "kernels (stft_mel, prior) run on a side stream concurrently with the pYIN chain, so a bracket "
                        "can include time queued behind the other stream and the brackets can sum to more than ms_per_step",
        "roofline": roofline,
        "roofline_kernels": roofline_kernels,
        "roofline_path": {"achieved_gbs": path_bytes / (step_ms * 1e-3) / 1e9, "hbm_frac": path_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak,
                          "achieved_tflops": path_flops / (step_ms * 1e-3) / 1e12, "fp_frac": path_flops / (step_ms * 1e-3) / 1e12 / fp_peak,
                          "algorithmic_bytes_per_step": path_bytes, "algorithmic_flops_per_step": path_flops,
                          "note": "rank 0's shard over the step time"},
        "pitch_stats": pstats,
        "rank0": {"utterances": n_utts, "audio_hours": audio_s / 3600, "chunks_per_step": len(bounds)},
    }
    if per_speaker:
        fin = final.cpu().numpy()
        line["speakers_with_stats"] = int((fin[1:, 2] > 1).sum())
    if e2e is not None:
        line["e2e"] = e2e
    if cli_e2e is not None:
        line["cli_e2e"] = cli_e2e
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cli_leg(args, audio, offs_h, lens_h, sr):
    """disk -> .pt: 16-bit wav files of the manifest head in tmpfs through `roar_b200.extract_sup_data.run` (the
    drop-in CLI, in-process), all five cached types.  Reported: audio-s/s over the streaming part (first decode
    submitted -> last cache file renamed) and over the whole call (manifest, extractor set-up, statistics)."""
    import shutil
    import struct
    import tempfile
    import torch
    from roar_b200 import extract_sup_data as X
    n = min(args.cli_utts, len(lens_h))
    root = tempfile.mkdtemp(prefix="roar_cli_", dir=args.cli_dir)
    try:
        wav_dir = os.path.join(root, "wavs")
        os.makedirs(wav_dir)
        rows = []
        hi = int(offs_h[n - 1] + lens_h[n - 1])
        pcm = (audio[:hi] * 32768.0).to(torch.int16).cpu().numpy()
        for i in range(n):
            x = pcm[int(offs_h[i]):int(offs_h[i] + lens_h[i])]
            p = os.path.join(wav_dir, f"utt{i:06d}.wav")
            with open(p, "wb") as f:
                f.write(b"RIFF" + struct.pack("<I", 36 + 2 * len(x)) + b"WAVEfmt " +
                        struct.pack("<IHHIIHH", 16, 1, 1, sr, 2 * sr, 2, 16) + b"data" + struct.pack("<I", 2 * len(x)))
                f.write(x.tobytes())
            rows.append(json.dumps({"audio_filepath": p, "duration": len(x) / sr, "text": "synthetic"}))
        mf = os.path.join(root, "train.json")
        with open(mf, "w") as f:
            f.write("\n".join(rows) + "\n")
        del pcm
        cfg = X.load_config([f"manifest_filepath={mf}", f"sup_data_path={os.path.join(root, 'sup')}",
                             "sup_data_types=[align_prior_matrix,pitch,voiced_mask,p_voiced,energy,log_mel]"])
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            res = X.run(cfg)
        run = res["run"]
        files = sum(len(os.listdir(os.path.join(root, "sup", t))) for t in CACHED)
        return {"value": run["audio_seconds"] / run["stream_seconds"], "unit": "audio-s/s", "utterances": run["utterances"],
                "audio_seconds": run["audio_seconds"], "stream_seconds": run["stream_seconds"], "total_seconds": run["seconds"],
                "value_incl_setup": run["audio_seconds"] / run["seconds"], "cache_files": files,
                "decode_threads": run["decode_threads"], "writer_threads": run["writer_threads"], "batches": run["batches"],
                "stage_seconds": run.get("stage_seconds"),
                "filesystem": args.cli_dir, "pitch_mean": res["pitch_mean"],
                "note": "16-bit wav files -> native decode threads -> pinned ring -> H2D int16 -> kernels -> D2H -> native "
                        ".pt writer (5 files per utterance); stream_seconds = first decode submitted -> last file renamed"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
