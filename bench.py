#!/usr/bin/env python
"""bench.py -- sup-data extraction throughput (audio-seconds processed per second) on B200.

A "step" is one pass of the whole hot path (log-mel + energy, pYIN f0 / voiced flag / voiced
probability, beta-binomial prior, pitch-stat partials [+ one all-reduce when N > 1]) over the
BASELINE.json config-2 workload: an LJSpeech-shaped 24 h synthetic manifest (13 100 utterances,
22.05 kHz, n_fft 1024 / hop 256 / 80 mels, 0-8000 Hz, pitch C2-C7).  With N GPUs every rank
processes its own 24 h shard (weak scaling, no data-path collective; one all-reduce of the five
pitch-stat partials per step).

  value     audio-s/s with the audio already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the public host API with the audio in pinned HOST memory: per step the
            packed audio is copied host->device and every tensor the reference caches on disk
            (log_mel, pitch, voiced_mask, p_voiced, energy) plus the pitch statistics is copied back
  roofline  the dominant kernel: its algorithmic HBM bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline  the reference-style CPU path (oracle port, dense Viterbi, 3 STFTs per utterance) on
            this box's host cores over a bounded sample of the same manifest (rank 0, N = 1 only)

`--impl reference` times that CPU path alone and prints the same JSON line shape.
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
                "prior", "stats", "fbank_norm", "pyin_energy"]
SR, HOP, N_MELS = 22050, 256, 80
WORKLOAD = "C2: LJSpeech-shaped 24 h synthetic manifest, 22.05 kHz, n_fft 1024/hop 256/80 mels"


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


def cpu_reference_run(steps, warmup, per_step=None):
    """The reference-style CPU extraction, `steps` timed bounded samples of the config-2 manifest."""
    from oracle import extract as oextract
    cores = os.cpu_count() or 1
    per_step = per_step or cores
    first = 0
    for _ in range(warmup):
        oextract.timed_cpu_extraction("C2", first, min(per_step, cores), cores)
    audio = wall = 0.0
    for s in range(steps):
        a, w, p = oextract.timed_cpu_extraction("C2", first + s * per_step, per_step, cores)
        audio += a
        wall += w
    return audio, wall, cores, per_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-utts", type=int, default=13100, help="utterances of the config-2 manifest per rank")
    ap.add_argument("--chunk-utts", type=int, default=0, help="utterances per device-resident call (0 = all)")
    ap.add_argument("--e2e-chunk-utts", type=int, default=2200, help="largest streamed host chunk (utterances)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-streams", type=int, default=2, choices=(1, 2),
                    help="extractor objects / compute streams the streamed chunks alternate over (2: the occupancy "
                         "tail of one chunk's kernels is filled by the next chunk's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cache-write-utts", type=int, default=512,
                    help="also time writing the .pt cache (5 files per utterance) for this many utterances; 0 = skip")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        audio, wall, cores, per_step = cpu_reference_run(args.steps, min(args.warmup, 1), None)
        v = audio / wall
        print(json.dumps({
            "impl": "reference", "metric": "audio-seconds processed/sec (sup-data extraction)", "value": v,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{per_step} utterances per step"},
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} x {per_step} utterances of the config-2 manifest, "
                                       "oracle port of TTSDataset.__getitem__ (3 STFTs, dense-Viterbi pyin, "
                                       "float32 prior), one process per core"},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ CPU baseline first (before CUDA init)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = max(8, min(4 * cores, 128))      # ~10-30 s of CPU work on the box's cores
        audio, wall, cores, per_step = cpu_reference_run(1, 0, n)
        cpu_baseline = {"value": audio / wall, "unit": "audio-s/s", "cores": cores, "kind": "port",
                        "sample": f"first {n} utterances ({audio:.0f} audio-s) of the config-2 manifest, oracle "
                                  "port of TTSDataset.__getitem__ (3 STFTs, dense-Viterbi pyin, float32 "
                                  "prior), one process per core"}

    import torch
    import torch.distributed as dist

    from roar_b200 import synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor, finalize_pitch_stats

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ex = SupDataExtractor(SupConfig(highfreq=8000.0), dev)
    # every rank its own 24 h shard (same manifest shape, different seed)
    synth.CORPORA["C2_rank"] = dict(synth.CORPORA["C2"], seed=synth.CORPORA["C2"]["seed"] + 1000 * rank)
    man, audio, offs, lens = synth.synth_corpus_device("C2_rank", dev, n_utts=args.n_utts)
    offs_h = offs.cpu().numpy()
    lens_h = lens.cpu().numpy().astype(np.int64)
    text_lens = np.array([u.text_len for u in man], dtype=np.int32)
    audio_s = float(lens_h.sum()) / SR
    n_utts = len(man)
    chunk = args.chunk_utts or n_utts
    bounds = [(a, min(a + chunk, n_utts)) for a in range(0, n_utts, chunk)]

    def chunk_batch(a, b):
        lo = int(offs_h[a])
        hi = int(offs_h[b - 1] + (lens_h[b - 1] + 3) // 4 * 4)
        return ex.batch_from_device(audio[lo:hi], offs_h[a:b] - lo, lens_h[a:b])

    batches = [chunk_batch(a, b) for a, b in bounds]
    stats = ex.new_pitch_partials(1)

    def step_device():
        ex.lib.roar_sup_pitch_partials_init(ex._h, ctypes.c_void_p(stats.data_ptr()), 1,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        for (a, b), bt in zip(bounds, batches):
            out = ex.extract(bt, text_lens=text_lens[a:b], stats=stats)
            del out
        if world > 1:
            red = stats.clone()
            dist.all_reduce(red[:, :3], op=dist.ReduceOp.SUM)
            mn, mx = red[:, 3].clone(), red[:, 4].clone()
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            red[:, 3], red[:, 4] = mn, mx
            return red
        return stats

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
    kms = (ctypes.c_double * 11)()
    kcnt = (ctypes.c_int64 * 11)()
    ex.lib.roar_sup_profile_read(ex._h, kms, kcnt, 1)
    ex.lib.roar_sup_set_profiling(ex._h, 0)
    pstats = finalize_pitch_stats(final)

    # ------------------------------------------------------------------ e2e: host buffers, copies inside
    e2e = None
    if not args.no_e2e:
        total_samples = int(audio.numel())
        host_audio = torch.empty(total_samples, dtype=torch.float32, pin_memory=True)
        host_audio.copy_(audio)
        frames_all = int((1 + lens_h // HOP).sum())
        host_out = {k: torch.empty(frames_all * (N_MELS if k == "log_mel" else 1), dtype=torch.float32,
                                   pin_memory=True)
                    for k in ("log_mel", "pitch", "voiced_mask", "p_voiced", "energy")}
        host_stats = torch.empty(1, 5, dtype=torch.float64, pin_memory=True)
        # streamed host chunks: small first and last chunks (the first H2D and the last D2H copy are the only
        # ones nothing hides behind), large ones in between (short occupancy tails in the kernels)
        ec = args.e2e_chunk_utts
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
        from roar_b200.extract_sup_data import merge_partials
        lanes = [ex] + [SupDataExtractor(SupConfig(highfreq=8000.0), dev) for _ in range(n_lane - 1)]
        lane_streams = [torch.cuda.Stream(dev) for _ in range(n_lane)]

        def step_e2e():
            main = torch.cuda.current_stream()
            sts = []
            for e, cs in zip(lanes, lane_streams):
                cs.wait_stream(main)
                with torch.cuda.stream(cs):
                    sts.append(e.new_pitch_partials(1))
            staged = {}

            def stage(i):
                a, b = eb[i]
                lo = int(offs_h[a])
                hi = int(offs_h[b - 1] + (lens_h[b - 1] + 3) // 4 * 4)
                with torch.cuda.stream(copy_stream):
                    d = host_audio[lo:hi].to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                staged[i] = (d, ev, lo)

            stage(0)
            pending = []
            for i, (a, b) in enumerate(eb):
                d, ev, lo = staged.pop(i)
                if i + 1 < len(eb):
                    stage(i + 1)
                e, cs = lanes[i % n_lane], lane_streams[i % n_lane]
                with torch.cuda.stream(cs):
                    cs.wait_event(ev)
                    bt = e.batch_from_device(d, offs_h[a:b] - lo, lens_h[a:b])
                    out = e.extract(bt, text_lens=text_lens[a:b], stats=sts[i % n_lane])
                    done = torch.cuda.Event()
                    done.record(cs)
                f_lo, f_hi = int(frame_cum[a]), int(frame_cum[b])
                with torch.cuda.stream(out_stream):
                    out_stream.wait_event(done)
                    for k in host_out:
                        m = N_MELS if k == "log_mel" else 1
                        host_out[k][m * f_lo:m * f_hi].copy_(out[k], non_blocking=True)
                pending.append((out, d, bt))     # keep device buffers alive until the copies are done
            for cs in lane_streams:
                main.wait_stream(cs)
            st = sts[0]
            for o in sts[1:]:
                st = merge_partials(st, o)
            if world > 1:
                dist.all_reduce(st[:, :3], op=dist.ReduceOp.SUM)
            host_stats.copy_(st, non_blocking=True)
            out_stream.synchronize()
            main.synchronize()
            del pending

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
        e2e = {"value": world * audio_s * args.steps / e2e_s, "unit": "audio-s/s",
               "h2d_bytes_per_step": total_samples * 4,
               "d2h_bytes_per_step": frames_all * 4 * (N_MELS + 4) + 40,
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "pitch_mean": finalize_pitch_stats(host_stats)["pitch_mean"]}

    # ------------------------------------------------------------------ .pt cache writing
    cache_write = None
    if rank == 0 and args.cache_write_utts > 0:
        import shutil
        import tempfile
        from pathlib import Path
        from roar_b200.extract_sup_data import ParallelCacheWriter
        nw = min(args.cache_write_utts, n_utts)
        lo = int(offs_h[0]); hi = int(offs_h[nw - 1] + (lens_h[nw - 1] + 3) // 4 * 4)
        bt = ex.batch_from_device(audio[lo:hi], offs_h[:nw] - lo, lens_h[:nw])
        out = ex.extract(bt, text_lens=text_lens[:nw])
        torch.cuda.synchronize()
        tmp = Path(tempfile.mkdtemp(prefix="roar_sup_cache_"))
        try:
            names = ("log_mel", "pitch", "voiced_mask", "p_voiced", "energy")
            for k in names:
                (tmp / k).mkdir()
            writer = ParallelCacheWriter(min(16, os.cpu_count() or 1))
            t0 = time.perf_counter()
            host = {k: out[k].cpu() for k in names}
            fo = out["frame_off"]
            jobs = [(k, int(fo[i]), int(fo[i + 1]), str(tmp / k / f"utt{i}.pt")) for i in range(nw) for k in names]
            writer.submit_batch(host, jobs, N_MELS)
            writer.drain()
            dt = time.perf_counter() - t0
            writer.close()
            cache_write = {"utterances": nw, "files": 5 * nw, "seconds": dt, "files_per_s": 5 * nw / dt,
                           "audio_s_per_s": float(lens_h[:nw].sum()) / SR / dt,
                           "writer_processes": min(16, os.cpu_count() or 1),
                           "note": "D2H + torch.save of the five cached types through the CLI's writer pool "
                                   "(worker processes over shared memory, temp file + rename); outside the timed extraction region"}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        del out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    hbm_peak, peak_src = measured_peaks()
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
    dom = max(("stft_mel", "pyin_energy", "pyin_cmnd", "pyin_probs", "viterbi", "backtrack", "prior"),
              key=lambda k: per[k])
    # the pYIN chain is one logical kernel split at two scratch hand-offs: its algorithmic traffic is
    # the audio in (4*hop B/frame) and f0 / flag / prob out (12 B/frame)
    pyin_alg = 4 * samples + 12 * frames
    if dom in ("pyin_energy", "pyin_cmnd", "pyin_probs", "viterbi", "backtrack"):
        dom_alg = pyin_alg
        dom_ms = per["pyin_energy"] + per["pyin_cmnd"] + per["pyin_probs"] + per["viterbi"] + per["backtrack"]
        dom_name = "pyin chain (pyin_energy + pyin_cmnd + pyin_probs + viterbi + backtrack); slowest member: " + dom
    else:
        dom_alg, dom_ms, dom_name = alg[dom], per[dom], dom
    n_launch = max(1, len(bounds))
    achieved = dom_alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # measured DRAM traffic of the same kernels (one ncu --set full capture of this workload, committed
    # under profiles/): scratch hand-offs between the chain's kernels make it larger than the
    # algorithmic bytes
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(tpath) and n_utts == 13100 and len(bounds) == 1:
        with open(tpath) as f:
            tk = json.load(f)["kernels"]
        names = {"pyin_energy": "k_pyin_energy", "pyin_cmnd": "k_pyin_cmnd", "pyin_probs": "k_pyin_probs",
                 "viterbi": "k_pyin_viterbi51", "backtrack": "k_pyin_backtrack", "stft_mel": "k_stft_mel",
                 "prior": "k_align_prior"}
        members = ["pyin_energy", "pyin_cmnd", "pyin_probs", "viterbi", "backtrack"] \
            if dom in ("pyin_energy", "pyin_cmnd", "pyin_probs", "viterbi", "backtrack") else [dom]
        vals = [tk.get(names[m], {}) for m in members]
        if all(v.get("dram_read_bytes") is not None and v.get("dram_write_bytes") is not None for v in vals):
            traffic = int(sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in vals))
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_alg // n_launch, "ms_per_launch": dom_ms / n_launch,
                "note": "FP64-pipe / latency bound, not HBM bound (DESIGN.md section 4); "
                        "frac is the honest HBM fraction of its algorithmic bytes"}
    # every kernel of the path against the same HBM peak (algorithmic bytes of that kernel alone)
    roofline_kernels = {}
    for k, b in alg.items():
        if per.get(k, 0) > 0:
            gbs = b / (per[k] * 1e-3) / 1e9
            roofline_kernels[k] = {"ms_per_step": round(per[k], 4), "algorithmic_bytes": int(b),
                                   "achieved_gbs": round(gbs, 2), "frac": round(gbs / hbm_peak, 5)}
    path_bytes = 4 * samples + 4 * (N_MELS + 4) * frames + 4 * prior_elems
    step_ms = total_ms / args.steps
    value = world * audio_s * args.steps / (total_ms * 1e-3)
    line = {
        "metric": "audio-seconds processed/sec (sup-data extraction)", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "utterances_per_gpu": n_utts, "audio_hours_per_gpu": audio_s / 3600,
                   "sup_data_types": ["log_mel", "align_prior_matrix", "pitch", "voiced_mask", "p_voiced", "energy"],
                   "l2": "inputs larger than L2 (%.1f GB audio per pass)" % (samples * 4 / 1e9),
                   "parallelism": f"utterance shards x{world}, one all-reduce of pitch partials"},
        "x_realtime_per_gpu": value / world,
        "gpu_launches": launches, "clocks": clocks,
        "kernels_ms_per_step": {k: round(v, 4) for k, v in per.items() if v > 0},
        "kernels_note": "CUDA-event brackets on each kernel's launching stream inside the timed region; the spectral "
                        "kernels (stft_mel, prior) run on a side stream concurrently with the pYIN chain, so a bracket "
                        "can include time queued behind the other stream (tile_offsets, pyin_energy) and the brackets "
                        "sum to more than ms_per_step",
        "roofline": roofline,
        "roofline_kernels": roofline_kernels,
        "roofline_path": {"bound": "hbm", "achieved": path_bytes / (step_ms * 1e-3) / 1e9, "peak": hbm_peak,
                          "unit": "GB/s", "frac": path_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak,
                          "algorithmic_bytes_per_step": path_bytes},
        "pitch_stats": pstats,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if cache_write is not None:
        line["cache_write"] = cache_write
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
