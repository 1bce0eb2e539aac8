"""Drop-in for ``scripts/dataset_processing/tts/extract_sup_data.py`` on B200.

    python -m roar_b200.extract_sup_data manifest_filepath=train.json sup_data_path=sup_data \\
        "sup_data_types=[align_prior_matrix,pitch,energy]" dataset.sample_rate=22050 dataset.n_fft=1024 ...
    torchrun --nproc-per-node 8 -m roar_b200.extract_sup_data ...          # one rank per GPU

Same contract as the reference (file:line relative to its root):
  * Hydra-style ``key=value`` overrides with the keys of ``ds_conf/ds_for_fastpitch_align.yaml:1-33``
    (an optional ``--config-path/--config-name`` YAML supplies defaults; hydra/omegaconf are not needed);
  * cache layout ``<sup_data_path>/<type>/<rel_audio_path_as_text_id>.pt`` holding CPU float32 tensors --
    log_mel ``[1, n_mels, T]``, pitch / voiced_mask / p_voiced / energy ``[T]``
    (``tts/data/dataset.py:581-586, 651-657, 685-708, 746-753``); an existing file is not recomputed;
  * stdout ``PITCH_MEAN=..., PITCH_STD=...`` / ``PITCH_MIN=..., PITCH_MAX=...``
    (``extract_sup_data.py:8-13``) over non-zero pitch frames, unbiased std.
Differences, on purpose: sup types are addressed by NAME (the reference unpacks the batch
positionally); statistics are accumulated in float64 and all-reduced across
ranks; ``pitch_stats.json`` (``default`` + per speaker, the format ``pitch_stats_path`` reads,
``dataset.py:485-487, 720-733``) is written next to the cache.
``dataset.trim=true`` (+ ``trim_ref``, ``trim_top_db``, ``trim_frame_length``, ``trim_hop_length``) runs
``librosa.effects.trim`` on the GPU (``roar_sup_trim``); wav files at another sample rate are converted to
``dataset.sample_rate`` on the GPU (``roar_sup_resample``, polyphase FIR -- the reference uses librosa's soxr_hq
there, so resampled audio agrees only as two good low-pass designs do: roar_b200/resample.py).

The run is a bounded-memory stream (the reference overlaps decode and compute through
``DataLoader(num_workers=...)``, ``extract_sup_data.py:66-71``):

    native wav decode threads -> ring of pinned staging buffers -> H2D (16-bit PCM stays int16, converted on the
    GPU) -> kernels -> async D2H into pinned output buffers -> native ``.pt`` writer threads

with ``pipeline_depth`` batches in flight, so host memory is a few staging buffers whatever the manifest size and
the GPU never waits for a whole shard to decode.  ``align_prior_matrix`` is accepted in ``sup_data_types`` (the
reference's ds_conf lists it) but nothing is computed for it here: the reference never caches it and discards the
batch element (``extract_sup_data.py:24,29-30``).  A manifest entry's ``mel_filepath`` that exists counts as the
cached log-mel (``dataset.py:646-649``).
"""
import json
import os
import re
import sys
import heapq
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from .config import PITCH_FMAX_C7, PITCH_FMIN_C2, SupConfig
from .dataset_utils import PackedCache, write_packed_batch

SUP_TYPES_ON_DISK = ("log_mel", "pitch", "voiced_mask", "p_voiced", "energy")
VALID_SUP_TYPES = SUP_TYPES_ON_DISK + ("align_prior_matrix", "speaker_id")

DEFAULTS = {
    "name": "ds_for_fastpitch_align",
    "manifest_filepath": "train_manifest.json",
    "sup_data_path": "sup_data",
    "sup_data_types": ["align_prior_matrix", "pitch", "energy"],
    "dataloader_params": {"num_workers": 16},
    "dataset": {
        "sample_rate": 22050, "n_fft": 1024, "win_length": 1024, "hop_length": 256, "window": "hann",
        "n_mels": 80, "lowfreq": 0, "highfreq": 8000, "max_duration": None, "min_duration": 0.1,
        "ignore_file": None, "trim": False, "pitch_fmin": PITCH_FMIN_C2, "pitch_fmax": PITCH_FMAX_C7,
    },
    "batch_audio_seconds": 12000.0,  # audio per device call (bounds the Viterbi scratch and the staging buffers)
    "pipeline_depth": 3,             # staging-buffer sets in flight: decode | compute | write
    "writer_threads": None,          # native .pt writer threads (default: min(16, cores))
    # "pt" (default): one torch-loadable file per utterance and type, the layout TTSDataset reads;
    # "packed": one raw float32 shard per batch and type + a JSON-lines index -- two
    # sequential writes per batch instead of five small files per utterance; read back with
    # roar_b200.dataset_utils.PackedCache (load(type, id) -> the tensor torch.load would have returned)
    "cache_format": "pt",
}


# ------------------------------------------------------------------------------------ config
def _parse_value(s: str):
    s = s.strip()
    if s.startswith("[") and s.endswith("]"):
        inner = s[1:-1].strip()
        return [_parse_value(x) for x in inner.split(",")] if inner else []
    if (s.startswith('"') and s.endswith('"')) or (s.startswith("'") and s.endswith("'")):
        return s[1:-1]
    low = s.lower()
    if low in ("null", "none", "~"):
        return None
    if low in ("true", "false"):
        return low == "true"
    for cast in (int, float):
        try:
            return cast(s)
        except ValueError:
            pass
    return s


def _set_path(cfg: dict, dotted: str, value):
    keys = dotted.lstrip("+").split(".")
    d = cfg
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def _interpolate(cfg: dict):
    """``${key}`` / ``${a.b}`` references as used by the reference YAMLs."""
    def lookup(path):
        d = cfg
        for k in path.split("."):
            d = d[k]
        return d

    def walk(node):
        if isinstance(node, dict):
            return {k: walk(v) for k, v in node.items()}
        if isinstance(node, list):
            return [walk(v) for v in node]
        if isinstance(node, str):
            m = re.fullmatch(r"\$\{([^}]+)\}", node)
            if m:
                return walk(lookup(m.group(1)))
        return node
    return walk(cfg)


def load_config(argv: List[str]) -> dict:
    import copy
    cfg = copy.deepcopy(DEFAULTS)
    cfg_path = cfg_name = None
    rest = []
    for a in argv:
        if a.startswith("--config-path"):
            cfg_path = a.split("=", 1)[1]
        elif a.startswith("--config-name"):
            cfg_name = a.split("=", 1)[1]
        else:
            rest.append(a)
    if cfg_path or cfg_name:
        import yaml
        name = cfg_name or "ds_for_fastpitch_align"
        p = Path(cfg_path or ".") / (name if name.endswith(".yaml") else name + ".yaml")
        with open(p) as f:
            y = yaml.safe_load(f)

        def merge(a, b):
            for k, v in b.items():
                if isinstance(v, dict) and isinstance(a.get(k), dict):
                    merge(a[k], v)
                else:
                    a[k] = v
        merge(cfg, y)
    for a in rest:
        if "=" not in a:
            raise SystemExit(f"expected key=value override, got {a!r}")
        k, v = a.split("=", 1)
        _set_path(cfg, k, _parse_value(v))
    return _interpolate(cfg)


def sup_config_from(cfg: dict) -> SupConfig:
    d = cfg["dataset"]
    return SupConfig(sample_rate=int(d["sample_rate"]), n_fft=int(d["n_fft"]),
                     win_length=d.get("win_length"), hop_length=d.get("hop_length"), window=d.get("window", "hann"),
                     n_mels=int(d["n_mels"]), lowfreq=float(d.get("lowfreq") or 0.0),
                     highfreq=(float(d["highfreq"]) if d.get("highfreq") else None),
                     pitch_fmin=float(d.get("pitch_fmin", PITCH_FMIN_C2)),
                     pitch_fmax=float(d.get("pitch_fmax", PITCH_FMAX_C7)))


# ------------------------------------------------------------------------------------ manifest
def get_base_dir(paths) -> Path:
    """Common parent directory of all audio files (``tts_dataset_utils.py:152-175``)."""
    base = None
    for p in paths:
        d = Path(p).parent
        if base is None:
            base = d
            continue
        while True:
            try:
                d.relative_to(base)
                break
            except ValueError:
                if base == base.parent:
                    break
                base = base.parent
    return base


def rel_audio_id(path: str, base_dir: Path) -> str:
    """``dataset.py:581-586``: relative path, suffix stripped, '/' -> '_'."""
    return str(Path(path).relative_to(base_dir).with_suffix("")).replace("/", "_")


def read_manifest(path, min_duration=None, max_duration=None, ignore_file=None) -> List[dict]:
    """Manifest loading and pruning with the reference's rules (``dataset.py:215-275, 367-406``):
    ``path`` is one manifest or a list of them; the duration filter applies only when EVERY entry carries a
    ``duration`` (``total_duration is not None``); ``ignore_file`` is a pickled list of ``audio_filepath``
    strings compared verbatim (each listed path prunes one entry)."""
    paths = [path] if isinstance(path, (str, os.PathLike)) else list(path)
    data = []
    all_have_duration = True
    for mf in paths:
        with open(Path(mf).expanduser(), encoding="utf-8") as f:
            for line in f:
                line = line.strip()
                if not line:
                    continue
                item = json.loads(line)
                if item.get("duration") is None:
                    all_have_duration = False
                data.append(item)
    ignore = set()
    if ignore_file:
        import pickle
        with open(Path(ignore_file).expanduser(), "rb") as f:
            ignore = set(pickle.load(f))
    out = []
    for item in data:
        if all_have_duration:
            dur = item["duration"]
            if (min_duration and dur < min_duration) or (max_duration and dur > max_duration):
                continue
        if ignore_file and item["audio_filepath"] in ignore:
            ignore.remove(item["audio_filepath"])
            continue
        out.append(item)
    return out


def text_length(item: dict) -> int:
    """Length the aligner prior needs.  The reference tokenises ``normalized_text`` with its own
    tokenizer (out of scope); a manifest may carry ``text_tokens`` (``dataset.py:635-637``), else the
    character count + 2 (``pad_with_space``) is used.  Only the in-memory prior depends on it."""
    if "text_tokens" in item:
        return len(item["text_tokens"])
    txt = item.get("normalized_text", item.get("text", ""))
    return len(txt) + 2


def load_wav(path: str, sample_rate: int) -> np.ndarray:
    """float32 mono in [-1, 1] like ``AudioSegment.from_file`` (``segment.py:156-278``) for wav files
    (native decoder, channels averaged)."""
    from . import host_io
    x, sr = host_io.load_wav(path)
    if sr != sample_rate:
        raise ValueError(f"{path}: sample rate {sr} != requested {sample_rate}")
    return x


def shard_indices(durations, world: int, rank: int) -> np.ndarray:
    """Utterance -> rank assignment with no inter-GPU traffic: longest-processing-time
    greedy on the durations -- utterances by decreasing duration, each to the rank with the least audio so far
    (ties: lowest rank).  Every rank ends within one utterance of the same total audio and gets the same length
    mix.  Deterministic; the ranks partition the set.  -> sorted indices of ``rank``."""
    d = np.asarray(durations, dtype=np.float64)
    if world <= 1:
        return np.arange(len(d))
    order = np.argsort(-d, kind="stable")
    owner = np.empty(len(d), dtype=np.int32)
    heap = [(0.0, r) for r in range(world)]
    for i in order:
        load, r = heap[0]
        owner[i] = r
        heapq.heapreplace(heap, (load + float(d[i]), r))
    return np.flatnonzero(owner == rank)


def make_batches(lengths, max_samples: int) -> List[np.ndarray]:
    """Group utterance indices (length-sorted descending) into device calls of at most ``max_samples``."""
    idx = np.argsort(-np.asarray(lengths), kind="stable")
    batches, cur, tot = [], [], 0
    for i in idx:
        if cur and tot + lengths[i] > max_samples:
            batches.append(np.array(cur))
            cur, tot = [], 0
        cur.append(i)
        tot += int(lengths[i])
    if cur:
        batches.append(np.array(cur))
    return batches


# ------------------------------------------------------------------------------------ statistics
def empty_partials(n_groups: int = 1) -> torch.Tensor:
    t = torch.zeros(n_groups, 5, dtype=torch.float64)
    t[:, 3] = float("inf")
    return t


def partials_from_pitch(p: np.ndarray) -> np.ndarray:
    v = p[p != 0].astype(np.float64)
    if v.size == 0:
        return np.array([0, 0, 0, np.inf, 0], dtype=np.float64)
    return np.array([v.sum(), (v * v).sum(), v.size, v.min(), v.max()], dtype=np.float64)


def merge_partials(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    out = a.clone()
    out[..., :3] += b[..., :3]
    out[..., 3] = torch.minimum(a[..., 3], b[..., 3])
    out[..., 4] = torch.maximum(a[..., 4], b[..., 4])
    return out


def allreduce_partials(t: torch.Tensor) -> torch.Tensor:
    """The one exchange step of the path: SUM on (sum, sumsq, count), MIN / MAX on the extremes."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    t = t.clone()
    sums = t[..., :3].contiguous()
    mn = t[..., 3].contiguous()
    mx = t[..., 4].contiguous()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t[..., :3], t[..., 3], t[..., 4] = sums, mn, mx
    return t


def stats_from_partials(row) -> Optional[Dict[str, float]]:
    s, q, n, mn, mx = [float(x) for x in row]
    if n < 1:
        return None
    mean = s / n
    std = float(np.sqrt(max((q - s * s / n) / (n - 1), 0.0))) if n > 1 else float("nan")
    return {"pitch_mean": mean, "pitch_std": std, "pitch_min": mn, "pitch_max": mx}


# ------------------------------------------------------------------------------------ staging
class StagingSlot:
    """One set of pinned host buffers: packed input audio (int16 PCM or float32) and one flat float32 output
    buffer per cached type.  Buffers grow to the largest batch seen and are reused."""

    def __init__(self):
        self.bufs: Dict[str, torch.Tensor] = {}
        self.free = threading.Event()
        self.free.set()

    def buf(self, name: str, n: int, dtype) -> torch.Tensor:
        t = self.bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            self.bufs[name] = None
            t = torch.empty(max(n, 1), dtype=dtype, pin_memory=torch.cuda.is_available())
            self.bufs[name] = t
        return t


def _zero_gaps(hv: np.ndarray, offs: np.ndarray, lens: np.ndarray, total: int):
    """The alignment gaps between packed utterances (<= 3 samples each) are staged by the kernels' bulk copies:
    keep them finite."""
    ends = offs + lens
    nxt = np.concatenate([offs[1:], [total]])
    for k in range(int((nxt - ends).max()) if len(offs) else 0):
        m = nxt - ends > k
        hv[ends[m] + k] = 0


class LoadedBatch:
    __slots__ = ("slot", "idx", "host", "offs", "lens", "total", "pcm16", "srs")


def load_batch(slot: StagingSlot, paths: List[str], sample_rate: int, threads: int) -> LoadedBatch:
    """Decode one batch of wav files into ``slot`` (native threads).  16-bit mono PCM stays int16 (2 bytes per
    sample over PCIe, ``x / 2**15`` on the GPU); any other encoding is decoded to float32 mono on the host."""
    from . import host_io
    from .extractor import pack_layout
    info = host_io.wav_probe(paths, threads)
    lens = info["n_frames"].astype(np.int64)
    offs, total = pack_layout(lens)
    pcm16 = bool(host_io.is_pcm16_mono(info).all())
    host = slot.buf("pcm16" if pcm16 else "f32", total, torch.int16 if pcm16 else torch.float32)
    host_io.wav_read(paths, info, np.zeros(len(paths), np.int64), lens, host, offs, -1, threads)
    _zero_gaps(host.numpy(), offs, lens, total)
    lb = LoadedBatch()
    lb.slot, lb.host, lb.offs, lb.lens, lb.total, lb.pcm16 = slot, host, offs, lens, total, pcm16
    # files at another rate are resampled on the GPU (target_sr, segment.py:68-75)
    lb.srs = None if bool((info["sample_rate"] == sample_rate).all()) else info["sample_rate"].astype(np.int64)
    return lb


# ------------------------------------------------------------------------------------ main
def _speaker(item: dict) -> str:
    return str(item.get("speaker", item.get("speaker_id", "default")))


def run(cfg: dict) -> Optional[Dict[str, float]]:
    import torch.distributed as dist

    from . import host_io
    from .extractor import SupDataExtractor
    from .resample import resample_batch

    t_start = time.perf_counter()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    types = list(cfg["sup_data_types"])
    for t in types:
        if t not in VALID_SUP_TYPES:
            raise NotImplementedError(f"sup_data_type {t!r} is outside the accelerated path "
                                      f"(supported: {VALID_SUP_TYPES})")
    d = cfg["dataset"]
    trim_kw = None
    if d.get("trim"):
        # librosa.effects.trim with the TTSDataset defaults (dataset.py:285-291); trim_ref=None means np.max
        trim_kw = dict(top_db=d.get("trim_top_db") if d.get("trim_top_db") is not None else 60,
                       ref=d.get("trim_ref"), frame_length=d.get("trim_frame_length") or 2048,
                       hop_length=d.get("trim_hop_length") or 512)
    scfg = sup_config_from(cfg)
    ex = SupDataExtractor(scfg)
    items = read_manifest(cfg["manifest_filepath"], d.get("min_duration"), d.get("max_duration"), d.get("ignore_file"))
    base_dir = get_base_dir([it["audio_filepath"] for it in items])
    sup_path = Path(cfg["sup_data_path"])
    folders = {t: Path(d.get(f"{t}_folder") or sup_path / t) for t in SUP_TYPES_ON_DISK if t in types}
    if rank == 0:
        for f in folders.values():
            f.mkdir(parents=True, exist_ok=True)
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(f"Processing {cfg['manifest_filepath']}:")

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_rank = max(1, cores // max(1, world))
    decode_threads = max(1, min(int(cfg.get("dataloader_params", {}).get("num_workers", 16) or 1), per_rank))
    writer_threads = int(cfg.get("writer_threads") or min(16, per_rank))

    durs = [it.get("duration") or os.path.getsize(it["audio_filepath"]) for it in items]
    mine = shard_indices(durs, world, rank)
    speakers = sorted({_speaker(it) for it in items})
    spk_index = {s: i + 1 for i, s in enumerate(speakers)}      # group 0 = "default" (all)
    n_groups = len(speakers) + 1
    stats = empty_partials(n_groups)                            # utterances whose pitch was already cached
    acc = ex.new_pitch_partials(n_groups)                       # newly computed, accumulated on the device

    # which cache files exist: one directory listing per type instead of five stat() calls per utterance
    existing = {t: set(os.listdir(f)) if f.is_dir() else set() for t, f in folders.items()}
    todo, todo_ids = [], []
    for i in mine:
        it = items[i]
        uid = rel_audio_id(it["audio_filepath"], base_dir)
        missing = [t for t in folders if f"{uid}.pt" not in existing[t]]
        mel_path = it.get("mel_filepath")
        if "log_mel" in missing and mel_path is not None and Path(mel_path).exists():
            missing.remove("log_mel")                           # dataset.py:646-649: the manifest's own mel file wins
        if missing:
            todo.append(i)
            todo_ids.append((uid, set(missing)))
        if "pitch" in folders and "pitch" not in missing:
            # already cached: its pitch still counts in the corpus statistics (the reference loads it)
            p = torch.load(folders["pitch"] / f"{uid}.pt").float().numpy()
            part = torch.from_numpy(partials_from_pitch(p))
            g = spk_index[_speaker(it)]
            stats[0] = merge_partials(stats[0], part)
            stats[g] = merge_partials(stats[g], part)

    packed = str(cfg.get("cache_format", "pt")).lower() == "packed"
    packed_dir = (sup_path / "packed") if packed else None
    packed_index: List[str] = []
    run_id = f"{int(time.time()):x}"                           # shards of a resumed run never overwrite earlier ones
    if packed:
        packed_dir.mkdir(parents=True, exist_ok=True)
        done = PackedCache.index_ids(packed_dir)              # utterances a previous packed run already holds
        keep = [k for k, (uid, _) in enumerate(todo_ids) if uid not in done]
        todo, todo_ids = [todo[k] for k in keep], [todo_ids[k] for k in keep]

    sr = scfg.sample_rate
    # batch composition from the manifest (duration, else file size): exact lengths come from the wav headers
    approx = np.array([int(items[i]["duration"] * sr) if items[i].get("duration") else
                       os.path.getsize(items[i]["audio_filepath"]) // 2 for i in todo], dtype=np.int64)
    batches = make_batches(approx, int(float(cfg.get("batch_audio_seconds", 12000.0)) * sr)) if len(todo) else []
    depth = max(2, int(cfg.get("pipeline_depth", 3)))
    slots = [StagingSlot() for _ in range(depth)]
    if batches:
        # pin the staging buffers up front (page-locking a few hundred MB takes longer than processing a batch):
        # 16-bit input and the float32 outputs of the largest batch, plus headroom for rounded durations
        big = max(int(approx[b].sum()) for b in batches)
        n_big = max(len(b) for b in batches)
        per_frame = sum((scfg.n_mels if t == "log_mel" else 1) for t in folders)
        for sl in slots[:min(depth, len(batches))]:
            sl.buf("pcm16", int(big * 1.02) + 4 * n_big + 4096, torch.int16)
            sl.buf("out", (int(big * 1.02) // scfg.hop + 2 * n_big) * per_frame, torch.float32)
    timers = {"load": 0.0, "write": 0.0, "wait_load": 0.0, "wait_slot": 0.0}
    pitch_types = {"pitch", "voiced_mask", "p_voiced"}
    spec_pad = scfg.n_fft // 2

    def prefetch(bi: int) -> LoadedBatch:
        slot = slots[bi % depth]
        t0 = time.perf_counter()
        slot.free.wait()
        slot.free.clear()
        t1 = time.perf_counter()
        idx = [todo[k] for k in batches[bi]]
        lb = load_batch(slot, [items[i]["audio_filepath"] for i in idx], sr, decode_threads)
        lb.idx = idx
        timers["wait_slot"] += t1 - t0
        timers["load"] += time.perf_counter() - t1
        return lb

    write_q: "queue.Queue" = queue.Queue()
    write_err: List[BaseException] = []

    def writer_loop():
        while True:
            job = write_q.get()
            if job is None:
                return
            slot, ev, base, offs, shapes, paths, keep = job
            try:
                ev.synchronize()
                del keep
                if not write_err:
                    t0 = time.perf_counter()
                    if packed_dir is None:
                        host_io.pt_write_batch(base, offs, shapes, paths, writer_threads)
                    else:
                        write_packed_batch(packed_dir, f"r{rank}_{run_id}_b{len(packed_index):06d}", base, offs, shapes, paths, packed_index)
                    timers["write"] += time.perf_counter() - t0
            except BaseException as e:      # surfaced by the main thread
                write_err.append(e)
            finally:
                slot.free.set()

    wt = threading.Thread(target=writer_loop, daemon=True)
    wt.start()
    # three streams: H2D of batch b+1 and D2H of batch b-1 overlap the kernels of batch b
    main_stream = torch.cuda.current_stream()
    in_stream, out_stream = torch.cuda.Stream(), torch.cuda.Stream()
    loader = ThreadPoolExecutor(max_workers=1)
    audio_seconds = 0.0
    n_done = 0
    t_stream = time.perf_counter()
    try:
        fut = loader.submit(prefetch, 0) if batches else None
        for bi, b in enumerate(batches):
            t0 = time.perf_counter()
            lb = fut.result()
            timers["wait_load"] += time.perf_counter() - t0
            fut = loader.submit(prefetch, bi + 1) if bi + 1 < len(batches) else None
            if write_err:
                raise write_err[0]
            need = set().union(*[todo_ids[k][1] for k in b])
            want = [t for t in SUP_TYPES_ON_DISK if t in need]
            if need & pitch_types:
                want = sorted(set(want) | pitch_types)
            with torch.cuda.stream(in_stream):
                raw = lb.host[:lb.total].to(ex.device, non_blocking=True)
            main_stream.wait_stream(in_stream)
            raw.record_stream(main_stream)
            batch = ex.batch_from_device(ex.pcm16_to_f32(raw) if lb.pcm16 else raw, lb.offs, lb.lens)
            if lb.srs is not None:
                batch = resample_batch(ex, batch, lb.srs, sr)
            if trim_kw is not None:
                batch = ex.trim(batch, **trim_kw)
            if {"log_mel", "energy"} & set(want) and int(batch.lens_host.min()) <= spec_pad:
                k = int(np.argmin(batch.lens_host))
                raise ValueError(f"{items[lb.idx[k]]['audio_filepath']}: {int(batch.lens_host[k])} samples, not longer "
                                 f"than the STFT reflect padding ({spec_pad}); torch.stft rejects it in the reference too")
            out = ex.extract(batch, types=want)
            if out.get("pitch") is not None:
                groups = np.array([spk_index[_speaker(items[i])] for i in lb.idx], dtype=np.int32)
                newly = np.array(["pitch" in todo_ids[k][1] or "pitch" not in folders for k in b])
                ex.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], np.where(newly, groups, -1), n_groups, acc)
                ex.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], np.where(newly, 0, -1), n_groups, acc)
            # D2H of every tensor this batch writes into ONE pinned buffer, then the native writer
            fo, pfo = out.get("frame_off"), out.get("pitch_frame_off")
            seg, pos = {}, 0
            for t in want:
                if t in folders and out.get(t) is not None:
                    seg[t] = pos
                    pos += out[t].numel()
            hostbuf = lb.slot.buf("out", pos, torch.float32)
            out_stream.wait_stream(main_stream)
            with torch.cuda.stream(out_stream):
                for t, o in seg.items():
                    hostbuf[o:o + out[t].numel()].copy_(out[t], non_blocking=True)
                    out[t].record_stream(out_stream)
                ev = torch.cuda.Event()
                ev.record(out_stream)
            offs, shapes, paths = [], [], []
            for n, k in enumerate(b):
                uid, missing = todo_ids[k]
                for t in missing:
                    if t not in seg:
                        continue
                    off = fo if t in ("log_mel", "energy") else pfo
                    a, e = int(off[n]), int(off[n + 1])
                    if t == "log_mel":
                        offs.append(seg[t] + scfg.n_mels * a); shapes.append((1, scfg.n_mels, e - a))
                    else:
                        offs.append(seg[t] + a); shapes.append((e - a,))
                    paths.append(f"{t}/{uid}" if packed else str(folders[t] / f"{uid}.pt"))
            write_q.put((lb.slot, ev, hostbuf, np.array(offs, dtype=np.int64), shapes, paths, (out, batch)))
            audio_seconds += float(batch.lens_host.sum()) / sr
            n_done += len(b)
    finally:
        write_q.put(None)
        wt.join()
        loader.shutdown(wait=True)
    if write_err:
        raise write_err[0]
    if packed and packed_index:
        with open(packed_dir / f"index_r{rank}.jsonl", "a", encoding="utf-8") as f:
            f.write("".join(packed_index))
    stream_seconds = time.perf_counter() - t_stream      # first decode submitted -> last cache file renamed

    stats = merge_partials(stats, acc.cpu())
    stats = allreduce_partials(stats.cuda()).cpu() if world > 1 else stats
    result = stats_from_partials(stats[0])
    if rank == 0:
        if result is not None:
            f32 = np.float32
            print(f"PITCH_MEAN={float(f32(result['pitch_mean']))}, PITCH_STD={float(f32(result['pitch_std']))}")
            print(f"PITCH_MIN={float(f32(result['pitch_min']))}, PITCH_MAX={float(f32(result['pitch_max']))}")
        table = {"default": result}
        for s, g in spk_index.items():
            r = stats_from_partials(stats[g])
            if r is not None:
                table[s] = r
        sup_path.mkdir(parents=True, exist_ok=True)
        with open(sup_path / "pitch_stats.json", "w", encoding="utf-8") as f:
            json.dump(table, f, indent=2)
    if world > 1:
        dist.barrier()
    if result is not None:
        result = dict(result)
        result["run"] = {"utterances": n_done, "audio_seconds": audio_seconds, "seconds": time.perf_counter() - t_start,
                         "stream_seconds": stream_seconds, "stage_seconds": {k: round(v, 4) for k, v in timers.items()},
                         "decode_threads": decode_threads, "writer_threads": writer_threads, "batches": len(batches)}
    return result


def main(argv=None):
    cfg = load_config(list(sys.argv[1:] if argv is None else argv))
    print(cfg["dataset"])
    return run(cfg)


if __name__ == "__main__":
    main()
