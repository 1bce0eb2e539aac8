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
``librosa.effects.trim`` on the GPU (``roar_sup_trim``).  Out of scope:
resampling -- the audio must already be at ``dataset.sample_rate``; wav decoding uses
``scipy.io.wavfile`` on the host.
"""
import json
import os
import re
import sys
import threading
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from .config import PITCH_FMAX_C7, PITCH_FMIN_C2, SupConfig

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
    "batch_audio_seconds": 4000.0,   # audio per device call (bounds the Viterbi scratch)
    "writer_processes": None,        # worker processes writing the .pt cache (default: min(16, cores))
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


def read_manifest(path: str, min_duration=None, max_duration=None, ignore_file=None) -> List[dict]:
    ignore = set()
    if ignore_file:
        import pickle
        with open(ignore_file, "rb") as f:
            ignore = set(pickle.load(f))
    out = []
    with open(path, encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            item = json.loads(line)
            dur = item.get("duration")
            if dur is not None:  # filter_by_duration, dataset.py:367-406
                if (min_duration and dur < min_duration) or (max_duration and dur > max_duration):
                    continue
            if Path(item["audio_filepath"]).stem in ignore:
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
    """float32 mono in [-1, 1] like ``AudioSegment.from_file`` (``segment.py:156-278``) for wav files."""
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if sr != sample_rate:
        raise ValueError(f"{path}: sample rate {sr} != dataset.sample_rate {sample_rate}; resampling is outside "
                         "the accelerated path (SURVEY.md section 8f, N3) -- resample the corpus first")
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(2 ** (8 * x.dtype.itemsize - 1))
    elif x.dtype.kind == "u":
        x = (x.astype(np.float32) - 128.0) / 128.0
    else:
        x = x.astype(np.float32)
    if x.ndim > 1:
        x = x.mean(axis=1)
    return np.ascontiguousarray(x, dtype=np.float32)


def shard_indices(durations, world: int, rank: int) -> np.ndarray:
    """Length-bucketed sharding: sort by duration, deal round-robin -> equal audio and equal length mix on
    every rank, no inter-GPU traffic.  Deterministic; ranks partition the set."""
    order = np.argsort(-np.asarray(durations, dtype=np.float64), kind="stable")
    return np.sort(order[rank::world])


def make_batches(lengths, max_samples: int) -> List[np.ndarray]:
    """Group utterance indices (already length-sorted descending within the shard) into device calls."""
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


# ------------------------------------------------------------------------------------ cache writer
class CacheWriter:
    """Writes ``torch.save`` files from a thread pool; temp file + rename keeps the cache idempotent
    under interruption (failure detection)."""

    def __init__(self, threads: int):
        self.pool = ThreadPoolExecutor(max_workers=max(1, threads))
        self.futures = []
        self.lock = threading.Lock()

    @staticmethod
    def _save(tensor: torch.Tensor, path: Path):
        tmp = path.with_suffix(path.suffix + f".tmp{os.getpid()}")
        torch.save(tensor, tmp)
        os.replace(tmp, path)

    def submit(self, tensor: torch.Tensor, path: Path):
        self.futures.append(self.pool.submit(self._save, tensor, path))

    def drain(self):
        for f in self.futures:
            f.result()
        self.futures = []


def _save_range(args):
    """Worker: write the cache files of utterances [lo, hi) of one batch (flat host tensors in shared memory)."""
    host, jobs, n_mels = args
    torch.set_num_threads(1)
    n = 0
    for t, a, e, path in jobs:
        if t == "log_mel":
            ten = host[t][n_mels * a: n_mels * e].view(1, n_mels, e - a).clone()
        else:
            ten = host[t][a:e].clone()
        CacheWriter._save(ten, Path(path))
        n += 1
    return n


class ParallelCacheWriter:
    """``torch.save`` of ~5 small files per utterance is pickling-bound (GIL), so a thread pool tops out at a few
    thousand files/s; this writer hands contiguous utterance ranges of a batch to worker PROCESSES, the flat
    host tensors travelling once per batch through shared memory."""

    def __init__(self, processes: int):
        import torch.multiprocessing as mp      # tensors in shared memory travel by handle, not by value
        self.n = max(1, processes)
        self.pool = mp.get_context("fork").Pool(self.n) if self.n > 1 else None
        self.pending = []

    def submit_batch(self, host: Dict[str, torch.Tensor], jobs: List[tuple], n_mels: int):
        """jobs: (type, frame_begin, frame_end, path) per file, grouped by utterance order."""
        if not jobs:
            return
        if self.pool is None:
            _save_range((host, jobs, n_mels))
            return
        for t in host.values():
            t.share_memory_()
        per = (len(jobs) + self.n - 1) // self.n
        for lo in range(0, len(jobs), per):
            self.pending.append(self.pool.apply_async(_save_range, ((host, jobs[lo:lo + per], n_mels),)))

    def drain(self):
        for r in self.pending:
            r.get()
        self.pending = []

    def close(self):
        self.drain()
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None


# ------------------------------------------------------------------------------------ main
def run(cfg: dict) -> Optional[Dict[str, float]]:
    import torch.distributed as dist

    from .extractor import SupDataExtractor

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # writer processes are forked before this process creates its CUDA context
    writer = ParallelCacheWriter(int(cfg.get("writer_processes") or min(16, (os.cpu_count() or 1) // max(1, world))))
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

    durs = [it.get("duration") or os.path.getsize(it["audio_filepath"]) for it in items]
    mine = shard_indices(durs, world, rank)
    speakers = sorted({str(it.get("speaker", it.get("speaker_id", "default"))) for it in items})
    spk_index = {s: i + 1 for i, s in enumerate(speakers)}      # group 0 = "default" (all)
    stats = empty_partials(len(speakers) + 1)

    todo, todo_ids = [], []
    pitch_types = [t for t in ("pitch", "voiced_mask", "p_voiced") if t in types]
    for i in mine:
        it = items[i]
        uid = rel_audio_id(it["audio_filepath"], base_dir)
        missing = [t for t in folders if not (folders[t] / f"{uid}.pt").exists()]
        if missing or "align_prior_matrix" in types and not folders:
            todo.append(i)
            todo_ids.append((uid, set(missing)))
        if "pitch" in folders and "pitch" not in missing:
            # already cached: its pitch still counts in the corpus statistics (the reference loads it)
            p = torch.load(folders["pitch"] / f"{uid}.pt").float().numpy()
            part = torch.from_numpy(partials_from_pitch(p))
            g = spk_index[str(it.get("speaker", it.get("speaker_id", "default")))]
            stats[0] = merge_partials(stats[0], part)
            stats[g] = merge_partials(stats[g], part)

    wavs = {}

    def get_wav(i):
        if i not in wavs:
            wavs[i] = load_wav(items[i]["audio_filepath"], scfg.sample_rate)
        return wavs[i]

    lengths = np.array([len(get_wav(i)) for i in todo], dtype=np.int64) if todo else np.zeros(0, np.int64)
    for b in make_batches(lengths, int(float(cfg.get("batch_audio_seconds", 4000.0)) * scfg.sample_rate)):
        idx = [todo[k] for k in b]
        batch = ex.pack([get_wav(i) for i in idx])
        if trim_kw is not None:
            batch = ex.trim(batch, **trim_kw)
        need = set().union(*[todo_ids[k][1] for k in b])
        want = [t for t in types if t in need or t == "align_prior_matrix"]
        if pitch_types and need & set(pitch_types):
            want = list(set(want) | {"pitch", "voiced_mask", "p_voiced"})
        tl = [text_length(items[i]) for i in idx]
        groups = np.array([spk_index[str(items[i].get("speaker", items[i].get("speaker_id", "default")))] for i in idx],
                          dtype=np.int32)
        out = ex.extract(batch, text_lens=tl, types=want)
        if "pitch" in out and out["pitch"] is not None:
            newly = np.array(["pitch" in todo_ids[k][1] or "pitch" not in folders for k in b])
            gp = ex.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], np.where(newly, groups, -1),
                                           len(speakers) + 1).cpu()
            stats = merge_partials(stats, gp)
            gp0 = ex.pitch_partials_grouped(out["pitch"], out["pitch_frame_off"], np.where(newly, 0, -1), 1).cpu()
            stats[0] = merge_partials(stats[0], gp0[0])
        host = {k: out[k].cpu() for k in SUP_TYPES_ON_DISK if k in out and out[k] is not None and k in folders}
        fo = out.get("frame_off")
        pfo = out.get("pitch_frame_off")
        jobs = []
        for n, k in enumerate(b):
            uid, missing = todo_ids[k]
            for t in missing:
                off = fo if t in ("log_mel", "energy") else pfo
                jobs.append((t, int(off[n]), int(off[n + 1]), str(folders[t] / f"{uid}.pt")))
        writer.submit_batch(host, jobs, scfg.n_mels)
        for i in idx:
            wavs.pop(i, None)
    writer.close()

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
    return result


def main(argv=None):
    cfg = load_config(list(sys.argv[1:] if argv is None else argv))
    print(cfg["dataset"])
    return run(cfg)


if __name__ == "__main__":
    main()
