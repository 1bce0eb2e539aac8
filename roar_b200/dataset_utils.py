"""Consumer-side helpers for the cached sup-data: what ``TTSDataset.__getitem__`` does AFTER loading the
``.pt`` files (host-side, trivial arithmetic -- kept here so a training loop that does not use the
reference's dataset class can still honour its contract).

* ``normalize_pitch``      -- ``roar/collections/tts/data/dataset.py:716-741``
* ``select_pitch_stats``   -- the ``pitch_mean``/``pitch_std`` vs ``pitch_stats_path`` (per speaker / "default")
                              precedence of the same lines; the JSON is what ``roar_b200.extract_sup_data`` and
                              ``compute_speaker_stats.py:105-132`` write.
"""
import json
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch


def select_pitch_stats(sample: Dict[str, Any], pitch_mean: Optional[float] = None, pitch_std: Optional[float] = None,
                       pitch_stats: Optional[Dict[str, Dict[str, float]]] = None) -> Tuple[float, float]:
    if pitch_mean is not None and pitch_std is not None:
        return float(pitch_mean), float(pitch_std)
    if pitch_stats:
        if "speaker_id" in sample and str(sample["speaker_id"]) in pitch_stats:
            st = pitch_stats[str(sample["speaker_id"])]
        elif "default" in pitch_stats:
            st = pitch_stats["default"]
        else:
            raise ValueError(f"Could not find pitch stats for {sample}.")
        return float(st["pitch_mean"]), float(st["pitch_std"])
    raise ValueError("Missing statistics for pitch normalization.")


def load_pitch_stats(path) -> Dict[str, Dict[str, float]]:
    with open(path, encoding="utf-8") as f:
        return json.load(f)


def normalize_pitch(pitch: torch.Tensor, mean: float, std: float) -> torch.Tensor:
    """In place, like the reference: ``pitch -= mean; pitch[pitch == -mean] = 0; pitch /= std`` -- frames
    that were 0 (unvoiced) stay 0."""
    pitch -= mean
    pitch[pitch == -mean] = 0.0
    pitch /= std
    return pitch


# ------------------------------------------------------------------------------------ packed cache (row N1)
def write_packed_batch(packed_dir, tag, base, elem_off, shapes, keys, index_lines):
    """One raw little-endian float32 shard for a whole batch (``<packed_dir>/shard_<tag>.bin``) and one index line
    per tensor: ``{"key": "<type>/<id>", "shard": ..., "offset": <elements>, "shape": [...]}``.  ``base`` is the flat
    CPU float32 tensor the batch's outputs were copied into; only the referenced ranges are written, back to back."""
    import json
    import os
    name = f"shard_{tag}.bin"
    tmp = os.path.join(str(packed_dir), name + f".tmp{os.getpid()}")
    pos = 0
    with open(tmp, "wb") as f:
        arr = base.numpy()
        for o, sh, key in zip(elem_off, shapes, keys):
            n = int(np.prod(sh))
            f.write(arr[int(o):int(o) + n].tobytes())
            index_lines.append(json.dumps({"key": key, "shard": name, "offset": pos, "shape": list(sh)}) + "\n")
            pos += n
    os.replace(tmp, os.path.join(str(packed_dir), name))


class PackedCache:
    """Reader of the packed cache: ``load("pitch", utt_id)`` returns the CPU float32 tensor ``torch.load`` of the
    per-file layout would (``log_mel`` ``[1, n_mels, T]``, the rest ``[T]``), memory-mapped shard by shard."""

    def __init__(self, packed_dir):
        import json
        from pathlib import Path
        self.dir = Path(packed_dir)
        self.index = {}
        for p in sorted(self.dir.glob("index_r*.jsonl")):
            with open(p, encoding="utf-8") as f:
                for line in f:
                    if line.strip():
                        e = json.loads(line)
                        self.index[e["key"]] = e
        self._maps = {}

    @staticmethod
    def index_ids(packed_dir):
        import json
        from pathlib import Path
        ids = set()
        for p in Path(packed_dir).glob("index_r*.jsonl"):
            with open(p, encoding="utf-8") as f:
                for line in f:
                    if line.strip():
                        ids.add(json.loads(line)["key"].split("/", 1)[1])
        return ids

    def __contains__(self, key):
        return key in self.index

    def load(self, sup_type: str, utt_id: str):
        import torch
        e = self.index[f"{sup_type}/{utt_id}"]
        m = self._maps.get(e["shard"])
        if m is None:
            m = self._maps[e["shard"]] = np.memmap(self.dir / e["shard"], dtype="<f4", mode="r")
        n = int(np.prod(e["shape"]))
        return torch.from_numpy(np.array(m[e["offset"]:e["offset"] + n])).view(*e["shape"])
