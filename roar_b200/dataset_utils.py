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
