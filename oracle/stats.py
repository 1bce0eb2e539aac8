"""Oracle: corpus pitch statistics and pitch normalisation.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``pitch_stats`` follows ``scripts/dataset_processing/tts/extract_sup_data.py:8-13,29-30``
  (float32 ``torch.cat`` -> ``mean``, unbiased ``std``, ``min``, ``max`` over ``pitch != 0``)
  and ``compute_speaker_stats.py:56-60``.
* ``pitch_stats_f64`` is the same statistic accumulated in float64 (what the CUDA path
  all-reduces); the 1e-5 gate compares against this and reports the float32 value beside it.
* ``normalize_pitch`` follows ``roar/collections/tts/data/dataset.py:716-741``.
"""
import numpy as np
import torch


def pitch_stats(pitch_list):
    t = torch.cat([torch.as_tensor(np.asarray(p, dtype=np.float32)) for p in pitch_list])
    t = t[t != 0]
    return dict(mean=t.mean().item(), std=t.std().item(), min=t.min().item(), max=t.max().item())


def pitch_stats_f64(pitch_list):
    x = np.concatenate([np.asarray(p, dtype=np.float32) for p in pitch_list]).astype(np.float64)
    x = x[x != 0]
    n = x.size
    mean = x.sum() / n
    var = ((x - mean) ** 2).sum() / (n - 1)
    return dict(mean=float(mean), std=float(np.sqrt(var)), min=float(x.min()), max=float(x.max()),
                count=int(n))


def normalize_pitch(pitch, mean, std):
    pitch = torch.as_tensor(np.array(pitch, dtype=np.float32))
    pitch -= mean
    pitch[pitch == -mean] = 0.0
    pitch /= std
    return pitch.numpy()
