"""Drop-in for ``scripts/dataset_processing/tts/compute_features.py`` on B200.

    python -m roar_b200.compute_features --feature_config_path=feature_22050.yaml \\
        --manifest_path=manifest.json --audio_dir=audio --feature_dir=features

Same arguments and the same output layout as the reference script
(``<feature_dir>/<feature_name>/<rel audio path>.pt``, ``compute_features.py:63-92``).  The feature
config is the reference's YAML shape -- entries with ``_target_: ...MelSpectrogramFeaturizer`` /
``EnergyFeaturizer`` / ``PitchFeaturizer``, ``${...}`` interpolation, a ``featurizers:`` mapping -- resolved
to the classes of ``roar_b200.featurizers`` by class name (no hydra / omegaconf needed).  Without a config
the defaults of those classes are used.  ``--num_workers`` is accepted and ignored: the reference spreads
single files over CPU processes, here every featurizer runs whole batches of utterances on the GPU.
"""
import argparse
import copy
import json
import re
from pathlib import Path
from typing import Any, Dict, List

import numpy as np

from . import featurizers as F

_CLASSES = {"MelSpectrogramFeaturizer": F.MelSpectrogramFeaturizer, "EnergyFeaturizer": F.EnergyFeaturizer,
            "PitchFeaturizer": F.PitchFeaturizer}


def _resolve(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """``${a.b}`` interpolation; a referenced node is resolved once and shared (so the energy featurizer
    receives the SAME spec featurizer object the config names)."""
    cache: Dict[str, Any] = {}

    def lookup(path):
        d = cfg
        for k in path.split("."):
            d = d[k]
        return d

    def walk(node, path=None):
        if isinstance(node, str):
            m = re.fullmatch(r"\$\{([^}]+)\}", node)
            if m:
                key = m.group(1)
                if key not in cache:
                    cache[key] = walk(copy.deepcopy(lookup(key)), key)
                return cache[key]
            return node
        if isinstance(node, list):
            return [walk(v) for v in node]
        if isinstance(node, dict):
            d = {k: walk(v) for k, v in node.items()}
            if "_target_" in d:
                cls_name = str(d.pop("_target_")).rsplit(".", 1)[-1]
                if cls_name not in _CLASSES:
                    raise ValueError(f"unsupported featurizer target {cls_name!r}; supported: {sorted(_CLASSES)}")
                obj = _CLASSES[cls_name](**d)
                if path is not None:
                    cache[path] = obj
                return obj
            return d
        return node

    out = {}
    for k in cfg:
        out[k] = cache[k] if k in cache else walk(cfg[k], k)
        cache.setdefault(k, out[k])
    return out


def default_featurizers() -> Dict[str, Any]:
    mel = F.MelSpectrogramFeaturizer()
    return {"mel_spec": mel, "energy": F.EnergyFeaturizer(spec_featurizer=mel),
            "pitch": F.PitchFeaturizer(voiced_prob_name="voiced_prob")}


def load_featurizers(feature_config_path) -> Dict[str, Any]:
    if feature_config_path is None:
        return default_featurizers()
    import yaml
    with open(feature_config_path) as f:
        cfg = yaml.safe_load(f)
    cfg = _resolve(cfg)
    if "featurizers" not in cfg:
        raise ValueError("feature config has no 'featurizers' mapping")
    return cfg["featurizers"]


def read_manifest(path: Path) -> List[Dict[str, Any]]:
    with open(path) as f:
        return [json.loads(line) for line in f if line.strip()]


def make_batches(lengths, max_samples: int) -> List[np.ndarray]:
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    batches, cur, tot = [], [], 0
    for i in order:
        if cur and tot + lengths[i] > max_samples:
            batches.append(np.array(cur)); cur, tot = [], 0
        cur.append(int(i)); tot += int(lengths[i])
    if cur:
        batches.append(np.array(cur))
    return batches


def run(featurizers: Dict[str, Any], entries: List[Dict[str, Any]], audio_dir: Path, feature_dir: Path,
        batch_audio_seconds: float = 2000.0) -> None:
    from concurrent.futures import ThreadPoolExecutor

    from . import host_io
    from .extract_sup_data import load_wav
    if not entries:
        return
    sr = next(iter(featurizers.values())).sample_rate
    paths = [str(F.get_abs_rel_paths(Path(e["audio_filepath"]), audio_dir)[0]) for e in entries]
    # batch composition from the wav headers; only one batch of audio (plus the one being prefetched) is ever
    # decoded in host memory, whatever the manifest size
    lengths = host_io.wav_probe(paths)["n_frames"].astype(np.int64)
    batches = make_batches(lengths, int(batch_audio_seconds * sr))

    def decode(b):
        return [load_wav(paths[i], sr) for i in b]

    with ThreadPoolExecutor(max_workers=1) as pool:
        fut = pool.submit(decode, batches[0])
        for k, b in enumerate(batches):
            sub_w = fut.result()
            if k + 1 < len(batches):
                fut = pool.submit(decode, batches[k + 1])
            sub_e = [entries[i] for i in b]
            for feature_name, featurizer in featurizers.items():
                featurizer.save_batch(sub_e, audio_dir, feature_dir, wavs=sub_w)


def main(argv=None):
    ap = argparse.ArgumentParser(description="Compute TTS features on B200.",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--feature_config_path", type=Path, default=None, help="Path to feature config file.")
    ap.add_argument("--manifest_path", required=True, type=Path, help="Path to training manifest.")
    ap.add_argument("--audio_dir", required=True, type=Path, help="Path to base directory with audio data.")
    ap.add_argument("--feature_dir", required=True, type=Path, help="Directory where features will be stored.")
    ap.add_argument("--num_workers", default=1, type=int, help="Accepted for compatibility; unused.")
    ap.add_argument("--batch_audio_seconds", default=2000.0, type=float, help="Audio per device batch.")
    args = ap.parse_args(argv)
    if not args.manifest_path.exists():
        raise ValueError(f"Manifest {args.manifest_path} does not exist.")
    if not args.audio_dir.exists():
        raise ValueError(f"Audio directory {args.audio_dir} does not exist.")
    featurizers = load_featurizers(args.feature_config_path)
    for name in featurizers:
        print(f"Computing: {name}")
    run(featurizers, read_manifest(args.manifest_path), args.audio_dir, args.feature_dir, args.batch_audio_seconds)


if __name__ == "__main__":
    main()
