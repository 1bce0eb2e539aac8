"""Golden fixtures for the consumer contract, produced by the REFERENCE's own
``TTSDataset`` code.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_collate.py

``roar/collections/tts/data/dataset.py`` imports librosa / hydra-era modules this image lacks, so the module is
loaded by file path with stub modules for those imports; the pieces exercised here --
``TTSDataset.general_collate_fn`` / ``join_data`` / ``_collate_fn`` (``dataset.py:799-1031``),
``TTSDataset.filter_files`` (``:367-406``) with the real ``tts_data_types.py`` and ``tts_dataset_utils.py`` --
touch none of the stubs: it is the reference's code that executes.

Outputs:
  * ``collate_ref.npz``      the joined batch tuple the reference builds from three per-utterance 18-tuples
                             (taken from ``supdata_oracle.npz`` + the reference's own prior) for two
                             ``sup_data_types`` orders, plus those inputs;
  * ``contract_ref.json``    the data-type table (names, which carry ``_lens``) and ``filter_files`` cases.
"""
import importlib.util
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

ORDERS = {
    "fastpitch": ["align_prior_matrix", "pitch", "energy"],                       # ds_for_fastpitch_align.yaml:5
    "all": ["log_mel", "align_prior_matrix", "pitch", "voiced_mask", "p_voiced", "energy", "speaker_id"],
}


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_dataset_module():
    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {"__init__": lambda self, *a, **kw: None})

    for name in ["librosa", "roar", "roar.collections", "roar.collections.asr", "roar.collections.asr.parts",
                 "roar.collections.asr.parts.preprocessing", "roar.collections.asr.parts.preprocessing.features",
                 "roar.collections.asr.parts.preprocessing.segment", "roar.collections.common",
                 "roar.collections.common.tokenizers", "roar.collections.common.tokenizers.text_to_speech",
                 "roar.collections.common.tokenizers.text_to_speech.tts_tokenizers", "roar.collections.tts",
                 "roar.collections.tts.parts", "roar.collections.tts.parts.utils", "roar.collections.tts.torch",
                 "roar.core", "roar.core.classes", "roar.utils"]:
        sys.modules[name] = _Any(name)
    logmod = types.ModuleType("roar.utils.logging")
    for fn in ("info", "debug", "warning", "error"):
        setattr(logmod, fn, lambda *a, **k: None)
    sys.modules["roar.utils"].logging = logmod
    sys.modules["roar.utils.logging"] = logmod
    # the real contract modules
    load_by_path("roar.collections.tts.parts.utils.tts_dataset_utils",
                 f"{REF}/roar/collections/tts/parts/utils/tts_dataset_utils.py")
    load_by_path("roar.collections.tts.torch.tts_data_types", f"{REF}/roar/collections/tts/torch/tts_data_types.py")
    return load_by_path("ref_tts_dataset", f"{REF}/roar/collections/tts/data/dataset.py")


def sample_tuples(utils):
    """Three __getitem__ tuples (dataset.py:775-794) from the committed oracle vectors."""
    g = np.load(f"{HERE}/supdata_oracle.npz")
    text_lens = [7, 12, 9]
    keep = [173, 150, 120]          # frames kept per utterance: unequal lengths exercise every padding rule
    tuples, inputs = [], {}
    for i in range(3):
        k = keep[i]
        audio = torch.from_numpy(g[f"audio{i}"][:k * 256])
        text = torch.arange(1, text_lens[i] + 1).long()
        log_mel = torch.from_numpy(g[f"logmel{i}"]).squeeze(0)[:, :k]
        pitch = torch.from_numpy(g[f"f0_{i}"]).float()[:k]
        voiced = torch.from_numpy(g[f"vflag{i}"]).float()[:k]
        pvoiced = torch.from_numpy(g[f"vprob{i}"]).float()[:k]
        energy = torch.from_numpy(g[f"energy{i}"]).float()[:k]
        prior = torch.from_numpy(utils.beta_binomial_prior_distribution(text_lens[i], log_mel.shape[1]))
        tuples.append((audio, torch.tensor(audio.shape[0]).long(), text, torch.tensor(len(text)).long(),
                       log_mel, torch.tensor(log_mel.shape[1]).long(), None, prior,
                       pitch, torch.tensor(len(pitch)).long(), energy, torch.tensor(len(energy)).long(),
                       torch.tensor(i % 2).long(), voiced, pvoiced, None, None, None))
        inputs[f"text{i}"] = text.numpy()
        inputs[f"keep{i}"] = np.array(k)
        inputs[f"prior{i}"] = prior.numpy()
    return tuples, inputs


def main():
    ds = reference_dataset_module()
    utils = sys.modules["roar.collections.tts.parts.utils.tts_dataset_utils"]
    types_mod = sys.modules["roar.collections.tts.torch.tts_data_types"]
    tuples, out = sample_tuples(utils)
    for tag, order in ORDERS.items():
        obj = object.__new__(ds.TTSDataset)                   # no __init__: only the collate state is needed
        obj.sup_data_types = [types_mod.DATA_STR2DATA_CLASS[t] for t in order]
        obj.sup_data_types_set = set(obj.sup_data_types)
        obj.text_tokenizer_pad_id = 0
        joined = obj._collate_fn(tuples)
        out[f"{tag}__n"] = np.array(len(joined))
        for k, t in enumerate(joined):
            out[f"{tag}__{k}"] = t.numpy()
        print(tag, [tuple(t.shape) for t in joined])
    np.savez_compressed(f"{HERE}/collate_ref.npz", **out)

    table = {"main": [t.name for t in types_mod.MAIN_DATA_TYPES],
             "with_lens": sorted(t.name for t in types_mod.DATA_STR2DATA_CLASS.values()
                                 if issubclass(t, types_mod.WithLens)),
             "valid_sup": [t.name for t in types_mod.VALID_SUPPLEMENTARY_DATA_TYPES],
             "orders": ORDERS, "filter_files": []}
    data = [{"audio_filepath": f"/d/spk{k % 2}/u{k}.wav", "duration": dur}
            for k, dur in enumerate([0.05, 3.0, 30.0, 7.5, 1.0])]
    with tempfile.TemporaryDirectory() as tmp:
        ig = os.path.join(tmp, "ignore.pkl")
        with open(ig, "wb") as f:
            pickle.dump(["/d/spk1/u3.wav", "u4"], f)
        for case in (dict(ignore=False, min_duration=0.1, max_duration=20.0, total=41.55),
                     dict(ignore=True, min_duration=0.1, max_duration=None, total=41.55),
                     # total_duration None = some entry lacks a duration: no duration pruning at all.  (With an
                     # ignore file as well the reference crashes -- `pruned_duration += ...` on None, :391 -- so
                     # that combination has no reference behaviour to pin.)
                     dict(ignore=False, min_duration=0.1, max_duration=20.0, total=None)):
            kept = ds.TTSDataset.filter_files([dict(d) for d in data], ig if case["ignore"] else None,
                                              case["min_duration"], case["max_duration"], case["total"])
            table["filter_files"].append(dict(case, data=data, ignore_list=["/d/spk1/u3.wav", "u4"],
                                              kept=[d["audio_filepath"] for d in kept]))
    with open(f"{HERE}/contract_ref.json", "w") as f:
        json.dump(table, f, indent=1)
    print(json.dumps(table["filter_files"], indent=1)[:600])


if __name__ == "__main__":
    main()
