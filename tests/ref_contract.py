"""Test-only restatement of the CONSUMER side of the cache contract: how the
reference's ``TTSDataset`` reads the ``.pt`` files this package writes and turns them into a training batch.

Follows ``roar/collections/tts/data/dataset.py``:
  * ``__getitem__`` load branches ``:643-660`` (log_mel: ``torch.load`` then ``squeeze(0)``), ``:680-714`` (pitch /
    voiced_mask / p_voiced: ``torch.load(...).float()``), ``:716-741`` (pitch normalisation), ``:746-755`` (energy),
    tuple order ``:775-794``;
  * ``general_collate_fn`` ``:809-1026``: zero padding to the batch maximum, log-mel padded with
    ``finfo(dtype).tiny`` (``:854, 930-935``), prior zero-filled ``[B, max T, max N]`` (``:856-864, 942-945``);
  * ``join_data`` ``:799-807``: ``[audio, audio_lens, text, text_lens]`` + the sup types in the USER's order, each
    followed by ``<name>_lens`` when the type is ``WithLens`` (``tts/torch/tts_data_types.py``).
Pinned: ``tests/golden/collate_ref.npz`` / ``contract_ref.json`` were produced by executing the reference's own
``_collate_fn`` and type table (``tests/golden/make_golden_collate.py``); ``tests/test_contract.py`` holds this
restatement to them.  Never imported by the product.
"""
import json
import os
from pathlib import Path

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "contract_ref.json")) as _f:
    TYPE_TABLE = json.load(_f)
WITH_LENS = set(TYPE_TABLE["with_lens"])
MAIN = TYPE_TABLE["main"]


def general_padding(item, item_len, max_len, pad_value=0):
    if item_len < max_len:
        item = torch.nn.functional.pad(item, (0, max_len - item_len), value=pad_value)
    return item


def getitem_from_cache(audio, text_tokens, uid, folders, sup_types, prior_fn=None, speaker_id=None,
                       pitch_mean=None, pitch_std=None):
    """The 18-tuple of ``__getitem__`` for an utterance whose sup data is already cached under ``folders``."""
    types = set(sup_types)
    audio = torch.as_tensor(audio)
    text = torch.as_tensor(text_tokens).long()
    log_mel = log_mel_len = None
    if "log_mel" in types:
        log_mel = torch.load(Path(folders["log_mel"]) / f"{uid}.pt").squeeze(0)
        log_mel_len = torch.tensor(log_mel.shape[1]).long()
    prior = None
    if "align_prior_matrix" in types:
        prior = torch.from_numpy(prior_fn(len(text)))
    vals = {}
    for name in ("pitch", "voiced_mask", "p_voiced"):
        if name in types:
            vals[name] = torch.load(Path(folders[name]) / f"{uid}.pt").float()
    pitch, pitch_len = vals.get("pitch"), None
    if pitch is not None:
        pitch_len = torch.tensor(len(pitch)).long()
        if pitch_mean is not None and pitch_std is not None:
            pitch -= pitch_mean
            pitch[pitch == -pitch_mean] = 0.0
            pitch /= pitch_std
    energy = energy_len = None
    if "energy" in types:
        energy = torch.load(Path(folders["energy"]) / f"{uid}.pt").float()
        energy_len = torch.tensor(len(energy)).long()
    spk = torch.tensor(speaker_id).long() if "speaker_id" in types else None
    return (audio, torch.tensor(audio.shape[0]).long(), text, torch.tensor(len(text)).long(), log_mel, log_mel_len,
            None, prior, pitch, pitch_len, energy, energy_len, spk, vals.get("voiced_mask"), vals.get("p_voiced"),
            None, None, None)


def collate(batch, sup_types, pad_id=0):
    """``_collate_fn``: ``general_collate_fn`` + ``join_data`` -> the tuple a training step unpacks."""
    types = set(sup_types)
    cols = list(zip(*batch))
    audio_lens, token_lens, mel_lens, pitch_lens, energy_lens = cols[1], cols[3], cols[5], cols[9], cols[11]
    max_audio, max_tok = max(audio_lens).item(), max(token_lens).item()
    d = {"audio": torch.stack([general_padding(s[0], s[1].item(), max_audio) for s in batch]),
         "audio_lens": torch.stack(audio_lens),
         "text": torch.stack([general_padding(s[2], s[3].item(), max_tok, pad_value=pad_id) for s in batch]),
         "text_lens": torch.stack(token_lens)}
    if "log_mel" in types:
        tiny = torch.finfo(batch[0][4].dtype).tiny
        mx = max(mel_lens)
        d["log_mel"] = torch.stack([general_padding(s[4], s[5], mx, pad_value=tiny) for s in batch])
        d["log_mel_lens"] = torch.stack(mel_lens)
    if "align_prior_matrix" in types:
        pr = torch.zeros(len(batch), max(s[7].shape[0] for s in batch), max(s[7].shape[1] for s in batch))
        for i, s in enumerate(batch):
            pr[i, :s[7].shape[0], :s[7].shape[1]] = s[7]
        d["align_prior_matrix"] = pr
    if "pitch" in types:
        mx = max(pitch_lens).item()
        d["pitch"] = torch.stack([general_padding(s[8], s[9].item(), mx) for s in batch])
        d["pitch_lens"] = torch.stack(pitch_lens)
        if "voiced_mask" in types:
            d["voiced_mask"] = torch.stack([general_padding(s[13], s[9].item(), mx) for s in batch])
        if "p_voiced" in types:
            d["p_voiced"] = torch.stack([general_padding(s[14], s[9].item(), mx) for s in batch])
    if "energy" in types:
        mx = max(energy_lens).item()
        d["energy"] = torch.stack([general_padding(s[10], s[11].item(), mx) for s in batch])
        d["energy_lens"] = torch.stack(energy_lens)
    if "speaker_id" in types:
        d["speaker_id"] = torch.stack([s[12] for s in batch])
    out = []
    for name in MAIN + list(sup_types):
        out.append(d[name])
        if name in WITH_LENS:
            out.append(d[f"{name}_lens"])
    return tuple(out)
