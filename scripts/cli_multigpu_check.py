"""Run the extract_sup_data drop-in on N GPUs (torchrun) and on one, compare the printed pitch statistics
and the cache.   usage (on a box with >= 2 GPUs): python scripts/cli_multigpu_check.py [n_gpus]"""
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
from scipy.io import wavfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roar_b200 import synth  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
root = Path(tempfile.mkdtemp(prefix="roar_cli_"))
(root / "wavs").mkdir()
man = synth.corpus_manifest("C1", 40)
lines = []
for u in man:
    w = synth.synth_utterance(1234, u.utt_id, min(u.n_samples, 22050 * 4), 22050, u.speaker)
    p = root / "wavs" / f"spk{u.speaker}_utt{u.utt_id}.wav"
    wavfile.write(p, 22050, w)
    lines.append(json.dumps({"audio_filepath": str(p), "text": "a" * u.text_len, "duration": len(w) / 22050,
                             "speaker": int(u.speaker)}))
(root / "manifest.json").write_text("\n".join(lines) + "\n")
common = [f"manifest_filepath={root / 'manifest.json'}", "sup_data_types=[align_prior_matrix,pitch,voiced_mask,p_voiced,energy,log_mel]"]
env = dict(os.environ, PYTHONPATH=os.getcwd())
r1 = subprocess.run([sys.executable, "-m", "roar_b200.extract_sup_data", f"sup_data_path={root / 'sup1'}"] + common,
                    capture_output=True, text=True, env=env)
rn = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}",
                     "--master-addr", "127.0.0.1", "--master-port", "29533", "-m", "roar_b200.extract_sup_data",
                     f"sup_data_path={root / 'supN'}"] + common, capture_output=True, text=True, env=env)
print("single rc", r1.returncode, "multi rc", rn.returncode)
if r1.returncode or rn.returncode:
    print(r1.stderr[-2000:], rn.stderr[-3000:])
    sys.exit(1)
l1 = [ln for ln in r1.stdout.splitlines() if ln.startswith("PITCH_")]
ln_ = [ln for ln in rn.stdout.splitlines() if ln.startswith("PITCH_")]
print(l1, ln_)
s1 = json.load(open(root / "sup1" / "pitch_stats.json"))
sn = json.load(open(root / "supN" / "pitch_stats.json"))
worst = 0.0
for k in s1:
    for f in ("pitch_mean", "pitch_std", "pitch_min", "pitch_max"):
        worst = max(worst, abs(s1[k][f] - sn[k][f]) / max(1.0, abs(s1[k][f])))
nfiles = 0
for t in ("log_mel", "pitch", "voiced_mask", "p_voiced", "energy"):
    a = sorted(os.listdir(root / "sup1" / t)); b = sorted(os.listdir(root / "supN" / t))
    assert a == b and len(a) == len(man), (t, len(a), len(b))
    for f in a:
        assert torch.equal(torch.load(root / "sup1" / t / f), torch.load(root / "supN" / t / f)), (t, f)
        nfiles += 1
print(f"OK: {n_gpus}-GPU run == 1-GPU run: {nfiles} cache files identical, per-speaker stats rel diff {worst:.2e}, "
      f"speakers {sorted(s1)}")
