"""On-GPU check of the pruned Viterbi kernel against the kernel that visits every in-band source
(ROAR_SUP_VITERBI=generic): f0 / voiced flag / voiced probability of a whole synthetic corpus must be
bit-identical.  usage: python scripts/viterbi_selfcheck.py [n_utts] [corpus]  -> one JSON line"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from roar_b200 import synth
from roar_b200.config import SupConfig
from roar_b200.extractor import SupDataExtractor


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 13100
    corpus = sys.argv[2] if len(sys.argv) > 2 else "C2"
    sr = synth.CORPORA[corpus]["sr"]
    cfg = SupConfig(highfreq=8000.0) if sr == 22050 else SupConfig(sample_rate=44100, n_fft=2048, hop_length=512)
    _, audio, offs, lens = synth.synth_corpus_device(corpus, "cuda", n_utts=n)
    res = {}
    for mode in ("fast", "generic"):
        if mode == "generic":
            os.environ["ROAR_SUP_VITERBI"] = "generic"
        else:
            os.environ.pop("ROAR_SUP_VITERBI", None)
        ex = SupDataExtractor(cfg)        # the switch is read when the handle is created
        b = ex.batch_from_device(audio, offs.cpu().numpy(), lens.cpu().numpy().astype(np.int64))
        f0, vf, vp, fo = ex.pyin(b)
        torch.cuda.synchronize()
        res[mode] = (f0.cpu(), vf.cpu(), vp.cpu())
        del ex, b
    a, g = res["fast"], res["generic"]
    frames = int(a[0].numel())
    out = {"corpus": corpus, "utterances": n, "frames": frames,
           "f0_mismatch_frames": int((a[0] != g[0]).sum()), "flag_mismatch_frames": int((a[1] != g[1]).sum()),
           "voiced_prob_mismatch_frames": int((a[2] != g[2]).sum()), "voiced_frames": int((a[1] != 0).sum())}
    print(json.dumps(out))
    return 0 if out["f0_mismatch_frames"] == 0 and out["flag_mismatch_frames"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
