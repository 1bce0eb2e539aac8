"""Step-category histogram of the Viterbi kernel on a corpus (diagnostic build, -DROAR_VIT_STATS):

    python -m roar_b200.build -DROAR_VIT_STATS --out=build/libroar_sup_stats.so
    ROAR_SUP_LIB=build/libroar_sup_stats.so python scripts/viterbi_step_stats.py [C2] [n_utts] [out.json]

Counts utterance-steps: sparse / dense steps, which source sets were scanned as a band and which walked as a
live list, candidate evaluations.  Used to decide where K3's instructions go."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    corpus = sys.argv[1] if len(sys.argv) > 1 else "C2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    import torch
    from roar_b200 import synth
    from roar_b200.config import SupConfig
    from roar_b200.extractor import SupDataExtractor
    cfg = SupConfig(highfreq=8000.0) if corpus != "C4" else SupConfig(sample_rate=44100, n_fft=2048, win_length=2048,
                                                                        hop_length=512, highfreq=None)
    ex = SupDataExtractor(cfg)
    man, audio, offs, lens = synth.synth_corpus_device(corpus, "cuda", n_utts=n)
    audio = torch.clamp(torch.round(audio * 32767.0), -32768, 32767) * (1.0 / 32768.0)
    b = ex.batch_from_device(audio, offs.cpu().numpy(), lens.cpu().numpy().astype(np.int64))
    out = (ctypes.c_uint64 * 16)()
    ex.lib.roar_sup_debug_counters(out, 16, 1)
    f0, vf, vp, fo = ex.pyin(b)
    torch.cuda.synchronize()
    ex.lib.roar_sup_debug_counters(out, 16, 0)
    c = [int(x) for x in out]
    steps = c[0] + c[1]
    names = ["sparse_steps", "dense_steps", "unvoiced_band_scans", "voiced_band_scans", "unvoiced_list_walks",
             "voiced_list_walks", "candidate_evaluations", "uniform_mode_steps", "dense_with_unvoiced_scan",
             "dense_with_voiced_scan", "sparse_with_voiced_scan", "dead_on_arrival_dense_steps", "reserved12", "steps_with_unvoiced_band_scan", "warp_steps_dead_segment_skipped",
             "warp_steps_band_scanned"]
    rep = {"corpus": corpus, "utterances": n, "frames": int(fo[-1]), "steps": steps,
           "voiced_prob_eq_1_frac": float((vp == 1).float().mean()), "voiced_frac": float((vf != 0).float().mean())}
    for k, v in zip(names, c):
        rep[k] = v
        rep[k + "_frac"] = v / max(1, steps)
    print(json.dumps(rep, indent=1))
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
