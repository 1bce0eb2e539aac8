"""Quick on-GPU timing of every kernel on a slice of the config-2 corpus (diagnostics)."""
import ctypes
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from roar_b200 import synth
from roar_b200.config import SupConfig
from roar_b200.extractor import SupDataExtractor

NAMES = ["tile_offsets", "stft_mel", "pyin_cmnd", "pyin_probs", "len_sort", "viterbi", "backtrack", "prior",
         "stats", "fbank_norm", "pyin_energy", "pcm16"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    corpus = sys.argv[2] if len(sys.argv) > 2 else "C2"
    sr = synth.CORPORA[corpus]["sr"]
    cfg = SupConfig(highfreq=8000.0) if sr == 22050 else SupConfig(sample_rate=44100, n_fft=2048, hop_length=512)
    ex = SupDataExtractor(cfg)
    t0 = time.time()
    man, audio, offs, lens = synth.synth_corpus_device(corpus, "cuda", n_utts=n)
    torch.cuda.synchronize()
    print("synth", time.time() - t0, "s; audio-seconds", audio.numel() / sr, flush=True)
    b = ex.batch_from_device(audio, offs.cpu().numpy(), lens.cpu().numpy().astype(np.int64))
    tl = [u.text_len for u in man]
    secs = float(b.lens_host.sum()) / sr
    for rep in range(3):
        ex.lib.roar_sup_set_profiling(ex._h, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ex.extract(b, text_lens=tl, stats=ex.new_pitch_partials())
        e1.record()
        torch.cuda.synchronize()
        ms = (ctypes.c_double * 12)()
        cnt = (ctypes.c_int64 * 12)()
        ex.lib.roar_sup_profile_read(ex._h, ms, cnt, 1)
        tot = e0.elapsed_time(e1)
        print(json.dumps({"rep": rep, "total_ms": tot, "audio_s": secs, "x_realtime": secs / (tot / 1e3),
                          "kernels_ms": {k: round(v, 3) for k, v in zip(NAMES, ms)}}), flush=True)
        del out
    ex.lib.roar_sup_set_profiling(ex._h, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ex.extract(b, text_lens=tl, stats=ex.new_pitch_partials())
    e1.record()
    torch.cuda.synchronize()
    print("unprofiled total ms", e0.elapsed_time(e1), "x realtime", secs / (e0.elapsed_time(e1) / 1e3))
    f0 = out["pitch"]
    print("voiced frac", float((f0 != 0).float().mean()), "mean f0", float(f0[f0 != 0].mean()))


if __name__ == "__main__":
    main()
