"""Host side of the extraction path: packed ragged batches in HBM -> libroar_sup.so kernels.

Mirrors what ``TTSDataset.__getitem__`` computes per utterance
(``roar/collections/tts/data/dataset.py:643-755``) -- ``get_log_mel``, ``librosa.pyin``, energy,
``beta_binomial_prior_distribution`` -- but for a whole batch of utterances per call.  PyTorch is used
for device memory and streams only; all arithmetic happens in the CUDA library.
"""
import ctypes
import os
from contextlib import nullcontext as _nullcontext
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import SupConfig

ALIGN = 4  # samples; utterance starts on 16-byte boundaries enable the TMA bulk-copy path
_H2D_MODE = os.environ.get("ROAR_SUP_H2D", "pinned")      # diagnostics: "pageable" / "pageable_async"


@dataclass
class PackedBatch:
    """Utterance i = ``audio[sample_off[i] : sample_off[i] + sample_len[i]]`` (float32, device)."""
    audio: torch.Tensor
    sample_off: torch.Tensor      # int64 [n] device
    sample_len: torch.Tensor      # int32 [n] device
    lens_host: np.ndarray         # int64 [n]
    offs_host: np.ndarray         # int64 [n]

    @property
    def n_utts(self) -> int:
        return int(self.lens_host.shape[0])

    @property
    def total_samples(self) -> int:
        return int(self.audio.shape[0])


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


_NP2TORCH = {np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32, np.dtype(np.float64): torch.float64,
             np.dtype(np.float32): torch.float32}


class MetaStager:
    """Small index arrays (offsets, lengths, frame prefix sums, groups) host -> device without the copy engine.

    A ``cudaMemcpyAsync`` of a few KB from pinned memory queues on the host-to-device copy engine behind whatever
    large audio copy is in flight there (hundreds of MB, several ms), and the kernels that need the array wait with
    it: measured, the streamed extraction was 25 % slower that way than with blocking pageable copies.  Here the
    arrays are written into a pinned arena and fetched by a tiny kernel that reads the arena directly
    (``roar_sup_upload``), in stream order, so the launching thread never blocks either.  The arena is a ring of
    segments; a segment is reused only after the uploads issued from it have completed (one event per stream that
    used it)."""

    SEG, NSEG = 1 << 20, 8

    def __init__(self, lib, handle, device):
        self.lib, self.h, self.device = lib, handle, device
        self.arena = torch.empty(self.SEG * self.NSEG, dtype=torch.uint8, pin_memory=True)
        self.view = self.arena.numpy()
        self.seg, self.off = 0, 0
        self.events = [[] for _ in range(self.NSEG)]
        self.streams = {}
        self.big = []

    def _leave_segment(self):
        self.events[self.seg] = []
        for st in self.streams.values():
            ev = torch.cuda.Event()
            ev.record(st)
            self.events[self.seg].append(ev)
        self.streams = {}
        self.seg = (self.seg + 1) % self.NSEG
        self.off = 0
        for ev in self.events[self.seg]:
            ev.synchronize()
        self.events[self.seg] = []

    def upload(self, a: np.ndarray, dtype) -> torch.Tensor:
        a = np.ascontiguousarray(a, dtype=dtype)
        tdt = _NP2TORCH[a.dtype]
        if a.size == 0:
            return torch.empty(0, dtype=tdt, device=self.device)
        nbytes = a.nbytes
        padded = (nbytes + 15) & ~15
        stream = torch.cuda.current_stream(self.device)
        dst = torch.empty(padded, dtype=torch.uint8, device=self.device)
        if padded > self.SEG:                      # oversize: its own pinned block, kept until the upload is done
            host = torch.empty(padded, dtype=torch.uint8, pin_memory=True)
            host.numpy()[:nbytes] = a.view(np.uint8).reshape(-1)
            src_ptr = host.data_ptr()
        else:
            if self.off + padded > self.SEG:
                self._leave_segment()
            o = self.seg * self.SEG + self.off
            self.view[o:o + nbytes] = a.view(np.uint8).reshape(-1)
            src_ptr = self.arena.data_ptr() + o
            self.off += padded
            self.streams[stream.cuda_stream] = stream
            host = None
        _lib.check(self.lib.roar_sup_upload(self.h, ctypes.c_void_p(src_ptr), ctypes.c_void_p(dst.data_ptr()), padded,
                                            ctypes.c_void_p(stream.cuda_stream)))
        if host is not None:
            ev = torch.cuda.Event()
            ev.record(stream)
            self.big = [(e, t) for e, t in self.big if not e.query()] + [(ev, host)]
        return dst[:nbytes].view(tdt)


def pack_layout(lens: np.ndarray, align: int = ALIGN):
    lens = np.asarray(lens, dtype=np.int64)
    padded = (lens + align - 1) // align * align
    offs = np.zeros(len(lens), dtype=np.int64)
    if len(lens) > 1:
        offs[1:] = np.cumsum(padded)[:-1]
    return offs, int(padded.sum())


class SupDataExtractor:
    """One handle (tables in HBM) per device and configuration."""

    def __init__(self, cfg: SupConfig = SupConfig(), device: Optional[torch.device] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.RoarSupError("roar_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._c = cfg.to_c()
        h = ctypes.c_void_p()
        _lib.check(self.lib.roar_sup_create(ctypes.byref(self._c), self.device.index or 0, ctypes.byref(h)))
        self._h = h
        geo = (ctypes.c_int32 * 8)()
        self.lib.roar_sup_pyin_geometry(self._h, ctypes.byref(geo))
        (self.min_period, self.max_period, self.n_pitch_bins, self.transition_width,
         self.pyin_hop, self.pyin_win, self.kmax, self.n_transition_rows) = [int(x) for x in geo]
        self._ws: Optional[torch.Tensor] = None
        self._meta = MetaStager(self.lib, self._h, self.device)
        self.kernel_launches = 0  # launches of our kernels issued through this object
        self._side = None
        self._ws_side = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.roar_sup_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _up(self, a, dtype) -> torch.Tensor:
        """Small index array -> device (see MetaStager); ``ROAR_SUP_H2D=pageable`` keeps the blocking copy."""
        if _H2D_MODE == "pageable":
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self.device)
        self.kernel_launches += 1
        return self._meta.upload(np.asarray(a), dtype)

    # ------------------------------------------------------------------ geometry
    def num_frames(self, n_samples):
        """``T = 1 + L // hop`` (``torch.stft(center=True)``, dataset.py:324-333)."""
        n = np.asarray(n_samples, dtype=np.int64)
        if self.cfg.exact_pad:
            p = (self.cfg.n_fft - self.cfg.hop) // 2
            return (n + 2 * p - self.cfg.n_fft) // self.cfg.hop + 1
        return 1 + n // self.cfg.hop

    def pyin_num_frames(self, n_samples):
        return 1 + np.asarray(n_samples, dtype=np.int64) // self.pyin_hop

    def mel_filterbank(self) -> np.ndarray:
        """float32 ``[n_mels, n_fft//2+1]`` -- what ``self.fb`` holds in the reference."""
        out = np.zeros((self.cfg.n_mels, self.cfg.n_fft // 2 + 1), dtype=np.float32)
        _lib.check(self.lib.roar_sup_host_mel_filterbank(ctypes.byref(self._c), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    # ------------------------------------------------------------------ batches
    def pack(self, wavs: Sequence, pinned: bool = True) -> PackedBatch:
        """Host waveforms (numpy / CPU tensors, float32) -> one packed device buffer (one H2D copy)."""
        lens = np.array([int(w.shape[0]) for w in wavs], dtype=np.int64)
        offs, total = pack_layout(lens)
        host = torch.zeros(total, dtype=torch.float32, pin_memory=pinned)
        hv = host.numpy()
        for w, o, n in zip(wavs, offs, lens):
            hv[o:o + n] = w.numpy() if isinstance(w, torch.Tensor) else np.asarray(w, dtype=np.float32)
        return self.pack_from_host_buffer(host, offs, lens)

    def pack_from_host_buffer(self, host: torch.Tensor, offs: np.ndarray, lens: np.ndarray) -> PackedBatch:
        audio = host.to(self.device, non_blocking=True)
        return self.batch_from_device(audio, offs, lens)

    def pcm16_to_f32(self, pcm: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device int16 PCM -> float32 ``x / 2**15`` (``segment.py:140-153``), on the current stream."""
        assert pcm.is_cuda and pcm.dtype == torch.int16 and pcm.is_contiguous()
        if out is None:
            out = torch.empty(pcm.numel(), dtype=torch.float32, device=pcm.device)
        _lib.check(self.lib.roar_sup_pcm16_to_f32(self._h, _ptr(pcm), pcm.numel(), _ptr(out), self._stream()))
        self.kernel_launches += 1
        return out

    def pack_pcm16_from_host_buffer(self, host: torch.Tensor, offs: np.ndarray, lens: np.ndarray) -> PackedBatch:
        """Packed 16-bit PCM in (pinned) host memory -> device float32 batch: the copy moves 2 bytes per
        sample, the integer -> float step of the wav decoder runs on the GPU."""
        assert host.dtype == torch.int16
        audio = self.pcm16_to_f32(host.to(self.device, non_blocking=True))
        return self.batch_from_device(audio, offs, lens)

    def batch_from_device(self, audio: torch.Tensor, offs: np.ndarray, lens: np.ndarray) -> PackedBatch:
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()
        return PackedBatch(audio, self._up(offs, np.int64), self._up(lens, np.int32),
                           np.asarray(lens, dtype=np.int64), np.asarray(offs, dtype=np.int64))

    def trim(self, batch: PackedBatch, top_db: float = 60.0, ref: Optional[float] = None, frame_length: int = 2048,
             hop_length: int = 512) -> PackedBatch:
        """``librosa.effects.trim`` for every utterance of the batch (``segment.py:76-88``): returns a batch
        over the SAME audio buffer with narrowed offsets / lengths (``ref=None`` = ``np.max`` like the
        reference's ``trim_ref`` default).  ``batch.trim_bounds`` holds the [start, end) pairs."""
        if batch.n_utts == 0:
            return batch
        n = batch.n_utts
        st = torch.empty(n, dtype=torch.int64, device=self.device)
        en = torch.empty(n, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.roar_sup_trim(self._h, _ptr(batch.audio), _ptr(batch.sample_off), _ptr(batch.sample_len), n,
                                          int(batch.lens_host.max()), float(top_db), float(ref) if ref else 0.0,
                                          int(frame_length), int(hop_length), _ptr(st), _ptr(en), self._stream()))
        self.kernel_launches += 1
        st_h, en_h = st.cpu().numpy(), en.cpu().numpy()
        out = PackedBatch(batch.audio, batch.sample_off + st, (en - st).to(torch.int32), en_h - st_h,
                          batch.offs_host + st_h)
        out.trim_bounds = np.stack([st_h, en_h], axis=1)
        return out

    def _check_lengths(self, batch: PackedBatch, what: str):
        pad = (self.cfg.n_fft - self.cfg.hop) // 2 if self.cfg.exact_pad else self.cfg.n_fft // 2
        if batch.n_utts and int(batch.lens_host.min()) <= pad:
            # torch.stft's reflect padding raises for these inputs in the reference
            raise ValueError(f"{what}: every utterance must be longer than the reflect padding ({pad} samples)")

    def _frame_off(self, T: np.ndarray) -> (torch.Tensor, np.ndarray):
        fo = np.zeros(len(T) + 1, dtype=np.int64)
        np.cumsum(T, out=fo[1:])
        return self._up(fo, np.int64), fo

    def _workspace(self, n_utts: int, total_samples: int, frames: int, extra: int = 0) -> torch.Tensor:
        need = int(self.lib.roar_sup_workspace_bytes(self._h, n_utts, total_samples, frames)) + int(extra)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    # ------------------------------------------------------------------ kernels
    def log_mel_energy(self, batch: PackedBatch, want_log_mel: bool = True, want_energy: bool = True):
        """-> (log_mel flat float32 [n_mels * sum T] or None, energy [sum T] or None, frame_off host int64).
        Utterance i's log-mel is ``log_mel[n_mels*fo[i] : n_mels*fo[i+1]].view(n_mels, T_i)``."""
        if batch.n_utts == 0:   # empty manifest shard: nothing to launch
            e = torch.empty(0, dtype=torch.float32, device=self.device)
            return (e if want_log_mel else None), (e.clone() if want_energy else None), np.zeros(1, dtype=np.int64)
        self._check_lengths(batch, "log_mel_energy")
        T = self.num_frames(batch.lens_host)
        d_fo, fo = self._frame_off(T)
        total = int(fo[-1])
        lm = torch.empty(self.cfg.n_mels * total, dtype=torch.float32, device=self.device) if want_log_mel else None
        en = torch.empty(total, dtype=torch.float32, device=self.device) if want_energy else None
        # room for the tile -> utterance map behind the tile prefix sums (tiles hold >= 4 frames)
        ws = self._workspace(batch.n_utts, batch.total_samples, 0, extra=total + 8 * batch.n_utts + 4096)
        _lib.check(self.lib.roar_sup_logmel_energy(
            self._h, _ptr(batch.audio), _ptr(batch.sample_off), _ptr(batch.sample_len), batch.n_utts,
            _ptr(d_fo), total, _ptr(lm), _ptr(en), _ptr(ws), ws.numel(), self._stream()))
        self.kernel_launches += 3      # tile offsets, tile map, stft_mel
        return lm, en, fo

    def pyin(self, batch: PackedBatch):
        """-> (f0, voiced_flag, voiced_prob) flat float32 [sum T], frame_off host int64."""
        if batch.n_utts == 0:
            e = torch.empty(0, dtype=torch.float32, device=self.device)
            return e, e.clone(), e.clone(), np.zeros(1, dtype=np.int64)
        T = self.pyin_num_frames(batch.lens_host)
        d_fo, fo = self._frame_off(T)
        total = int(fo[-1])
        f0 = torch.empty(total, dtype=torch.float32, device=self.device)
        vf = torch.empty(total, dtype=torch.float32, device=self.device)
        vp = torch.empty(total, dtype=torch.float32, device=self.device)
        ws = self._workspace(batch.n_utts, batch.total_samples, total)
        _lib.check(self.lib.roar_sup_pyin(
            self._h, _ptr(batch.audio), _ptr(batch.sample_off), _ptr(batch.sample_len), batch.n_utts,
            _ptr(d_fo), total, int(T.max()), _ptr(f0), _ptr(vf), _ptr(vp), _ptr(ws), ws.numel(), self._stream()))
        # 2 x tile offsets, tile map, energy, cmnd, probs, 3 x length sort, viterbi, backtrack
        self.kernel_launches += 11
        return f0, vf, vp, fo

    def align_prior(self, text_lens, mel_lens, scaling_factor: float = 1.0):
        """-> (prior flat float32, out_off host int64 [n+1]); utterance i is
        ``prior[oo[i]:oo[i+1]].view(mel_len_i, text_len_i)`` (tts_dataset_utils.py:140-149)."""
        tl = np.asarray(text_lens, dtype=np.int32)
        ml = np.asarray(mel_lens, dtype=np.int32)
        oo = np.zeros(len(tl) + 1, dtype=np.int64)
        np.cumsum(tl.astype(np.int64) * ml.astype(np.int64), out=oo[1:])
        out = torch.empty(int(oo[-1]), dtype=torch.float32, device=self.device)
        d_tl = self._up(tl, np.int32)
        d_ml = self._up(ml, np.int32)
        d_oo = self._up(oo, np.int64)
        _lib.check(self.lib.roar_sup_align_prior(self._h, _ptr(d_tl), _ptr(d_ml), len(tl), _ptr(d_oo),
                                                 int(ml.max()) if len(ml) else 0, float(scaling_factor),
                                                 _ptr(out), self._stream()))
        self.kernel_launches += (len(tl) + 65534) // 65535
        return out, oo

    def align_prior_interp(self, text_lens, mel_lens, round_mel_len_to: int = 50, round_text_len_to: int = 10):
        """``BetaBinomialInterpolator.__call__(mel_len, text_len)`` for a batch
        (tts_dataset_utils.py:69-92): same return layout as :meth:`align_prior`."""
        tl = np.asarray(text_lens, dtype=np.int32)
        ml = np.asarray(mel_lens, dtype=np.int32)
        oo = np.zeros(len(tl) + 1, dtype=np.int64)
        np.cumsum(tl.astype(np.int64) * ml.astype(np.int64), out=oo[1:])
        out = torch.empty(int(oo[-1]), dtype=torch.float32, device=self.device)
        d_tl = self._up(tl, np.int32)
        d_ml = self._up(ml, np.int32)
        d_oo = self._up(oo, np.int64)
        _lib.check(self.lib.roar_sup_align_prior_interp(self._h, _ptr(d_tl), _ptr(d_ml), len(tl), _ptr(d_oo),
                                                        int(ml.max()) if len(ml) else 0, int(round_mel_len_to),
                                                        int(round_text_len_to), _ptr(out), self._stream()))
        self.kernel_launches += (len(tl) + 65534) // 65535
        return out, oo

    def new_pitch_partials(self, n_groups: int = 1) -> torch.Tensor:
        out = torch.empty(n_groups, 5, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.roar_sup_pitch_partials_init(self._h, _ptr(out), n_groups, self._stream()))
        self.kernel_launches += 1
        return out

    def pitch_partials(self, f0: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Accumulate (sum, sumsq, count, min, max) over ``f0 != 0`` into ``out[0]``."""
        if out is None:
            out = self.new_pitch_partials(1)
        _lib.check(self.lib.roar_sup_pitch_partials(self._h, _ptr(f0), f0.numel(), _ptr(out), self._stream()))
        self.kernel_launches += 1
        return out

    def pitch_partials_grouped(self, f0: torch.Tensor, frame_off: np.ndarray, groups, n_groups: int,
                               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = self.new_pitch_partials(n_groups)
        d_fo = self._up(frame_off, np.int64)
        d_g = self._up(groups, np.int32)
        _lib.check(self.lib.roar_sup_pitch_partials_grouped(self._h, _ptr(f0), _ptr(d_fo), _ptr(d_g), len(groups),
                                                            n_groups, _ptr(out), self._stream()))
        self.kernel_launches += 1
        return out

    # ------------------------------------------------------------------ whole hot path
    def extract(self, batch: PackedBatch, text_lens=None,
                types=("log_mel", "align_prior_matrix", "pitch", "voiced_mask", "p_voiced", "energy"),
                stats: Optional[torch.Tensor] = None) -> Dict[str, object]:
        """All supplementary data of one packed batch.  Returns flat device tensors + host offsets."""
        types = set(types)
        out: Dict[str, object] = {}
        want_pitch = bool(types & {"pitch", "voiced_mask", "p_voiced"})
        want_spec = bool(types & {"log_mel", "energy", "align_prior_matrix"})
        # The spectral kernels and the pYIN chain are independent: when both are requested the spectral
        # side runs on a second stream (its own workspace), forked from / joined to the caller's stream,
        # so its CTAs fill the SMs the latency-bound Viterbi leaves idle.
        main = torch.cuda.current_stream(self.device)
        overlap = want_pitch and want_spec and batch.n_utts > 0 and os.environ.get("ROAR_SUP_NO_OVERLAP") != "1"
        side = self._side_stream() if overlap else None
        if side is not None:
            side.wait_stream(main)
        if want_spec:
            with torch.cuda.stream(side) if side is not None else _nullcontext():
                if side is not None:
                    self._ws, self._ws_side = self._ws_side, self._ws      # spectral side uses its own workspace
                try:
                    lm, en, fo = self.log_mel_energy(batch, "log_mel" in types, "energy" in types) \
                        if ("log_mel" in types or "energy" in types) else (None, None, None)
                    if fo is None:
                        fo = np.concatenate([[0], np.cumsum(self.num_frames(batch.lens_host))])
                    out["log_mel"], out["energy"], out["frame_off"] = lm, en, fo
                    if "align_prior_matrix" in types:
                        if text_lens is None:
                            raise ValueError("align_prior_matrix needs text_lens")
                        mel_lens = np.diff(out["frame_off"])
                        out["align_prior_matrix"], out["prior_off"] = self.align_prior(text_lens, mel_lens)
                finally:
                    if side is not None:
                        self._ws, self._ws_side = self._ws_side, self._ws
        if want_pitch:
            f0, vf, vp, pfo = self.pyin(batch)
            out["pitch"], out["voiced_mask"], out["p_voiced"], out["pitch_frame_off"] = f0, vf, vp, pfo
            if stats is not None:
                self.pitch_partials(f0, stats)
        if side is not None:
            main.wait_stream(side)
            for k in ("log_mel", "energy", "align_prior_matrix"):
                t = out.get(k)
                if isinstance(t, torch.Tensor):
                    t.record_stream(main)      # allocated on the side stream, consumed on the caller's
        return out


def split_frames(flat: torch.Tensor, frame_off: np.ndarray, rows: int = 1) -> List[torch.Tensor]:
    """Per-utterance views of a flat output (``rows`` = n_mels for log-mel)."""
    out = []
    for i in range(len(frame_off) - 1):
        a, b = int(frame_off[i]), int(frame_off[i + 1])
        v = flat[rows * a: rows * b]
        out.append(v.view(rows, b - a) if rows > 1 else v)
    return out


def finalize_pitch_stats(partials: torch.Tensor) -> Dict[str, float]:
    """(sum, sumsq, count, min, max) -> mean / unbiased std / min / max
    (``extract_sup_data.py:8-13``: ``torch.std`` default is the n-1 estimator)."""
    s, q, n, mn, mx = [float(x) for x in partials.reshape(-1, 5)[0].tolist()]
    if n < 1:
        return dict(pitch_mean=float("nan"), pitch_std=float("nan"), pitch_min=float("nan"),
                    pitch_max=float("nan"), count=0)
    mean = s / n
    var = (q - s * s / n) / (n - 1) if n > 1 else float("nan")
    return dict(pitch_mean=mean, pitch_std=float(np.sqrt(max(var, 0.0))) if n > 1 else float("nan"),
                pitch_min=mn, pitch_max=mx, count=int(n))
