// K1  stft_mel_energy: reflect-padded framed STFT (shared-memory Stockham real FFT over a window
// table) fused with |X|, the sparse banded mel projection, log guard and per-frame L2 energy.
//
// Replaces (reference paths): TTSDataset.get_spec / get_log_mel / energy
//   roar/collections/tts/data/dataset.py:324-333, 524-537, 751-753
// and the STFT..log part of FilterbankFeatures.forward
//   roar/collections/asr/parts/preprocessing/features.py:384-434.
//
// Work decomposition: one CTA = one tile of FT consecutive frames of one utterance.  The tile's
// audio span ((FT-1)*hop + n_fft samples) is staged once in shared memory -- by ONE 1-D TMA bulk
// copy (cp.async.bulk + mbarrier) when the span is interior and 16-B aligned, by mirrored
// per-thread loads at the utterance edges -- so every sample is read from HBM/L2 once per tile.
// P = M/8 threads own a frame (M = n_fft/2 complex points, radix-8 Stockham passes, real-FFT
// untangle), G = blockDim/P frames are in flight per pass.  The result tile [n_mels, FT] is
// staged in shared memory and written as row segments (coalesced).
//
// Algorithmic HBM bytes per frame: 4*hop in + 4*(n_mels+1) out.
#pragma once
#include "common.cuh"
#include "fft.cuh"
#include "../../include/roar_sup.h"

namespace roar {

struct StftParams {
  // batch
  const float* audio;
  const int64_t* sample_off;
  const int32_t* sample_len;
  const int64_t* frame_off;   // [n_utts+1]
  const int32_t* tile_off;    // [n_utts+1] prefix sum of ceil(T_i / FT)
  const int32_t* tile_map;    // [n_tiles] tile -> utterance (optional: null = binary search of tile_off)
  int32_t n_utts;
  float* logmel;              // may be null
  float* energy;              // may be null
  int64_t out_utt_stride;     // 0 => ragged: block of utterance i starts at n_mels*frame_off[i], row stride T_i
  int64_t out_row_stride;
  // geometry
  int32_t n_fft, hop, M, n_bins, n_mels, FT, span, pad_left;
  int32_t P, G;               // threads per frame, frames in flight
  // options
  float floor_, mag_power, log_guard, preemph;
  int32_t log_mode, has_preemph, use_tma, preemph_after_pad, energy_mode;
  // tables (device memory)
  const float* window;        // [n_fft]
  const cf32* tw;             // per-pass Stockham twiddles (fft.cuh layout), read through L1
  const cf32* tw_post;        // [M+1]  W_{n_fft}^k, read through L1
  const int32_t* mel_start;   // [n_mels]
  const int32_t* mel_count;
  const int32_t* mel_offset;
  const float* mel_w;         // packed band weights
  int32_t mel_nw;
  int32_t tw_total;           // fft_tw_total(M), computed once on the host (0: derive it)
  int32_t part_per_slot;      // energy partials per frame slot in StftSmem::part (0: one per thread = P)
};

// shared-memory carve-up (all offsets in bytes, 16-B aligned)
struct StftSmem {
  float* audio;     // [span]
  float* window;    // [n_fft]
  cf32* tw;         // [tw_total] per-pass Stockham twiddles
  cf32* bufA;       // [G*MP]  MP = pidx(M): padded slot stride
  cf32* bufB;       // [G*MP]  (spec[G][n_bins] float aliases the buffer not holding the FFT result)
  float* part;      // [blockDim] energy partials
  float* out;       // [n_mels][FT+1] staging
  float* en;        // [FT]
  int32_t* mel_start; int32_t* mel_count; int32_t* mel_offset;  // [n_mels]
  float* mel_w;     // [mel_nw]
  unsigned long long* mbar;
};

HD size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

HD size_t stft_smem_carve(const StftParams& p, int nthreads, unsigned char* base, StftSmem* s) {
  size_t o = 0;
#define CARVE(field, type, count) { if (s) s->field = (type*)(base + o); o = align16(o + sizeof(type) * (size_t)(count)); }
  CARVE(mbar, unsigned long long, 2)
  CARVE(audio, float, p.span)
  CARVE(window, float, p.n_fft)
  CARVE(tw, cf32, (p.tw_total ? p.tw_total : fft_tw_total(p.M)) + 1)
  CARVE(bufA, cf32, (size_t)p.G * pidx(p.M) + 8)
  CARVE(bufB, cf32, (size_t)p.G * pidx(p.M) + 8)
  CARVE(part, float, nthreads)
  CARVE(out, float, (size_t)p.n_mels * (p.FT + 1))
  CARVE(en, float, p.FT)
  CARVE(mel_start, int32_t, p.n_mels)
  CARVE(mel_count, int32_t, p.n_mels)
  CARVE(mel_offset, int32_t, p.n_mels)
  CARVE(mel_w, float, p.mel_nw)
#undef CARVE
  return o;
}

struct StftTile {
  int32_t utt, t0, T, nf;     // frames [t0, t0+nf) of utterance utt with T frames
  int64_t off; int32_t L;
  int64_t p0;                 // unpadded sample position of span[0] (may be negative)
};

// locate the tile: binary search of blockIdx in tile_off
HD bool stft_locate(const StftParams& p, int tile, StftTile* t) {
  int lo = 0, hi = p.n_utts;
  if (tile >= p.tile_off[p.n_utts]) return false;
  if (p.tile_map) {
    lo = p.tile_map[tile];
  } else {
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.tile_off[mid] <= tile) lo = mid; else hi = mid;
    }
  }
  t->utt = lo;
  t->T = (int32_t)(p.frame_off[lo + 1] - p.frame_off[lo]);
  t->t0 = (tile - p.tile_off[lo]) * p.FT;
  t->nf = t->T - t->t0 < p.FT ? t->T - t->t0 : p.FT;
  t->off = p.sample_off[lo];
  t->L = p.sample_len[lo];
  t->p0 = (int64_t)t->t0 * p.hop - p.pad_left;
  return true;
}

HD int64_t stft_reflect(int64_t q, int64_t L) {
  // reflect (no edge repeat) into [0, L); -1 when out of reach (L <= pad, rejected by the host)
  if (q < 0) q = -q;
  if (q >= L) q = 2 * (L - 1) - q;
  return (q < 0 || q >= L) ? -1 : q;
}

HD float stft_sample(const StftParams& p, const StftTile& t, int64_t pos) {
  const int64_t q = stft_reflect(pos, t.L);
  if (q < 0) return 0.f;
  float v = p.audio[t.off + q];
  if (p.has_preemph) {
    if (p.preemph_after_pad) {
      // FilterbankFeatures with exact_pad pads first and pre-emphasises the PADDED signal
      // (features.py:387-400): the previous sample is the previous padded sample
      if (pos > -(int64_t)p.pad_left) {
        const int64_t q1 = stft_reflect(pos - 1, t.L);
        if (q1 >= 0) v = v - p.preemph * p.audio[t.off + q1];
      }
    } else if (q >= 1) {
      // center=True: pre-emphasis on the raw signal, torch.stft reflects the result
      v = v - p.preemph * p.audio[t.off + q - 1];
    }
  }
  return v;
}

// ---- phase A: tables + audio tile (generic path; the TMA path replaces the audio part on device)
HD void stft_phase_tables(const StftParams& p, StftSmem& s, int tid, int nthr) {
  for (int i = tid; i < p.n_fft; i += nthr) s.window[i] = p.window[i];
  {
    const int nt = p.tw_total ? p.tw_total : fft_tw_total(p.M);
    for (int i = tid; i < nt; i += nthr) s.tw[i] = p.tw[i];
  }
  for (int i = tid; i < p.n_mels; i += nthr) {
    s.mel_start[i] = p.mel_start[i]; s.mel_count[i] = p.mel_count[i]; s.mel_offset[i] = p.mel_offset[i];
  }
  for (int i = tid; i < p.mel_nw; i += nthr) s.mel_w[i] = p.mel_w[i];
}

HD void stft_phase_audio(const StftParams& p, const StftTile& t, StftSmem& s, int tid, int nthr,
                         int lo, int hi) {
  // fill span indices [lo, hi)
  for (int i = lo + tid; i < hi; i += nthr) s.audio[i] = stft_sample(p, t, t.p0 + i);
}

// ---- phase B1: window + first radix pass straight from the audio tile
template <int R>
HD void stft_first_pass(const StftParams& p, const StftTile& t, StftSmem& s, int g, int tid) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const float* a = s.audio + f * p.hop;
  cf32* out = s.bufA + (size_t)slot * pidx(p.M);
  const int nb = p.M / R;
  for (int j = u; j < nb; j += p.P) {
    cf32 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int n = j + r * nb;
      v[r].x = a[2 * n] * s.window[2 * n];
      v[r].y = a[2 * n + 1] * s.window[2 * n + 1];
    }
    dftR<R, false>(v);
    stockham_store<R>(v, out, 1, j);
  }
}

template <int R>
HD void stft_pass(const StftParams& p, const StftTile& t, StftSmem& s, int g, int tid, int Ns,
                  const cf32* twp, const cf32* in_base, cf32* out_base) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const cf32* in = in_base + (size_t)slot * pidx(p.M);
  cf32* out = out_base + (size_t)slot * pidx(p.M);
  const int nb = p.M / R;
  for (int j = u; j < nb; j += p.P) {
    cf32 v[R];
    stockham_load<R>(v, in, p.M, j);
    stockham_twiddle_dft<R, false, cf32, false>(v, Ns, j, twp);   // twiddles staged in shared memory
    stockham_store<R>(v, out, Ns, j);
  }
}

// magnitude root: the reference's float32 spectrum is only reproducible to ~1e-6 relative anyway (FFT
// rounding), so the device uses the 2-ulp hardware square root instead of the IEEE sequence
HD float stft_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}

HD float stft_pow(float mag, float power) {
  if (power == 1.0f) return mag;
  if (power == 2.0f) return mag * mag;
  return powf(mag, power);
}

// ---- untangle the real FFT, magnitude (+ floor), power, energy partials.
// Bins k and M-k come from the same pair (Z[k], Z[M-k]):  X[k] = e + t,  X[M-k] = conj(e - t)  with
// e = (Z[k] + conj Z[M-k]) / 2,  t = W_N^k (Z[k] - conj Z[M-k]) / (2i)   (W_N^{M-k} = -conj W_N^k),
// so one thread does both: half the shared-memory reads and one complex product per two bins.
// Returns the thread's share of sum |X|^2 (also parked in s.part[tid] when there is one partial per thread).
HD float stft_phase_post(const StftParams& p, const StftTile& t, StftSmem& s, int g, int tid,
                         const cf32* z_base, float* spec_base) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  float esum = 0.f;
  if (f < t.nf) {
    const cf32* Z = z_base + (size_t)slot * pidx(p.M);
    float* spec = spec_base + (size_t)slot * (2 * pidx(p.M));
    const int half = p.M >> 1;
    for (int k = u; k <= half; k += p.P) {
      const cf32 zk = Z[pidx(k)];
      const cf32 zc = cconj(Z[pidx((p.M - k) & (p.M - 1))]);
      cf32 e; e.x = 0.5f * (zk.x + zc.x); e.y = 0.5f * (zk.y + zc.y);
      cf32 d; d.x = 0.5f * (zk.y - zc.y); d.y = -0.5f * (zk.x - zc.x);   // (zk-zc)/(2i)
      const cf32 w = ld_ro(p.tw_post + k);
      const float tx = w.x * d.x - w.y * d.y, ty = w.x * d.y + w.y * d.x;
      const float re = e.x + tx, im = e.y + ty;
      const float mag = stft_sqrt(re * re + im * im + p.floor_);
      esum += mag * mag;
      spec[k] = stft_pow(mag, p.mag_power);
      if (k != p.M - k) {
        const float re2 = e.x - tx, im2 = e.y - ty;
        const float mag2 = stft_sqrt(re2 * re2 + im2 * im2 + p.floor_);
        esum += mag2 * mag2;
        spec[p.M - k] = stft_pow(mag2, p.mag_power);
      }
    }
  }
  if (p.part_per_slot == 0) s.part[tid] = esum;
  return esum;
}

// ---- sparse mel rows + log guard (+ the energy reduction as extra work items).
// One work item = one mel band for a PAIR of frame slots: the band's weights are read once for both frames,
// and G/2 * n_mels items fit the CTA in one round (the per-slot mapping needed two, the second a quarter full).
HD float stft_mel_log(const StftParams& p, float acc) {
  if (p.log_mode == ROAR_LOG_CLAMP) return logf(acc < p.log_guard ? p.log_guard : acc);
  if (p.log_mode == ROAR_LOG_ADD) return logf(acc + p.log_guard);
  return acc;
}
HD void stft_phase_mel(const StftParams& p, const StftTile& t, StftSmem& s, int g, int tid, int nthr,
                       const float* spec_base) {
  const int pairs = (p.G + 1) >> 1;
  const int n_items = pairs * p.n_mels;
  const int pps = p.part_per_slot ? p.part_per_slot : p.P;
  const size_t sstride = (size_t)(2 * pidx(p.M));
  for (int w = tid; w < n_items + p.G; w += nthr) {
    if (w >= n_items) {                      // energy of one frame slot
      const int slot = w - n_items;
      const int f = g * p.G + slot;
      if (f >= t.nf) continue;
      float e = 0.f;
      const float* pp = s.part + slot * pps;
      for (int i = 0; i < pps; ++i) e += pp[i];
      s.en[f] = sqrtf(e);
      continue;
    }
    const int sp = w / p.n_mels, m = w - sp * p.n_mels;
    const int slot0 = 2 * sp, slot1 = slot0 + 1;
    const int f0 = g * p.G + slot0, f1 = f0 + 1;
    if (f0 >= t.nf) continue;
    const bool two = slot1 < p.G && f1 < t.nf;
    const float* sa = spec_base + (size_t)slot0 * sstride + s.mel_start[m];
    const float* sb = two ? sa + sstride : sa;
    const float* wgt = s.mel_w + s.mel_offset[m];
    const int cnt = s.mel_count[m];
    float a0 = 0.f, a1 = 0.f;
    int c = 0;
    for (; c + 1 < cnt; c += 2) {
      const float w0 = wgt[c], w1 = wgt[c + 1];
      a0 += w0 * sa[c]; a1 += w0 * sb[c];
      a0 += w1 * sa[c + 1]; a1 += w1 * sb[c + 1];
    }
    if (c < cnt) { const float w0 = wgt[c]; a0 += w0 * sa[c]; a1 += w0 * sb[c]; }
    s.out[m * (p.FT + 1) + f0] = stft_mel_log(p, a0);
    if (two) s.out[m * (p.FT + 1) + f1] = stft_mel_log(p, a1);
  }
}

HD void stft_phase_store(const StftParams& p, const StftTile& t, StftSmem& s, int tid, int nthr) {
  if (p.logmel) {
    const int64_t base = p.out_utt_stride ? (int64_t)t.utt * p.out_utt_stride
                                          : (int64_t)p.n_mels * p.frame_off[t.utt];
    const int64_t rs = p.out_row_stride ? p.out_row_stride : t.T;
    for (int i = tid; i < p.n_mels * p.FT; i += nthr) {
      const int m = i / p.FT, f = i - m * p.FT;
      if (f < t.nf) p.logmel[base + (int64_t)m * rs + t.t0 + f] = s.out[m * (p.FT + 1) + f];
    }
  }
  if (p.energy) {
    for (int f = tid; f < t.nf; f += nthr) {
      float e = s.en[f];
      if (p.energy_mode == 1) {   // EnergyFeaturizer: torch.linalg.norm(features, axis=0)
        float acc = 0.f;
        for (int m = 0; m < p.n_mels; ++m) { const float v = s.out[m * (p.FT + 1) + f]; acc += v * v; }
        e = sqrtf(acc);
      }
      p.energy[p.frame_off[t.utt] + t.t0 + f] = e;
    }
  }
}

// One FFT pass of radix R over all frames in flight (dispatch on the runtime radix).
HD void stft_pass_any(int R, const StftParams& p, const StftTile& t, StftSmem& s, int g, int tid,
                      int Ns, const cf32* twp, const cf32* src, cf32* dst) {
  if (R == 8) stft_pass<8>(p, t, s, g, tid, Ns, twp, src, dst);
  else if (R == 4) stft_pass<4>(p, t, s, g, tid, Ns, twp, src, dst);
  else stft_pass<2>(p, t, s, g, tid, Ns, twp, src, dst);
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// LG = log2(M): the radix plan (8,...,8[,4|2]) is a compile-time constant, so the pass loop unrolls and
// nothing about the plan lives in local memory.
template <int LG>
__global__ void __launch_bounds__(256) k_stft_mel(const StftParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StftSmem s;
  const int nthr = blockDim.x, tid = threadIdx.x;
  stft_smem_carve(p, nthr, smem_raw, &s);
  StftTile t;
  if (!stft_locate(p, blockIdx.x, &t)) return;

  // ---- audio tile: one TMA bulk copy for the 16-B aligned interior, mirrored loads for the rest
  const int64_t need_lo = t.p0, need_hi = t.p0 + (int64_t)(t.nf - 1) * p.hop + p.n_fft;  // [lo, hi)
  int lo_i = 0, hi_i = (int)(need_hi - need_lo);     // span indices to fill
  int tma_lo = 0, tma_hi = 0;                        // span indices covered by the bulk copy
  if (p.use_tma && !p.has_preemph) {
    int64_t a = need_lo < 0 ? 0 : need_lo, b = need_hi > t.L ? t.L : need_hi;
    // 16-B alignment of both the global source and the shared destination
    const int64_t g0 = t.off + a;
    int64_t a4 = a + ((4 - (g0 & 3)) & 3);
    if (((a4 - need_lo) & 3) != 0) a4 = b;           // destination not alignable: skip TMA
    int64_t b4 = a4 + ((b - a4) & ~(int64_t)3);
    if (b4 - a4 >= 64 && ((reinterpret_cast<uintptr_t>(p.audio) & 15) == 0)) {
      tma_lo = (int)(a4 - need_lo); tma_hi = (int)(b4 - need_lo);
    }
  }
  if (tma_hi > tma_lo) {
    const unsigned bar = smem_u32(s.mbar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const unsigned bytes = (unsigned)(tma_hi - tma_lo) * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(s.audio + tma_lo)),
          "l"(p.audio + t.off + (need_lo + tma_lo)), "r"(bytes), "r"(bar)
          : "memory");
    }
    stft_phase_tables(p, s, tid, nthr);
    stft_phase_audio(p, t, s, tid, nthr, lo_i, tma_lo);
    stft_phase_audio(p, t, s, tid, nthr, tma_hi, hi_i);
    // wait for the bulk copy (phase parity 0)
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
          : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
  } else {
    stft_phase_tables(p, s, tid, nthr);
    stft_phase_audio(p, t, s, tid, nthr, lo_i, hi_i);
  }
  __syncthreads();

  const int n_groups = (t.nf + p.G - 1) / p.G;
  for (int g = 0; g < n_groups; ++g) {
    stft_first_pass<8>(p, t, s, g, tid);
    __syncthreads();
    constexpr int NP8 = LG / 3, REM = LG % 3;
    int Ns = 8, off = 0;
    const cf32* src = s.bufA;
    cf32* dst = s.bufB;
#pragma unroll
    for (int ps = 1; ps < NP8; ++ps) {
      stft_pass<8>(p, t, s, g, tid, Ns, s.tw + off, src, dst);
      __syncthreads();
      off += 7 * Ns; Ns *= 8;
      const cf32* tmp = src; src = dst; dst = const_cast<cf32*>(tmp);
    }
    if (REM != 0) {
      if (REM == 2) stft_pass<4>(p, t, s, g, tid, Ns, s.tw + off, src, dst);
      else stft_pass<2>(p, t, s, g, tid, Ns, s.tw + off, src, dst);
      __syncthreads();
      const cf32* tmp = src; src = dst; dst = const_cast<cf32*>(tmp);
    }
    const cf32* zfin = src;            // FFT result
    float* spec = (float*)dst;         // the other buffer holds |X| for the mel phase
    {
      float esum = stft_phase_post(p, t, s, g, tid, zfin, spec);
      if (p.part_per_slot) {          // P >= 32: a warp lies inside one frame slot -> one partial per warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
        if ((tid & 31) == 0) s.part[tid >> 5] = esum;
      }
    }
    __syncthreads();
    stft_phase_mel(p, t, s, g, tid, nthr, spec);
    __syncthreads();
  }
  stft_phase_store(p, t, s, tid, nthr);
}

// prefix sum of ceil(T_i/FT) -> tile_off[n_utts+1]; one CTA
__global__ void k_tile_offsets(const int64_t* frame_off, int32_t n_utts, int32_t FT, int32_t* tile_off) {
  __shared__ int32_t s_carry;
  __shared__ int32_t s_scan[1024];
  const int tid = threadIdx.x;
  if (tid == 0) { s_carry = 0; tile_off[0] = 0; }
  __syncthreads();
  for (int base = 0; base < n_utts; base += 1024) {
    const int i = base + tid;
    int32_t v = 0;
    if (i < n_utts) { int64_t T = frame_off[i + 1] - frame_off[i]; v = (int32_t)((T + FT - 1) / FT); }
    s_scan[tid] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      int32_t x = tid >= d ? s_scan[tid - d] : 0;
      __syncthreads();
      s_scan[tid] += x;
      __syncthreads();
    }
    if (i < n_utts) tile_off[i + 1] = s_carry + s_scan[tid];
    __syncthreads();
    if (tid == 1023) s_carry += s_scan[1023];
    __syncthreads();
  }
}
// tile -> utterance map from the tile prefix sums (one thread per tile, binary search once instead of in
// every CTA of the consumer kernels)
__global__ void k_tile_map(const int32_t* tile_off, int32_t n_utts, int32_t* map) {
  const int n_tiles = tile_off[n_utts];
  for (int tile = blockIdx.x * blockDim.x + threadIdx.x; tile < n_tiles; tile += gridDim.x * blockDim.x) {
    int lo = 0, hi = n_utts;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tile_off[mid] <= tile) lo = mid; else hi = mid;
    }
    map[tile] = lo;
  }
}
#endif  // __CUDACC__

}  // namespace roar
