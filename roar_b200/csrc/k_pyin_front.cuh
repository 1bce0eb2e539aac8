// K2  pYIN front end.
//   K2a  pyin_cmnd : per frame, YIN difference function -> cumulative-mean-normalised difference.
//        The autocorrelation  acf[tau] = sum_{j=1..W} y[j] * y[j+tau]  (what the reference obtains
//        through float64 FFTs) is evaluated DIRECTLY in float64: the products of two float32 samples
//        are exact in float64, so the only rounding is the accumulation.  Consecutive frames overlap
//        by W - hop samples, so the sum is split into hop-sized blocks Q_m[tau] shared by the W/hop
//        frames that contain block m (half the work at the default W = 2*hop).  Each thread owns
//        R = 12 consecutive lags of one block: a register-resident sliding window gives 12 DFMA per
//        two shared-memory loads.  The energy terms replay the reference's float32 sequential cumsum
//        bit for bit (it carries ~1e-4 relative noise into the trough heights, so replaying it is
//        what makes the thresholds agree).
//   K2b  pyin_probs: per frame (one warp), parabolic shifts, troughs, threshold-beta / Boltzmann
//        probabilities, pitch-bin quantisation -> sparse observation list + voiced probability.
//
// Replaces librosa.pyin's `_cumulative_mean_normalized_difference`, `_parabolic_interpolation`,
// `__pyin_helper` as called from roar/collections/tts/data/dataset.py:696-703.
#pragma once
#include "common.cuh"

namespace roar {

constexpr int ACF_R = 11;   // lags per thread (odd: a lane stride of 11 doubles is bank-conflict-free, 31 lanes cover 341 lags)

struct PyinParams {
  const float* audio;
  const int64_t* sample_off;
  const int32_t* sample_len;
  const int64_t* frame_off;   // [n_utts+1] pyin frames
  const int32_t* tile_off;    // [n_utts+1]
  const int32_t* tile_map;    // [n_tiles] tile -> utterance (optional: null = binary search of tile_off)
  int32_t n_utts;
  // geometry
  int32_t F, W, hop;          // frame, win, hop
  int32_t min_period, max_period, n_lags;
  int32_t FT;                 // frames per tile
  int32_t BL, nb;             // autocorrelation block length, blocks per frame (W = nb*BL when shared)
  int32_t n_groups;           // ceil((max_period+1) / ACF_R) lag groups
  int32_t ylen;               // staged samples per tile (incl. zero tail read by the padded lag groups)
  int32_t npb, nbps, kmax, n_thr;
  double sr, fmin, no_trough_prob;
  // tables
  const double* thresholds;   // [n_thr+1]
  const double* beta_probs;   // [n_thr]
  const double* beta_cum;     // [n_thr+1]
  const double* boltz_exp;    // [kmax+1]
  const double* boltz_fact;   // [kmax+1]
  // scratch / outputs
  float* energy;              // [max_period+1, total_frames]  E[tau] = cs[W+tau] - cs[tau] (float32 scratch, frame-contiguous)
  const int32_t* etile_off;   // [n_utts+1] prefix sum of ceil(T_i / 32): energy-kernel tiles
  const int32_t* etile_map;   // [n_etiles] energy tile -> utterance (optional: null = binary search of etile_off)
  double* cmnd;               // [total_frames, n_lags]
  uint16_t* cand_bin;         // [total_frames, kmax]
  double* cand_lp;            // [total_frames, kmax]
  int32_t* n_cand;            // [total_frames]
  double* lp_unvoiced;        // [total_frames]
  float* voiced_prob;         // [total_frames]  (final output)
  int64_t total_frames;
};

// autocorrelation blocking: blocks of `hop` samples shared between frames when W is a multiple of
// hop, one private block of W samples per frame otherwise
HD void cmnd_blocking(int W, int hop, int* BL, int* nb) {
  if (W % hop == 0) { *BL = hop; *nb = W / hop; } else { *BL = W; *nb = 1; }
}
HD int cmnd_n_blocks(const PyinParams& p, int nf) { return nf + p.nb - 1; }
// samples of the tile the kernel touches: frames, plus the zero tail the padded lag groups read
HD int cmnd_ylen(int FT, int F, int hop, int BL, int nb, int n_groups) {
  const int a = (FT - 1) * hop + F;
  const int b = (FT + nb - 2) * hop + BL + n_groups * ACF_R + 1;
  return ((a > b ? a : b) + 3) & ~3;
}

struct CmndSmem {
  double* yd;      // [ylen]   tile samples as float64 (exact), zero outside the utterance
  double* Q;       // [FT+nb-1][QS]  block partial autocorrelations; then d[f][tau] in place
  float* E;        // [max_period+1][FT]  the tile's energy terms (from K2a-0's scratch)
  // aliases of yd, valid once the autocorrelation is done
  double* dsum;    // [n_slots][max_period+1]
  double* chunk;   // [n_slots][64]
};

HD size_t cmnd_align16(size_t x) { return (x + 15) & ~(size_t)15; }
HD int cmnd_qs(const PyinParams& p) { return p.n_groups * ACF_R; }
constexpr int CMND_SLOTS = 8;   // frames scanned concurrently (one warp each)

HD size_t cmnd_smem_carve(const PyinParams& p, unsigned char* base, CmndSmem* s) {
  size_t o = 0;
#define CARVE(field, type, count) { if (s) s->field = (type*)(base + o); o = cmnd_align16(o + sizeof(type) * (size_t)(count)); }
  size_t ny = (size_t)p.ylen + 4;
  const size_t alias = (size_t)CMND_SLOTS * (p.max_period + 1) + (size_t)CMND_SLOTS * 64;
  if (ny < alias) ny = alias;
  CARVE(yd, double, ny)
  CARVE(Q, double, (size_t)(p.FT + p.nb - 1) * cmnd_qs(p))
  CARVE(E, float, (size_t)p.FT * (p.max_period + 1))
#undef CARVE
  if (s) { s->dsum = s->yd; s->chunk = s->yd + (size_t)CMND_SLOTS * (p.max_period + 1); }
  return o;
}

struct PyinTile {
  int32_t utt, t0, T, nf;
  int64_t off; int32_t L;
  int64_t p0;
};

HD bool pyin_locate(const PyinParams& p, int tile, PyinTile* t) {
  int lo = 0, hi = p.n_utts;
  if (tile >= p.tile_off[p.n_utts]) return false;
  if (p.tile_map) {
    lo = p.tile_map[tile];      // one load instead of log2(n_utts) dependent ones
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
  t->p0 = (int64_t)t->t0 * p.hop - p.F / 2;   // librosa center=True, pad_mode="constant"
  return true;
}

constexpr int CMND_LOAD_U = 6;   // sample groups (of 4) a thread keeps in flight
HD void cmnd_phase_load(const PyinParams& p, const PyinTile& t, CmndSmem& s, int tid, int nthr) {
  // energy terms first: asynchronous 4-byte copies on the device, consumed after the autocorrelation
  // (consecutive threads take consecutive frames of one lag: the global reads stay sector-coalesced)
  const float* Eg = p.energy + (p.frame_off[t.utt] + t.t0);
  {   // LW lanes per lag (LW >= FT, a power of two): frame = tid & (LW-1), no division per element
    const int lw_log = p.FT <= 16 ? 4 : 5, LW = 1 << lw_log;
    const int f = tid & (LW - 1);
    if (f < t.nf) {
      for (int tau = tid >> lw_log; tau <= p.max_period; tau += nthr >> lw_log) {
        const float* src = Eg + (size_t)tau * p.total_frames + f;
        float* dst = s.E + tau * p.FT + f;
#if defined(__CUDA_ARCH__)
        const unsigned d32 = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d32), "l"(src) : "memory");
#else
        *dst = *src;
#endif
      }
    }
  }
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
  // samples in groups of 4: a 16-byte load when the group lies inside the utterance and the source is
  // aligned, scalar loads at the edges
  const float* src = p.audio + t.off;
  // (src + p0 + i) is 16-byte aligned for every i % 4 == 0 iff (address / 4 + p0) % 4 == 0
  const bool aligned = (((reinterpret_cast<uintptr_t>(src) >> 2) + (uintptr_t)(t.p0 & 3)) & 3) == 0;
  const int n4 = p.ylen >> 2;      // ylen is a multiple of 4
  // in batches of CMND_LOAD_U groups per thread: all global loads of a batch are issued before the first
  // conversion waits on one (one exposed memory latency per batch instead of one per group)
  for (int g0 = tid; g0 < n4; g0 += CMND_LOAD_U * nthr) {
    float v[CMND_LOAD_U][4];
#pragma unroll
    for (int u = 0; u < CMND_LOAD_U; ++u) {
      const int g4 = g0 + u * nthr;
      const int64_t q = t.p0 + ((int64_t)g4 << 2);
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (g4 < n4) {
        if (aligned && q >= 0 && q + 3 < t.L) {
#if defined(__CUDA_ARCH__)
          const float4 x = __ldg(reinterpret_cast<const float4*>(src + q));
          v[u][0] = x.x; v[u][1] = x.y; v[u][2] = x.z; v[u][3] = x.w;
#else
          v[u][0] = src[q]; v[u][1] = src[q + 1]; v[u][2] = src[q + 2]; v[u][3] = src[q + 3];
#endif
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (q + e >= 0 && q + e < t.L) v[u][e] = src[q + e];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < CMND_LOAD_U; ++u) {
      const int g4 = g0 + u * nthr;
      if (g4 < n4) {
        double* dst = s.yd + (g4 << 2);
        dst[0] = (double)v[u][0]; dst[1] = (double)v[u][1]; dst[2] = (double)v[u][2]; dst[3] = (double)v[u][3];
      }
    }
  }
}

// K2a-0  float32 sequential cumsum of y^2 (np.cumsum(y_frames**2, axis=-2) on float32 frames): an
// inherently serial chain per frame, so it runs as its own small kernel -- one warp per tile of 32
// consecutive frames, one lane per frame, the tile's squared samples staged in shared memory -- instead
// of stalling a tile of K2a.  E[tau] = cs[W+tau] - cs[tau] (|.| < 1e-6 flushed to 0 like the reference):
// cs[tau] is not parked anywhere, a second accumulator replays the same additions tau = 0..max_period
// in lockstep (bit-identical by construction), so the chain never waits on memory it wrote itself.
HD float f32_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
HD float f32_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
HD float f32_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
constexpr int ENERGY_FT = 32;   // frames per energy tile (one warp)
HD int epad(int i, int hop) { return i + i / hop; }   // frame stride hop+1 floats: conflict-free lanes
HD int energy_span(const PyinParams& p) { return (ENERGY_FT - 1) * p.hop + p.W + p.max_period + 1; }
// stage the tile's squared samples (zero outside the utterance; librosa center=True, pad_mode="constant")
HD void pyin_energy_stage(const PyinParams& p, const float* y, int L, int64_t q0, int n, float* sq, int tid,
                          int nthr) {
  const int hop = p.hop;
#if defined(__CUDA_ARCH__)
  // groups of 4 samples: one 16-byte load when the group lies inside the utterance, inside one hop-sized
  // segment of the padded layout and the source is aligned; scalar loads at the edges
  if ((hop & 3) == 0) {
    const bool aligned = (((reinterpret_cast<uintptr_t>(y) >> 2) + (uintptr_t)(q0 & 3)) & 3) == 0;
    const int n4 = (n + 3) >> 2;
    for (int g = tid; g < n4; g += nthr) {
      const int i0 = g << 2;                       // sample index in the tile; hop % 4 == 0: one segment
      const int sg = i0 / hop;
      float* dst = sq + (size_t)sg * (hop + 1) + (i0 - sg * hop);
      const int64_t q = q0 + i0;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      if (aligned && q >= 0 && q + 3 < L) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(y + q));
        v0 = x.x; v1 = x.y; v2 = x.z; v3 = x.w;
      } else {
        if (q >= 0 && q < L) v0 = y[q];
        if (q + 1 >= 0 && q + 1 < L) v1 = y[q + 1];
        if (q + 2 >= 0 && q + 2 < L) v2 = y[q + 2];
        if (q + 3 >= 0 && q + 3 < L) v3 = y[q + 3];
      }
      dst[0] = f32_mul(v0, v0);
      if (i0 + 1 < n) dst[1] = f32_mul(v1, v1);
      if (i0 + 2 < n) dst[2] = f32_mul(v2, v2);
      if (i0 + 3 < n) dst[3] = f32_mul(v3, v3);
    }
    return;
  }
#endif
  for (int s0 = 0, sg = 0; s0 < n; s0 += hop, ++sg) {
    const int len = n - s0 < hop ? n - s0 : hop;
    float* dst = sq + (size_t)sg * (hop + 1);
    for (int i = tid; i < len; i += nthr) {
      const int64_t q = q0 + s0 + i;
      const float v = (q >= 0 && q < L) ? y[q] : 0.f;
      dst[i] = f32_mul(v, v);
    }
  }
}
// one frame's chain; E = &energy[frame], rows `stride` floats apart (frame-contiguous rows).
// The epad layout is contiguous inside a hop-sized segment, so the chain is walked run by run:
// plain unit-stride inner loops, no per-sample index arithmetic.
HD void pyin_energy_frame(const PyinParams& p, const float* sq, int f, float* E, size_t stride) {
  const int hop = p.hop, seg = hop + 1;
  float cs = 0.f, cs2 = 0.f;
  // phase A: n = 0 .. W-1
  for (int n0 = 0; n0 < p.W; n0 += hop) {
    const float* ph = sq + (size_t)(f + n0 / hop) * seg;
    const int len = p.W - n0 < hop ? p.W - n0 : hop;
    for (int i = 0; i < len; ++i) cs = f32_add(cs, ph[i]);
  }
  // phase C: tau = 0 .. max_period; hi stream at sample W + tau, lo stream at sample tau
  int tau = 0;
  while (tau <= p.max_period) {
    const int nh = p.W + tau, nl = tau;
    const float* ph = sq + (size_t)(f + nh / hop) * seg + nh % hop;
    const float* pl = sq + (size_t)(f + nl / hop) * seg + nl % hop;
    int len = p.max_period + 1 - tau;
    if (hop - nh % hop < len) len = hop - nh % hop;
    if (hop - nl % hop < len) len = hop - nl % hop;
    for (int i = 0; i < len; ++i) {
      cs = f32_add(cs, ph[i]);
      cs2 = f32_add(cs2, pl[i]);
      float e = f32_sub(cs, cs2);
      if (fabsf(e) < 1e-6f) e = 0.f;
      *E = e;
      E += stride;
    }
    tau += len;
  }
}

// Autocorrelation work is handed out in warp-sized chunks: chunk c = (block m, part) where a block's
// lag groups are split into `cpb` parts of at most 32 lanes, so all lanes of a warp share the block
// (one broadcast load for y[j]) and read windows 15 doubles apart (conflict-free).
HD int cmnd_cpb(const PyinParams& p) { return (p.n_groups + 31) / 32; }
HD int cmnd_lpc(const PyinParams& p) { const int c = cmnd_cpb(p); return (p.n_groups + c - 1) / c; }

// One autocorrelation unit: block m, lag group g ->  Q[m][tau] = sum_{j=1..BL} y[m*hop+j] * y[m*hop+j+tau]
// for tau = g*R .. g*R+R-1.  The window y[j+tau0 .. j+tau0+R-1] lives in registers and slides by one
// sample per step (static rotation: the loop is unrolled by R; every shared-memory offset inside the
// unrolled body is a compile-time constant).
HD void cmnd_acf_unit(const PyinParams& p, CmndSmem& s, int m, int g) {
  constexpr int R = ACF_R;
  const int b0 = m * p.hop;                // j = 1..BL -> samples b0+1 .. b0+BL
  const double* ya = s.yd + b0;
  const double* yw = s.yd + b0 + g * R;
  double acc[R], win[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
#pragma unroll
  for (int r = 0; r < R - 1; ++r) win[r] = yw[r + 1];
  const int full = p.BL / R * R;
  int jj = 0;
  for (; jj < full; jj += R) {
    const double* a_ = ya + jj;
    const double* w_ = yw + jj;
#pragma unroll
    for (int st = 0; st < R; ++st) {
      const double a = a_[st + 1];
      win[(st + R - 1) % R] = w_[st + R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fma(a, win[(st + r) % R], acc[r]);
    }
  }
  {   // tail (BL not a multiple of R): same body, guarded
    const double* a_ = ya + jj;
    const double* w_ = yw + jj;
#pragma unroll
    for (int st = 0; st < R; ++st) {
      if (jj + st < p.BL) {
        const double a = a_[st + 1];
        win[(st + R - 1) % R] = w_[st + R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fma(a, win[(st + r) % R], acc[r]);
      }
    }
  }
  double* q = s.Q + (size_t)m * cmnd_qs(p) + g * R;
#pragma unroll
  for (int r = 0; r < R; ++r) q[r] = acc[r];
}

// difference function d[tau] = (E[0] + E[tau]) - 2*acf[tau], acf = sum of the frame's blocks, written
// in place over Q[f]; one thread per lag walks the frames in order (Q[f+1..] is still intact).
HD void cmnd_phase_diff(const PyinParams& p, const PyinTile& t, CmndSmem& s, int tid, int nthr) {
  const int qs = cmnd_qs(p);
  for (int tau = tid; tau <= p.max_period; tau += nthr) {
    for (int f = 0; f < t.nf; ++f) {
      double acf = s.Q[(size_t)f * qs + tau];
      for (int b = 1; b < p.nb; ++b) acf += s.Q[(size_t)(f + b) * qs + tau];
      if (fabs(acf) < 1e-6) acf = 0.0;
      const float e2 = f32_add(s.E[f], s.E[tau * p.FT + f]);
      s.Q[(size_t)f * qs + tau] = (double)e2 - 2.0 * acf;
    }
  }
}

// cumulative sum of d[1..max_period] per frame: chunk-local prefix, chunk offsets, combine.
// One 32-lane slot per frame.
HD int cmnd_chunk(const PyinParams& p) { return (p.max_period + 31) / 32; }

HD void cmnd_phase_scan1(const PyinParams& p, CmndSmem& s, int f, int slot, int u) {
  const int ch = cmnd_chunk(p);
  const double* d = s.Q + (size_t)f * cmnd_qs(p);
  double* ds = s.dsum + (size_t)slot * (p.max_period + 1);
  double acc = 0.0;
  for (int i = 0; i < ch; ++i) {
    const int tau = 1 + u * ch + i;
    if (tau > p.max_period) break;
    acc += d[tau];
    ds[tau] = acc;
  }
  s.chunk[slot * 64 + u] = acc;
}
HD void cmnd_phase_scan2(const PyinParams& p, CmndSmem& s, int slot, int u) {
  double acc = 0.0;
  for (int v = 0; v < u; ++v) acc += s.chunk[slot * 64 + v];
  s.chunk[slot * 64 + 32 + u] = acc;
}
HD void cmnd_phase_emit(const PyinParams& p, const PyinTile& t, CmndSmem& s, int f, int slot, int u) {
  const int ch = cmnd_chunk(p);
  const double* d = s.Q + (size_t)f * cmnd_qs(p);
  const double* ds = s.dsum + (size_t)slot * (p.max_period + 1);
  double* out = p.cmnd + (size_t)(p.frame_off[t.utt] + t.t0 + f) * p.n_lags;
  // (tau - 1) / ch by a multiply-shift: exact for tau - 1 < 2^20 / ch (tau <= max_period ~ 1.4 k, ch >= 1)
  const unsigned chm = ((1u << 20) + (unsigned)ch - 1u) / (unsigned)ch;
  for (int i = u; i < p.n_lags; i += 32) {
    const int tau = p.min_period + i;
    const double c = ds[tau] + s.chunk[slot * 64 + 32 + (int)(((unsigned)(tau - 1) * chm) >> 20)];
    out[i] = d[tau] / (c / (double)tau + 2.2250738585072014e-308);
  }
}

// =================================================================================== K2b
struct ProbSmem {
  double* x;          // [n_lags]
  double* prob;       // [kmax]
  uint16_t* tr;       // [kmax] trough lag indices (increasing)
  uint16_t* sorted;   // [kmax] active troughs (c_r < n_thr) in lag order
  int16_t* bin;       // [kmax]
  uint8_t* cr;        // [kmax] first threshold index the trough is below (n_thr = never)
  int32_t* cnt;       // [36]
  double* red;        // [32]
  uint8_t* live;      // [kmax]
  int32_t* hist;      // [2 * n_thr] per-threshold counts of active troughs: all / already-processed chunks
};
HD size_t prob_smem_carve(const PyinParams& p, unsigned char* base, ProbSmem* s) {
  size_t o = 0;
#define CARVE(field, type, count) { if (s) s->field = (type*)(base + o); o = cmnd_align16(o + sizeof(type) * (size_t)(count)); }
  CARVE(x, double, p.n_lags + 2)
  CARVE(prob, double, p.kmax)
  CARVE(red, double, 32)
  CARVE(tr, uint16_t, p.kmax)
  CARVE(sorted, uint16_t, p.kmax + 2)
  CARVE(live, uint8_t, p.kmax)
  CARVE(hist, int32_t, 2 * p.n_thr)
  CARVE(bin, int16_t, p.kmax)
  CARVE(cr, uint8_t, p.kmax)
  CARVE(cnt, int32_t, 36)
#undef CARVE
  return o;
}

HD bool prob_is_trough(const double* x, int i, int n) {
  if (i == 0) return x[0] < x[1];
  if (i == n - 1) return x[n - 1] < x[n - 2];
  return (x[i] < x[i - 1]) && (x[i] <= x[i + 1]);
}

// lane-phased per-frame pipeline; `lane` in [0, 32); WSYNC between phases
HD void prob_phase0(const PyinParams& p, ProbSmem& s, int64_t frame, int lane) {
  const double* src = p.cmnd + (size_t)frame * p.n_lags;
  for (int i = lane; i < p.n_lags; i += 32) s.x[i] = src[i];
}
HD void prob_phase1(const PyinParams& p, ProbSmem& s, int lane) {
  const int ch = (p.n_lags + 31) / 32;
  int c = 0;
  for (int i = lane * ch; i < (lane + 1) * ch && i < p.n_lags; ++i) c += prob_is_trough(s.x, i, p.n_lags) ? 1 : 0;
  s.cnt[lane] = c;
}
HD void prob_phase2(const PyinParams& p, ProbSmem& s, int lane) {
  const int ch = (p.n_lags + 31) / 32;
  int off = 0;
  for (int v = 0; v < lane; ++v) off += s.cnt[v];
  for (int i = lane * ch; i < (lane + 1) * ch && i < p.n_lags; ++i)
    if (prob_is_trough(s.x, i, p.n_lags)) s.tr[off++] = (uint16_t)i;
  if (lane == 31) s.cnt[32] = off;   // R
}
HD void prob_phase3(const PyinParams& p, ProbSmem& s, int lane, const double* thr) {
  const int R = s.cnt[32];
  double best = 1e300; int bi = 0x7fffffff;
  for (int r = lane; r < R; r += 32) {
    const double h = s.x[s.tr[r]];
    // c_r = #{c in [1, n_thr] : thr[c] <= h}  (thresholds ascending)
    int lo = 0, hi = p.n_thr;   // invariant: thr[1..lo] <= h, thr[hi+1..] > h
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (thr[mid] <= h) lo = mid; else hi = mid - 1;
    }
    s.cr[r] = (uint8_t)lo;
    if (h < best) { best = h; bi = r; }
  }
  s.red[lane] = best;
  s.cnt[lane] = bi;   // cnt[0..31] reused: per-lane argmin
}
// global minimum trough (first index of the minimum height) from the per-lane partials
HD void prob_phase4(const PyinParams& p, ProbSmem& s, int lane) {
  if (lane != 0) return;
  double best = 1e300; int bi = 0x7fffffff;
  for (int v = 0; v < 32; ++v) {
    const double h = s.red[v]; const int i = s.cnt[v];
    if (i != 0x7fffffff && (h < best || (h == best && i < bi))) { best = h; bi = i; }
  }
  s.cnt[33] = bi;
}

// Threshold-beta / Boltzmann probability of trough r is finished here: the no-trough bonus on the
// global minimum, parabolic period refinement and pitch-bin quantisation.
HD void prob_trough_finish(const PyinParams& p, ProbSmem& s, int r, double acc) {
  const int cr = s.cr[r];
  if (r == s.cnt[33]) acc += p.no_trough_prob * p.beta_cum[cr];
  s.prob[r] = acc;
  int bin = -1;
  if (acc != 0.0) {
    const int i = s.tr[r];
    double shift = 0.0;
    if (i > 0 && i < p.n_lags - 1) {
      const double a = s.x[i + 1] + s.x[i - 1] - 2.0 * s.x[i];
      const double b = (s.x[i + 1] - s.x[i - 1]) / 2.0;
      if (!(fabs(b) >= fabs(a))) shift = -b / a;
    }
    const double period = (double)(p.min_period + i) + shift;
    const double f0 = p.sr / period;
    const double bf = (double)(12 * p.nbps) * log2(f0 / p.fmin);
    double rb = rint(bf);
    if (rb < 0.0) rb = 0.0;
    if (rb > (double)p.npb) rb = (double)p.npb;
    bin = (int)rb;
  }
  s.bin[r] = (int16_t)bin;
}

// Only ACTIVE troughs (height below the last threshold, c_r < n_thr) take part in the threshold sums:
// an inactive trough is never "below", so it adds nothing to n_c or to anybody's rank.  The active
// troughs are compacted in lag order (s.sorted[0..Na)); for trough a
//   probs_a = sum_{c = c_a .. n_thr-1} (fact[n_c] * exp(-lambda * pos(a, c))) * beta[c],
//   n_c = #{a' : c_a' <= c},  pos(a, c) = #{a' < a : c_a' <= c},
// accumulated in ascending c exactly like the dense formulation.  This is the sequential statement the
// host harness runs; the kernel evaluates the same sums with ballots (lanes own active troughs).
HD int prob_compact_active(const PyinParams& p, ProbSmem& s) {
  const int R = s.cnt[32];
  int na = 0;
  for (int r = 0; r < R; ++r) if ((int)s.cr[r] < p.n_thr) s.sorted[na++] = (uint16_t)r;
  return na;
}
HD double prob_active_sum(const PyinParams& p, const ProbSmem& s, int na, int a) {
  const int cra = s.cr[s.sorted[a]];
  double acc = 0.0;
  for (int c = cra; c < p.n_thr; ++c) {
    int n = 0, pos = 0;
    for (int q = 0; q < na; ++q) {
      const bool below = (int)s.cr[s.sorted[q]] <= c;
      n += below ? 1 : 0;
      pos += (below && q < a) ? 1 : 0;
    }
    acc += (p.boltz_fact[n] * p.boltz_exp[pos]) * p.beta_probs[c];
  }
  return acc;
}

// live = survives NumPy's last-write-wins scatter: no later trough with a non-zero probability
// lands on the same pitch bin (bins are non-increasing in lag order, so only the run of following
// zero-probability troughs has to be skipped)
HD void prob_phase6a(const PyinParams& p, ProbSmem& s, int lane) {
  const int R = s.cnt[32];
  for (int r = lane; r < R; r += 32) {
    const int b = s.bin[r];
    bool live = b >= 0 && b < p.npb;
    if (live) {
      for (int q = r + 1; q < R; ++q) {
        const int bq = s.bin[q];
        if (bq < 0) continue;
        live = bq != b;
        if (live) {
          // bins are monotone, but stay exact even if they were not: finish the scan
          for (int q2 = q + 1; q2 < R; ++q2) if (s.bin[q2] == b) { live = false; break; }
        }
        break;
      }
    }
    s.live[r] = live ? 1 : 0;
  }
}
// lane 0: voiced probability (ascending pitch bin = descending lag, like np.sum over the bins) and
// the compact observation list
HD void prob_phase6b(const PyinParams& p, ProbSmem& s, int64_t frame, int lane) {
  if (lane != 0) return;
  const int R = s.cnt[32];
  uint16_t* ob = p.cand_bin + (size_t)frame * p.kmax;
  double* ol = p.cand_lp + (size_t)frame * p.kmax;
  int n = 0;
  double vp = 0.0;
  for (int r = R - 1; r >= 0; --r) {
    if (!s.live[r]) continue;
    vp += s.prob[r];
    ob[n] = (uint16_t)s.bin[r];
    ol[n] = log(s.prob[r] + 2.2250738585072014e-308);
    ++n;
  }
  if (vp < 0.0) vp = 0.0;
  if (vp > 1.0) vp = 1.0;
  p.n_cand[frame] = n;
  p.lp_unvoiced[frame] = log((1.0 - vp) / (double)p.npb + 2.2250738585072014e-308);
  p.voiced_prob[frame] = (float)vp;
}

#ifdef __CUDACC__
// one CTA per tile of 32 consecutive frames of one utterance: 128 threads stage the squared samples,
// then warp 0 runs the 32 serial chains (one lane per frame)
constexpr int ENERGY_THREADS = 128;
__global__ void __launch_bounds__(ENERGY_THREADS) k_pyin_energy(const PyinParams p) {
  extern __shared__ __align__(16) float e_smem[];
  const int tile = blockIdx.x, tid = threadIdx.x;
  if (tile >= p.etile_off[p.n_utts]) return;
  int lo = 0, hi = p.n_utts;
  if (p.etile_map) {
    lo = p.etile_map[tile];
  } else {
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.etile_off[mid] <= tile) lo = mid; else hi = mid;
    }
  }
  const int T = (int)(p.frame_off[lo + 1] - p.frame_off[lo]);
  const int t0 = (tile - p.etile_off[lo]) * ENERGY_FT;
  const int nf = T - t0 < ENERGY_FT ? T - t0 : ENERGY_FT;
  const int n = (nf - 1) * p.hop + p.W + p.max_period + 1;
  pyin_energy_stage(p, p.audio + p.sample_off[lo], p.sample_len[lo], (int64_t)t0 * p.hop - p.F / 2, n, e_smem,
                    tid, ENERGY_THREADS);
  __syncthreads();
  if (tid < nf)
    pyin_energy_frame(p, e_smem, tid, p.energy + (p.frame_off[lo] + t0 + tid), (size_t)p.total_frames);
}

constexpr int CMND_THREADS = 256;   // 8 warps, each takes every 8th autocorrelation chunk
__global__ void __launch_bounds__(CMND_THREADS, 2) k_pyin_cmnd(const PyinParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  CmndSmem s;
  cmnd_smem_carve(p, smem_raw, &s);
  const int nthr = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  PyinTile t;
  if (!pyin_locate(p, blockIdx.x, &t)) return;
  cmnd_phase_load(p, t, s, tid, nthr);
  __syncthreads();
  {
    const int cpb = cmnd_cpb(p), lpc = cmnd_lpc(p);
    const int n_chunks = cmnd_n_blocks(p, t.nf) * cpb;
    for (int c = warp; c < n_chunks; c += CMND_SLOTS) {
      const int m = c / cpb, g = (c - m * cpb) * lpc + lane;
      if (lane < lpc && g < p.n_groups) cmnd_acf_unit(p, s, m, g);
    }
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();
  cmnd_phase_diff(p, t, s, tid, nthr);
  __syncthreads();
  for (int f = warp; f < t.nf; f += CMND_SLOTS) {
    cmnd_phase_scan1(p, s, f, warp, lane);
    __syncwarp();
    cmnd_phase_scan2(p, s, warp, lane);
    __syncwarp();
    cmnd_phase_emit(p, t, s, f, warp, lane);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) k_pyin_probs(const PyinParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int wpb = blockDim.x / 32;
  const size_t per = prob_smem_carve(p, nullptr, nullptr);
  ProbSmem s;
  prob_smem_carve(p, smem_raw + (size_t)warp * per, &s);
  // thresholds staged once per CTA behind the per-warp areas
  double* thr = (double*)(smem_raw + (size_t)wpb * per);
  double* beta = thr + p.n_thr + 2;
  double* bexp = beta + p.n_thr + 2;
  double* bfact = bexp + p.kmax + 2;
  for (int i = threadIdx.x; i <= p.n_thr; i += blockDim.x) thr[i] = p.thresholds[i];
  for (int i = threadIdx.x; i < p.n_thr; i += blockDim.x) beta[i] = p.beta_probs[i];
  for (int i = threadIdx.x; i <= p.kmax; i += blockDim.x) { bexp[i] = p.boltz_exp[i]; bfact[i] = p.boltz_fact[i]; }
  __syncthreads();
  for (int64_t frame = (int64_t)blockIdx.x * wpb + warp; frame < p.total_frames;
       frame += (int64_t)gridDim.x * wpb) {
    prob_phase0(p, s, frame, lane); __syncwarp();
    {   // troughs in lag order: lane-strided test + ballot compaction (same list as prob_phase1/2)
      int base = 0;
      for (int i0 = 0; i0 < p.n_lags; i0 += 32) {
        const int i = i0 + lane;
        const bool tr = i < p.n_lags && prob_is_trough(s.x, i, p.n_lags);
        const unsigned m = __ballot_sync(0xffffffffu, tr);
        if (tr) s.tr[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
        base += __popc(m);
      }
      if (lane == 0) s.cnt[32] = base;
    }
    __syncwarp();
    {   // prob_phase3 + prob_phase4: first threshold per trough; global minimum trough (lowest height, first index)
        // by a shuffle arg-min over the per-lane partials instead of a serial pass of lane 0
      const int R = s.cnt[32];
      double best = 1e300; int bi = 0x7fffffff;
      for (int r = lane; r < R; r += 32) {
        const double h = s.x[s.tr[r]];
        int lo = 0, hi = p.n_thr;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (thr[mid] <= h) lo = mid; else hi = mid - 1;
        }
        s.cr[r] = (uint8_t)lo;
        if (h < best) { best = h; bi = r; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double oh = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oh < best || (oh == best && oi < bi)) { best = oh; bi = oi; }
      }
      if (lane == 0) s.cnt[33] = bi;
    }
    __syncwarp();
    {
      const int R = s.cnt[32];
      // compact the active troughs in lag order
      int na = 0;
      for (int base = 0; base < R; base += 32) {
        const int r = base + lane;
        const bool act = r < R && (int)s.cr[r] < p.n_thr;
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (act) s.sorted[na + __popc(m & ((1u << lane) - 1u))] = (uint16_t)r;
        na += __popc(m);
      }
      __syncwarp();
      // inactive troughs: probability 0 (the global minimum still gets the no-trough bonus)
      for (int r = lane; r < R; r += 32) if ((int)s.cr[r] >= p.n_thr) prob_trough_finish(p, s, r, 0.0);
      const unsigned lt_mask = (1u << lane) - 1u;
      // hist[c] = #{active a : c_a == c} (n_c is its running sum); carry[c] = the same over the
      // chunks already processed (ranks of later chunks start from it)
      int* hist = s.hist;
      int* carry = s.hist + p.n_thr;
      for (int c = lane; c < 2 * p.n_thr; c += 32) hist[c] = 0;
      __syncwarp();
      int cmin = 0x7fff;
      for (int a = lane; a < na; a += 32) {
        const int cr = (int)s.cr[s.sorted[a]];
        atomicAdd(&hist[cr], 1);
        cmin = cr < cmin ? cr : cmin;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const int x = __shfl_xor_sync(0xffffffffu, cmin, o); cmin = x < cmin ? x : cmin; }
      __syncwarp();
      for (int base = 0; base < na; base += 32) {
        const int a = base + lane;
        const int r = a < na ? (int)s.sorted[a] : -1;
        const int cr = r >= 0 ? (int)s.cr[r] : 0x7fff;
        const bool more = base + 32 < na;
        double acc = 0.0, K = 0.0;     // K = fact[n_c] * exp(-lambda * pos): changes only where a trough starts
        int pos = 0, n = 0;
        if (p.n_thr <= 128) {
          // thresholds where some active trough starts (hist != 0) as a bit mask; between two of them K is
          // constant, so the inner loop is just  acc += K * beta[c]  (same additions, same order)
          // the remaining change points as four warp-uniform words, popped lowest bit first
          unsigned rem0 = __ballot_sync(0xffffffffu, lane < p.n_thr && hist[lane] != 0);
          unsigned rem1 = __ballot_sync(0xffffffffu, 32 + lane < p.n_thr && hist[32 + lane] != 0);
          unsigned rem2 = __ballot_sync(0xffffffffu, 64 + lane < p.n_thr && hist[64 + lane] != 0);
          unsigned rem3 = __ballot_sync(0xffffffffu, 96 + lane < p.n_thr && hist[96 + lane] != 0);
#define PROB_POP_CP(dst)                                                              \
          if (rem0) { dst = __ffs(rem0) - 1; rem0 &= rem0 - 1; }                       \
          else if (rem1) { dst = 31 + __ffs(rem1); rem1 &= rem1 - 1; }                 \
          else if (rem2) { dst = 63 + __ffs(rem2); rem2 &= rem2 - 1; }                 \
          else if (rem3) { dst = 95 + __ffs(rem3); rem3 &= rem3 - 1; }                 \
          else dst = p.n_thr;
          int c;
          PROB_POP_CP(c)
          while (c < p.n_thr) {
            {   // change point c
              const unsigned m = __ballot_sync(0xffffffffu, cr == c);
              const int cy = carry[c];
              n += hist[c];
              pos += cy + __popc(m & lt_mask);
              if (more && lane == 0 && m) carry[c] = cy + __popc(m);
              K = bfact[n] * bexp[pos];
            }
            int cn;
            PROB_POP_CP(cn)
            int cc = c > cr ? c : cr;
            for (; cc < cn; ++cc) acc += K * beta[cc];
            c = cn;
          }
#undef PROB_POP_CP
        } else {
          for (int c = cmin; c < p.n_thr; ++c) {
            const int hc = hist[c];      // warp-uniform
            if (hc) {
              const unsigned m = __ballot_sync(0xffffffffu, cr == c);
              const int cy = carry[c];
              n += hc;
              pos += cy + __popc(m & lt_mask);
              if (more && lane == 0 && m) carry[c] = cy + __popc(m);
              K = bfact[n] * bexp[pos];
            }
            if (c >= cr) acc += K * beta[c];
          }
        }
        if (r >= 0) prob_trough_finish(p, s, r, acc);
        __syncwarp();
      }
    }
    __syncwarp();
    {   // prob_phase6a: a trough survives NumPy's last-write-wins scatter unless the NEXT trough with a non-zero
        // probability lands on the same pitch bin.  Bins are non-increasing in lag order (troughs are at least two
        // lags apart and the parabolic shift is < 1 in magnitude, so refined periods strictly increase), hence
        // duplicates of a bin are adjacent among the non-zero troughs: one look at the successor decides.
        // Chunks of 32 troughs from the last to the first, the successor's bin carried across chunks.
      const int R = s.cnt[32];
      int nb = -2;                                   // bin of the nearest non-zero trough after this chunk
      for (int base = ((R - 1) >> 5) << 5; base >= 0; base -= 32) {
        const int r = base + lane;
        const int b = r < R ? (int)s.bin[r] : -1;
        const unsigned nz = __ballot_sync(0xffffffffu, b >= 0);
        const unsigned above = lane == 31 ? 0u : nz >> (lane + 1);
        const int src = above ? lane + __ffs(above) : lane;
        const int sb = __shfl_sync(0xffffffffu, b, src);
        const int nxt = above ? sb : nb;
        if (r < R) s.live[r] = (b >= 0 && b < p.npb && nxt != b) ? 1 : 0;
        if (nz) nb = __shfl_sync(0xffffffffu, b, __ffs(nz) - 1);
      }
    }
    __syncwarp();
    {   // prob_phase6b with the logarithms spread over the lanes; the voiced-probability sum stays serial
      const int R = s.cnt[32];
      uint16_t* ob = p.cand_bin + (size_t)frame * p.kmax;
      double* ol = p.cand_lp + (size_t)frame * p.kmax;
      int n = 0;
      double vp = 0.0;     // summed in descending trough order (= ascending pitch bin), every lane the same chain
      for (int b0 = 0; b0 < R; b0 += 32) {
        const int r = R - 1 - (b0 + lane);
        const bool live = r >= 0 && s.live[r];
        const unsigned m = __ballot_sync(0xffffffffu, live);
        const double pr = live ? s.prob[r] : 0.0;
        if (live) {
          const int pos = n + __popc(m & ((1u << lane) - 1u));
          ob[pos] = (uint16_t)s.bin[r];
          ol[pos] = log(pr + 2.2250738585072014e-308);
        }
        for (unsigned mm = m; mm; mm &= mm - 1) vp += __shfl_sync(0xffffffffu, pr, __ffs(mm) - 1);
        n += __popc(m);
      }
      if (lane == 0) {
        if (vp < 0.0) vp = 0.0;
        if (vp > 1.0) vp = 1.0;
        p.n_cand[frame] = n;
        p.lp_unvoiced[frame] = log((1.0 - vp) / (double)p.npb + 2.2250738585072014e-308);
        p.voiced_prob[frame] = (float)vp;
      }
    }
    __syncwarp();
  }
}
#endif  // __CUDACC__

}  // namespace roar
