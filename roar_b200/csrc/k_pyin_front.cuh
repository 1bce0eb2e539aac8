// K2  pYIN front end.
//   K2a  pyin_cmnd : per frame, YIN difference function -> cumulative-mean-normalised difference.
//        The autocorrelation uses the reference's own formulation (float64 FFT of the frame and of
//        its reversed first half, product, inverse FFT) as ONE packed complex FFT of size F plus one
//        half-size inverse, in shared memory; the energy terms replay the reference's float32
//        sequential cumsum bit for bit (it carries ~1e-4 relative noise into the trough heights, so
//        replaying it is what makes the thresholds agree).
//   K2b  pyin_probs: per frame (one warp), parabolic shifts, troughs, threshold-beta / Boltzmann
//        probabilities, pitch-bin quantisation -> sparse observation list + voiced probability.
//
// Replaces librosa.pyin's `_cumulative_mean_normalized_difference`, `_parabolic_interpolation`,
// `__pyin_helper` as called from roar/collections/tts/data/dataset.py:696-703.
#pragma once
#include "common.cuh"
#include "fft.cuh"

namespace roar {

struct PyinParams {
  const float* audio;
  const int64_t* sample_off;
  const int32_t* sample_len;
  const int64_t* frame_off;   // [n_utts+1] pyin frames
  const int32_t* tile_off;    // [n_utts+1]
  int32_t n_utts;
  // geometry
  int32_t F, W, hop, H;       // frame, win, hop, F/2
  int32_t min_period, max_period, n_lags;
  int32_t FT, span, P, G;     // frames per tile, audio span, threads per frame, frames in flight
  int32_t npb, nbps, kmax, n_thr;
  double sr, fmin, no_trough_prob;
  // tables
  const cf64* tw;             // [F] W_F^k (float64), read through L1 (inverse packing)
  const cf64* tw_f;           // per-pass Stockham twiddles, size F (fft.cuh layout)
  const cf64* tw_h;           // per-pass Stockham twiddles, size F/2
  const double* thresholds;   // [n_thr+1]
  const double* beta_probs;   // [n_thr]
  const double* beta_cum;     // [n_thr+1]
  const double* boltz_exp;    // [kmax+1]
  const double* boltz_fact;   // [kmax+1]
  // scratch / outputs
  double* cmnd;               // [total_frames, n_lags]
  uint16_t* cand_bin;         // [total_frames, kmax]
  double* cand_lp;            // [total_frames, kmax]
  int32_t* n_cand;            // [total_frames]
  double* lp_unvoiced;        // [total_frames]
  float* voiced_prob;         // [total_frames]  (final output)
  int64_t total_frames;
};

struct CmndSmem {
  float* audio;    // [span]
  cf64* buf;       // [G][pidx(F)]
  float* E;        // [FT][max_period+1]
  double* d;       // [G][max_period+1]   difference function, then chunk-local prefix sums
  double* dsum;    // [G][max_period+1]
  double* chunk;   // [G][32] chunk totals, [G][32] chunk offsets
};

HD size_t cmnd_align16(size_t x) { return (x + 15) & ~(size_t)15; }

HD size_t cmnd_smem_carve(const PyinParams& p, unsigned char* base, CmndSmem* s) {
  size_t o = 0;
#define CARVE(field, type, count) { if (s) s->field = (type*)(base + o); o = cmnd_align16(o + sizeof(type) * (size_t)(count)); }
  CARVE(buf, cf64, (size_t)p.G * pidx(p.F) + 8)
  CARVE(d, double, (size_t)p.G * (p.max_period + 1))
  CARVE(dsum, double, (size_t)p.G * (p.max_period + 1))
  CARVE(chunk, double, (size_t)p.G * 64)
  CARVE(audio, float, p.span)
  CARVE(E, float, (size_t)p.FT * (p.max_period + 1))
#undef CARVE
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
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (p.tile_off[mid] <= tile) lo = mid; else hi = mid;
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

HD void cmnd_phase_load(const PyinParams& p, const PyinTile& t, CmndSmem& s, int tid, int nthr) {
  const int n = (t.nf - 1) * p.hop + p.F;
  for (int i = tid; i < n; i += nthr) {
    const int64_t q = t.p0 + i;
    s.audio[i] = (q >= 0 && q < t.L) ? p.audio[t.off + q] : 0.f;
  }
}

// float32 sequential cumsum of y^2 (np.cumsum(y_frames**2, axis=-2) on float32 frames), one thread
// per frame; keeps only what the difference function needs: E[tau] = cs[W+tau] - cs[tau].
// The frames of a tile are spread over warps (lane 0/1 of each) so the strided reads do not pile
// onto one bank.
HD float f32_sq_add(float cs, float y) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(cs, __fmul_rn(y, y));
#else
  volatile float sq = y * y;
  volatile float r = cs + sq;
  return r;
#endif
}
HD void cmnd_phase_energy(const PyinParams& p, const PyinTile& t, CmndSmem& s, int tid, int nthr) {
  const int nwarp = nthr / 32 > 0 ? nthr / 32 : 1;
  const int warp = tid / 32, lane = tid % 32;
  const int f = lane * nwarp + warp;
  if (lane >= (p.FT + nwarp - 1) / nwarp || f >= t.nf) return;
  const float* y = s.audio + f * p.hop;
  float* E = s.E + (size_t)f * (p.max_period + 1);
  // three straight loops (no load/store aliasing inside a loop, so the loads pipeline):
  //   A: cs[n] for n <= max_period, kept in E;  B: keep accumulating up to W-1;
  //   C: n = W + tau: E[tau] = cs[W+tau] - cs[tau]
  float cs = 0.f;
  int n = 0;
  for (; n <= p.max_period && n < p.W; ++n) { cs = f32_sq_add(cs, y[n]); E[n] = cs; }
  for (; n < p.W; ++n) cs = f32_sq_add(cs, y[n]);
  for (int tau = 0; tau <= p.max_period; ++tau) {
    cs = f32_sq_add(cs, y[p.W + tau]);
    if (p.W + tau <= p.max_period) {   // only when max_period >= W (non-default win_length)
      const float keep = cs;
      const float lo0 = E[tau];
      E[p.W + tau] = keep;
      float e0 = cs - lo0;
      if (fabsf(e0) < 1e-6f) e0 = 0.f;
      E[tau] = e0;
      continue;
    }
#if defined(__CUDA_ARCH__)
    float e = __fsub_rn(cs, E[tau]);
#else
    float e = cs - E[tau];
#endif
    if (fabsf(e) < 1e-6f) e = 0.f;
    E[tau] = e;
  }
}

// first radix-8 pass of the packed FFT: z[n] = y[n] + i * yrev[n], yrev[k] = y[W-k] (k < W), 0 after
HD void cmnd_first_pass(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const float* y = s.audio + f * p.hop;
  cf64* out = s.buf + (size_t)slot * pidx(p.F);
  const int nb = p.F / 8;
  for (int j = u; j < nb; j += p.P) {
    cf64 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int n = j + r * nb;
      v[r].x = (double)y[n];
      v[r].y = n < p.W ? (double)y[p.W - n] : 0.0;
    }
    dft8<false>(v);
    stockham_store<8>(v, out, 1, j);
  }
}

// in-place pass: load+compute into registers (phase a), barrier, store (phase b).
// Each thread owns at most 8/R butterflies (M/P <= 8), fully unrolled so `regs` stays in registers.
template <int R, bool INV>
HD void cmnd_pass_compute(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid, int M,
                          int Ns, const cf64* twp, const cf64* in_base, size_t slot_stride, cf64* regs) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const cf64* in = in_base + (size_t)slot * slot_stride;
  const int nb = M / R;
#pragma unroll
  for (int q = 0; q < 8 / R; ++q) {
    const int j = u + q * p.P;
    if (j < nb) {
      cf64* v = regs + q * R;
      stockham_load<R>(v, in, M, j);
      stockham_twiddle_dft<R, INV>(v, Ns, j, twp);
    }
  }
}
template <int R>
HD void cmnd_pass_store(const PyinParams& p, const PyinTile& t, int g, int tid, int M, int Ns,
                        cf64* out_base, size_t slot_stride, const cf64* regs) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  cf64* out = out_base + (size_t)slot * slot_stride;
  const int nb = M / R;
#pragma unroll
  for (int q = 0; q < 8 / R; ++q) {
    const int j = u + q * p.P;
    if (j < nb) stockham_store<R>(regs + q * R, out, Ns, j);
  }
}

// spectra of the two real sequences from the packed transform, their product, and the packing of
// the half-size inverse:  C[k] = A[k]*B[k];  Zr[k] = (C[k]+conj(C[H-k])) + i e^{+2 pi i k/F} (C[k]-conj(C[H-k]))
HD void cmnd_phase_product(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  cf64* Z = s.buf + (size_t)slot * pidx(p.F);
  for (int k = u; k <= p.H; k += p.P) {
    const cf64 zk = Z[pidx(k)];
    const cf64 zc = cconj(Z[pidx((p.F - k) & (p.F - 1))]);
    cf64 A, B;
    A.x = 0.5 * (zk.x + zc.x); A.y = 0.5 * (zk.y + zc.y);
    B.x = 0.5 * (zk.y - zc.y); B.y = -0.5 * (zk.x - zc.x);   // (zk - zc)/(2i)
    Z[pidx(k)] = cmul(A, B);
  }
}
HD void cmnd_phase_pack_inverse(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  cf64* C = s.buf + (size_t)slot * pidx(p.F);
  for (int k = u; k <= p.H / 2; k += p.P) {
    const int k2 = p.H - k;
    const cf64 ck = C[pidx(k)], c2 = C[pidx(k2)];
    // k
    {
      cf64 a = cadd(ck, cconj(c2)), b = csub(ck, cconj(c2));
      cf64 w = ld_ro(p.tw + k); w.y = -w.y;          // e^{+2 pi i k / F}
      cf64 wb = cmul(w, b);
      cf64 r; r.x = a.x - wb.y; r.y = a.y + wb.x;    // a + i*wb
      C[pidx(k)] = r;
    }
    if (k2 != k && k2 < p.H) {
      cf64 a = cadd(c2, cconj(ck)), b = csub(c2, cconj(ck));
      cf64 w = ld_ro(p.tw + k2); w.y = -w.y;
      cf64 wb = cmul(w, b);
      cf64 r; r.x = a.x - wb.y; r.y = a.y + wb.x;
      C[pidx(k2)] = r;
    }
  }
}

// difference function from the inverse transform (real sequence r[m] = interleaved re/im of z2)
HD void cmnd_phase_diff(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid,
                        const cf64* z2_base) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const cf64* z2 = z2_base + (size_t)slot * pidx(p.F);
  const float* E = s.E + (size_t)f * (p.max_period + 1);
  double* d = s.d + (size_t)slot * (p.max_period + 1);
  const double scale = 1.0 / p.F;
  for (int tau = u; tau <= p.max_period; tau += p.P) {
    const int m = p.W + tau;
    const cf64 zz = z2[pidx(m >> 1)];
    double acf = ((m & 1) ? zz.y : zz.x) * scale;
    if (fabs(acf) < 1e-6) acf = 0.0;
#if defined(__CUDA_ARCH__)
    const float e2 = __fadd_rn(E[0], E[tau]);
#else
    volatile float e2v = E[0] + E[tau]; const float e2 = e2v;
#endif
    d[tau] = (double)e2 - 2.0 * acf;
  }
}

// cumulative sum of d[1..max_period] in three steps (chunk-local prefix, chunk offsets, combine)
HD int cmnd_chunk(const PyinParams& p) { return (p.max_period + 31) / 32; }

HD void cmnd_phase_scan1(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf || u >= 32) return;
  const int ch = cmnd_chunk(p);
  const double* d = s.d + (size_t)slot * (p.max_period + 1);
  double* ds = s.dsum + (size_t)slot * (p.max_period + 1);
  double acc = 0.0;
  for (int i = 0; i < ch; ++i) {
    const int tau = 1 + u * ch + i;
    if (tau > p.max_period) break;
    acc += d[tau];
    ds[tau] = acc;
  }
  s.chunk[slot * 32 + u] = acc;
}
HD void cmnd_phase_scan2(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid,
                         double* offs /* [G][32] */) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf || u >= 32) return;
  double acc = 0.0;
  for (int v = 0; v < u; ++v) acc += s.chunk[slot * 32 + v];
  offs[slot * 32 + u] = acc;
}
HD void cmnd_phase_emit(const PyinParams& p, const PyinTile& t, CmndSmem& s, int g, int tid,
                        const double* offs) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const int ch = cmnd_chunk(p);
  const double* d = s.d + (size_t)slot * (p.max_period + 1);
  const double* ds = s.dsum + (size_t)slot * (p.max_period + 1);
  double* out = p.cmnd + (size_t)(p.frame_off[t.utt] + t.t0 + f) * p.n_lags;
  for (int i = u; i < p.n_lags; i += p.P) {
    const int tau = p.min_period + i;
    const double c = ds[tau] + offs[slot * 32 + (tau - 1) / ch];
    out[i] = d[tau] / (c / (double)tau + 2.2250738585072014e-308);
  }
}

// =================================================================================== K2b
struct ProbSmem {
  double* x;          // [n_lags]
  double* prob;       // [kmax]
  uint16_t* tr;       // [kmax] trough lag indices (increasing)
  uint16_t* sorted;   // [kmax] trough ranks ordered by (c_r, r)
  int16_t* bin;       // [kmax]
  uint8_t* cr;        // [kmax] first threshold index the trough is below (n_thr = never)
  int32_t* cnt;       // [36]
  double* red;        // [32]
  int32_t* carry;     // [n_thr] troughs of earlier chunks per first-threshold index
  uint8_t* live;      // [kmax]
};
HD size_t prob_smem_carve(const PyinParams& p, unsigned char* base, ProbSmem* s) {
  size_t o = 0;
#define CARVE(field, type, count) { if (s) s->field = (type*)(base + o); o = cmnd_align16(o + sizeof(type) * (size_t)(count)); }
  CARVE(x, double, p.n_lags + 2)
  CARVE(prob, double, p.kmax)
  CARVE(red, double, 32)
  CARVE(tr, uint16_t, p.kmax)
  CARVE(sorted, uint16_t, (p.kmax > 2 * p.n_thr ? p.kmax : 2 * p.n_thr) + 2)   // also n_all[n_thr] ints
  CARVE(carry, int32_t, p.n_thr)
  CARVE(live, uint8_t, p.kmax)
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

// Threshold-beta / Boltzmann probabilities.  Lanes own troughs (chunks of 32 in lag order); the loop
// over the thresholds c is warp-uniform.  n_c = #{r : c_r <= c} is uniform; pos(r, c) =
// #{r' < r : c_r' <= c} advances by the number of earlier troughs whose first threshold is c, which
// is a ballot + popcount inside the chunk plus a per-threshold carry from the previous chunks.
//   probs_r = sum_{c >= c_r} (fact[n_c] * exp(-lambda * pos)) * beta[c]   (+ no-trough bonus on the global min)
// `hist` [n_thr] (ints) aliases s.sorted (unused otherwise).
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

// n_all[c] = #{r : c_r <= c} over ALL troughs (kept in smem, [n_thr] ints, aliases s.sorted)
HD void prob_phase5a(const PyinParams& p, ProbSmem& s, int lane) {
  const int R = s.cnt[32];
  int* n_all = reinterpret_cast<int*>(s.sorted);
  for (int c = lane; c < p.n_thr; c += 32) {
    int n = 0;
    for (int r = 0; r < R; ++r) n += ((int)s.cr[r] <= c) ? 1 : 0;
    n_all[c] = n;
  }
}
// one chunk of 32 troughs; `carry[c]` (ints, [n_thr], aliases s.bin's tail? no: own array s.carry)
HD void prob_phase5b_lane(const PyinParams& p, ProbSmem& s, int base, int lane, const int* chunk_cr /*[32]*/) {
  // reference formulation used by the host harness: identical arithmetic, ballot replaced by a scan
  const int R = s.cnt[32];
  const int r = base + lane;
  if (r >= R) return;
  const int* n_all = reinterpret_cast<const int*>(s.sorted);
  const int cr = s.cr[r];
  double acc = 0.0;
  int pos = 0;
  // earlier chunks: every trough r' < base counts when c_r' <= c
  for (int c = 0; c < p.n_thr; ++c) {
    int inc = s.carry[c];
    for (int l = 0; l < lane; ++l) inc += (chunk_cr[l] == c) ? 1 : 0;
    pos += inc;
    if (c >= cr) acc += (p.boltz_fact[n_all[c]] * p.boltz_exp[pos]) * p.beta_probs[c];
  }
  prob_trough_finish(p, s, r, acc);
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
__global__ void __launch_bounds__(256) k_pyin_cmnd(const PyinParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  CmndSmem s;
  cmnd_smem_carve(p, smem_raw, &s);
  const int nthr = blockDim.x, tid = threadIdx.x;
  PyinTile t;
  if (!pyin_locate(p, blockIdx.x, &t)) return;
  cmnd_phase_load(p, t, s, tid, nthr);
  __syncthreads();
  cmnd_phase_energy(p, t, s, tid, nthr);
  __syncthreads();
  const FftPlan planF = make_plan(p.F);
  const FftPlan planH = make_plan(p.H);
  const size_t FP = (size_t)pidx(p.F);
  cf64 regs[8];
  const int n_groups = (t.nf + p.G - 1) / p.G;
  for (int g = 0; g < n_groups; ++g) {
    cmnd_first_pass(p, t, s, g, tid);
    __syncthreads();
    int Ns = 8;
    for (int ps = 1; ps < planF.n_pass; ++ps) {
      const int R = planF.radix[ps];
      const cf64* twp = p.tw_f + planF.tw_off[ps];
      if (R == 8) cmnd_pass_compute<8, false>(p, t, s, g, tid, p.F, Ns, twp, s.buf, FP, regs);
      else if (R == 4) { cmnd_pass_compute<4, false>(p, t, s, g, tid, p.F, Ns, twp, s.buf, FP, regs); }
      else { cmnd_pass_compute<2, false>(p, t, s, g, tid, p.F, Ns, twp, s.buf, FP, regs); }
      __syncthreads();
      if (R == 8) cmnd_pass_store<8>(p, t, g, tid, p.F, Ns, s.buf, FP, regs);
      else if (R == 4) cmnd_pass_store<4>(p, t, g, tid, p.F, Ns, s.buf, FP, regs);
      else cmnd_pass_store<2>(p, t, g, tid, p.F, Ns, s.buf, FP, regs);
      __syncthreads();
      Ns *= R;
    }
    cmnd_phase_product(p, t, s, g, tid);
    __syncthreads();
    cmnd_phase_pack_inverse(p, t, s, g, tid);
    __syncthreads();
    // half-size inverse, in place as well (keeps one buffer per frame)
    Ns = 1;
    for (int ps = 0; ps < planH.n_pass; ++ps) {
      const int R = planH.radix[ps];
      const cf64* twp = p.tw_h + planH.tw_off[ps];
      if (R == 8) cmnd_pass_compute<8, true>(p, t, s, g, tid, p.H, Ns, twp, s.buf, FP, regs);
      else if (R == 4) cmnd_pass_compute<4, true>(p, t, s, g, tid, p.H, Ns, twp, s.buf, FP, regs);
      else cmnd_pass_compute<2, true>(p, t, s, g, tid, p.H, Ns, twp, s.buf, FP, regs);
      __syncthreads();
      if (R == 8) cmnd_pass_store<8>(p, t, g, tid, p.H, Ns, s.buf, FP, regs);
      else if (R == 4) cmnd_pass_store<4>(p, t, g, tid, p.H, Ns, s.buf, FP, regs);
      else cmnd_pass_store<2>(p, t, g, tid, p.H, Ns, s.buf, FP, regs);
      __syncthreads();
      Ns *= R;
    }
    cmnd_phase_diff(p, t, s, g, tid, s.buf);
    __syncthreads();
    cmnd_phase_scan1(p, t, s, g, tid);
    __syncthreads();
    cmnd_phase_scan2(p, t, s, g, tid, s.chunk + (size_t)p.G * 32);
    __syncthreads();
    cmnd_phase_emit(p, t, s, g, tid, s.chunk + (size_t)p.G * 32);
    __syncthreads();
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
  for (int i = threadIdx.x; i <= p.n_thr; i += blockDim.x) thr[i] = p.thresholds[i];
  __syncthreads();
  for (int64_t frame = (int64_t)blockIdx.x * wpb + warp; frame < p.total_frames;
       frame += (int64_t)gridDim.x * wpb) {
    prob_phase0(p, s, frame, lane); __syncwarp();
    prob_phase1(p, s, lane); __syncwarp();
    prob_phase2(p, s, lane); __syncwarp();
    prob_phase3(p, s, lane, thr); __syncwarp();
    prob_phase4(p, s, lane); __syncwarp();
    prob_phase5a(p, s, lane);
    for (int c = lane; c < p.n_thr; c += 32) s.carry[c] = 0;
    __syncwarp();
    {
      const int R = s.cnt[32];
      const int* n_all = reinterpret_cast<const int*>(s.sorted);
      const unsigned lt_mask = (1u << lane) - 1u;
      for (int base = 0; base < R; base += 32) {
        const int r = base + lane;
        const int cr = r < R ? (int)s.cr[r] : 0x7fff;
        double acc = 0.0;
        int pos = 0;
        for (int c = 0; c < p.n_thr; ++c) {
          const unsigned m = __ballot_sync(0xffffffffu, cr == c);
          const int carry = s.carry[c];
          pos += carry + __popc(m & lt_mask);
          if (c >= cr) acc += (p.boltz_fact[n_all[c]] * p.boltz_exp[pos]) * p.beta_probs[c];
          if (lane == 0 && m) s.carry[c] = carry + __popc(m);
        }
        if (r < R) prob_trough_finish(p, s, r, acc);
        __syncwarp();
      }
    }
    __syncwarp();
    prob_phase6a(p, s, lane); __syncwarp();
    prob_phase6b(p, s, frame, lane); __syncwarp();
  }
}
#endif  // __CUDACC__

}  // namespace roar
