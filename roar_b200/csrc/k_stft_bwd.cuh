// K1b  backward of K1 for FilterbankFeatures(use_grads=True): d loss / d audio from d loss / d features.
//
// Replaces autograd through FilterbankFeatures.forward (asr/parts/preprocessing/features.py:384-434)
// as used by the JETS / HiFi-GAN / BigVGAN / RoarTTS mel losses (tts/models/jets.py:175-177,
// hifigan.py:56-58, bigvgan.py:57-59, roar_tts.py:174-176: `use_grads=True`, normalize None, no
// pre-emphasis).  SURVEY.md section 8f, row N4.
//
// Same tiling as the forward kernel (one CTA = FT frames of one utterance) and the forward spectrum is
// recomputed rather than stored:
//   forward  : window -> n_fft/2 complex FFT -> untangle X[k] -> mag = sqrt(re^2+im^2+guard) -> S = mag^p
//              -> mel = fb S -> out = log(clamp | add)
//   backward : g_mel = g_out * d log            (0 on masked frames / below the clamp)
//              g_S[k] = sum_m fb[m,k] g_mel[m]  (shared-memory atomics over the CSR band rows)
//              G[k]   = g_S[k] * p * mag^(p-2) * X[k]
//              g_x[n] = Re sum_{k=0..M} G[k] e^{+2 pi i k n / N}  -- evaluated as an inverse REAL FFT:
//                       C[0] = Re G[0], C[M] = Re G[M], C[k] = G[k]/2, packed into M complex points,
//                       inverse half-size FFT, z[n] = (g_x[2n], g_x[2n+1])
//              overlap-add of window * g_x into the tile's span (shared memory), then atomicAdd into
//              the gradient buffer through the same reflect mapping the forward pass reads with.
#pragma once
#include "k_stft_mel.cuh"

namespace roar {

struct StftBwdParams {
  StftParams f;               // forward geometry / tables / audio (logmel, energy unused)
  const float* grad_out;      // [B, n_mels, Tpad] (f.out_utt_stride / f.out_row_stride address it)
  const int64_t* valid_len;   // [B] frames that are not masked (features.py:446-452); may be null
  float* grad_audio;          // [B, Lmax] zeroed by the caller
};

struct StftBwdSmem {
  float* spec;      // [G][M+1]  mag^p
  float* mag;       // [G][M+1]
  float* gS;        // [G][M+1]
  float* gspan;     // [span]    overlap-added gradient of the tile's (padded) samples
};
HD size_t stft_bwd_extra_carve(const StftParams& p, unsigned char* base, StftBwdSmem* s) {
  size_t o = 0;
#define CARVE(field, type, count) { if (s) s->field = (type*)(base + o); o = align16(o + sizeof(type) * (size_t)(count)); }
  CARVE(spec, float, (size_t)p.G * (p.M + 1))
  CARVE(mag, float, (size_t)p.G * (p.M + 1))
  CARVE(gS, float, (size_t)p.G * (p.M + 1))
  CARVE(gspan, float, p.span)
#undef CARVE
  return o;
}

HD void smem_add(float* addr, float v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(addr, v);
#else
  *addr += v;
#endif
}

// X[k] for k = 0..M from the half-size FFT (same untangle as stft_phase_post); X -> xbuf[slot][k]
HD void bwd_phase_spectrum(const StftParams& p, const StftTile& t, StftBwdSmem& b, int g, int tid,
                           const cf32* z_base, cf32* x_base) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const cf32* Z = z_base + (size_t)slot * pidx(p.M);
  cf32* X = x_base + (size_t)slot * pidx(p.M);
  for (int k = u; k <= p.M; k += p.P) {
    const cf32 zk = Z[pidx(k & (p.M - 1))];
    const cf32 zc = cconj(Z[pidx((p.M - k) & (p.M - 1))]);
    cf32 e; e.x = 0.5f * (zk.x + zc.x); e.y = 0.5f * (zk.y + zc.y);
    cf32 d; d.x = 0.5f * (zk.y - zc.y); d.y = -0.5f * (zk.x - zc.x);
    const cf32 w = ld_ro(p.tw_post + k);
    cf32 x;
    x.x = e.x + (w.x * d.x - w.y * d.y);
    x.y = e.y + (w.x * d.y + w.y * d.x);
    const float mag = sqrtf(x.x * x.x + x.y * x.y + p.floor_);
    // X is consumed un-padded (k up to M inclusive fits: pidx(M) >= M + 1 for M >= 8)
    X[k] = x;
    b.mag[(size_t)slot * (p.M + 1) + k] = mag;
    b.spec[(size_t)slot * (p.M + 1) + k] = stft_pow(mag, p.mag_power);
    b.gS[(size_t)slot * (p.M + 1) + k] = 0.f;
  }
}

// mel forward + d log, scattered back over the band: gS[k] += fb[m,k] * g_mel[m]
HD void bwd_phase_mel(const StftBwdParams& q, const StftTile& t, StftSmem& s, StftBwdSmem& b, int g, int tid, int nthr) {
  const StftParams& p = q.f;
  for (int w = tid; w < p.G * p.n_mels; w += nthr) {
    const int slot = w / p.n_mels, m = w - slot * p.n_mels;
    const int f = g * p.G + slot;
    if (f >= t.nf) continue;
    const int tt = t.t0 + f;
    if (q.valid_len && tt >= q.valid_len[t.utt]) continue;       // masked frame: constant output
    const float* spec = b.spec + (size_t)slot * (p.M + 1) + s.mel_start[m];
    const float* wgt = s.mel_w + s.mel_offset[m];
    const int cnt = s.mel_count[m];
    float mel = 0.f;
    for (int c = 0; c < cnt; ++c) mel += wgt[c] * spec[c];
    float dlog = 1.f;
    if (p.log_mode == ROAR_LOG_CLAMP) dlog = mel >= p.log_guard ? 1.f / mel : 0.f;   // torch.clamp passes grad where x >= min
    else if (p.log_mode == ROAR_LOG_ADD) dlog = 1.f / (mel + p.log_guard);
    const int64_t base = p.out_utt_stride ? (int64_t)t.utt * p.out_utt_stride : (int64_t)p.n_mels * p.frame_off[t.utt];
    const int64_t rs = p.out_row_stride ? p.out_row_stride : t.T;
    const float gm = q.grad_out[base + (int64_t)m * rs + tt] * dlog;
    if (gm == 0.f) continue;
    float* gs = b.gS + (size_t)slot * (p.M + 1) + s.mel_start[m];
    for (int c = 0; c < cnt; ++c) smem_add(gs + c, wgt[c] * gm);
  }
}

// G[k] = gS[k] * p * mag^(p-2) * X[k], halved / real-only at the ends (C of the header), in place in X
HD void bwd_phase_gx(const StftParams& p, const StftTile& t, StftBwdSmem& b, int g, int tid, cf32* x_base) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  cf32* X = x_base + (size_t)slot * pidx(p.M);
  for (int k = u; k <= p.M; k += p.P) {
    const float mag = b.mag[(size_t)slot * (p.M + 1) + k];
    const float gs = b.gS[(size_t)slot * (p.M + 1) + k];
    float fac;
    if (p.mag_power == 1.0f) fac = gs / mag;
    else if (p.mag_power == 2.0f) fac = 2.0f * gs;
    else fac = gs * p.mag_power * powf(mag, p.mag_power - 2.0f);
    cf32 c = X[k];
    c.x *= fac; c.y *= fac;
    if (k == 0 || k == p.M) c.y = 0.f; else { c.x *= 0.5f; c.y *= 0.5f; }
    X[k] = c;
  }
}

// pack the Hermitian half spectrum C[0..M] into M complex points for the half-size inverse FFT:
//   Zc[k] = (C[k] + conj(C[M-k])) + i e^{+2 pi i k / N} (C[k] - conj(C[M-k]))
HD void bwd_phase_pack(const StftParams& p, const StftTile& t, int g, int tid, const cf32* x_base, cf32* z_base) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const cf32* C = x_base + (size_t)slot * pidx(p.M);
  cf32* Z = z_base + (size_t)slot * pidx(p.M);
  for (int k = u; k < p.M; k += p.P) {
    const cf32 ck = C[k], cm = cconj(C[p.M - k]);
    const cf32 a = cadd(ck, cm), d = csub(ck, cm);
    cf32 w = ld_ro(p.tw_post + k); w.y = -w.y;             // e^{+2 pi i k / N}
    const cf32 wd = cmul(w, d);
    cf32 r; r.x = a.x - wd.y; r.y = a.y + wd.x;            // a + i * wd
    Z[pidx(k)] = r;
  }
}

template <int R>
HD void bwd_pass_inv(const StftParams& p, const StftTile& t, int g, int tid, int Ns, const cf32* twp,
                     const cf32* in_base, cf32* out_base) {
  const int slot = tid / p.P, u = tid - slot * p.P;
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const cf32* in = in_base + (size_t)slot * pidx(p.M);
  cf32* out = out_base + (size_t)slot * pidx(p.M);
  const int nb = p.M / R;
  for (int j = u; j < nb; j += p.P) {
    cf32 v[R];
    stockham_load<R>(v, in, p.M, j);
    stockham_twiddle_dft<R, true, cf32, false>(v, Ns, j, twp);
    stockham_store<R>(v, out, Ns, j);
  }
}
HD void bwd_pass_inv_any(int R, const StftParams& p, const StftTile& t, int g, int tid, int Ns, const cf32* twp,
                         const cf32* in_base, cf32* out_base) {
  if (R == 8) bwd_pass_inv<8>(p, t, g, tid, Ns, twp, in_base, out_base);
  else if (R == 4) bwd_pass_inv<4>(p, t, g, tid, Ns, twp, in_base, out_base);
  else bwd_pass_inv<2>(p, t, g, tid, Ns, twp, in_base, out_base);
}

// window * g_x of one frame slot added into the tile span (slots are taken one at a time: they overlap)
HD void bwd_phase_accumulate(const StftParams& p, const StftTile& t, StftSmem& s, StftBwdSmem& b, int g, int slot,
                             int tid, int nthr, const cf32* z_base) {
  const int f = g * p.G + slot;
  if (f >= t.nf) return;
  const cf32* Z = z_base + (size_t)slot * pidx(p.M);
  float* dst = b.gspan + (size_t)f * p.hop;
  for (int n = tid; n < p.M; n += nthr) {
    const cf32 z = Z[pidx(n)];
    dst[2 * n] += s.window[2 * n] * z.x;
    dst[2 * n + 1] += s.window[2 * n + 1] * z.y;
  }
}

// span gradient -> global gradient through the forward read mapping (reflect padding folds back)
HD void bwd_phase_scatter(const StftBwdParams& q, const StftTile& t, StftBwdSmem& b, int tid, int nthr) {
  const StftParams& p = q.f;
  const int n = (t.nf - 1) * p.hop + p.n_fft;
  for (int i = tid; i < n; i += nthr) {
    const float v = b.gspan[i];
    if (v == 0.f) continue;
    const int64_t pos = stft_reflect(t.p0 + i, t.L);
    if (pos < 0) continue;
    smem_add(q.grad_audio + t.off + pos, v);       // atomicAdd on the device (global address)
  }
}

#ifdef __CUDACC__
template <int LG>
__global__ void __launch_bounds__(256) k_stft_mel_bwd(const StftBwdParams q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const StftParams& p = q.f;
  StftSmem s;
  const int nthr = blockDim.x, tid = threadIdx.x;
  const size_t fwd_bytes = stft_smem_carve(p, nthr, smem_raw, &s);
  StftBwdSmem b;
  stft_bwd_extra_carve(p, smem_raw + fwd_bytes, &b);
  StftTile t;
  if (!stft_locate(p, blockIdx.x, &t)) return;
  stft_phase_tables(p, s, tid, nthr);
  const int span_n = (t.nf - 1) * p.hop + p.n_fft;
  stft_phase_audio(p, t, s, tid, nthr, 0, span_n);
  for (int i = tid; i < p.span; i += nthr) b.gspan[i] = 0.f;
  __syncthreads();
  constexpr int NP8 = LG / 3, REM = LG % 3;
  const int n_groups = (t.nf + p.G - 1) / p.G;
  for (int g = 0; g < n_groups; ++g) {
    // ---- forward spectrum of the group's frames
    stft_first_pass<8>(p, t, s, g, tid);
    __syncthreads();
    int Ns = 8, off = 0;
    cf32* src = s.bufA;
    cf32* dst = s.bufB;
#pragma unroll
    for (int ps = 1; ps < NP8; ++ps) {
      stft_pass<8>(p, t, s, g, tid, Ns, s.tw + off, src, dst);
      __syncthreads();
      off += 7 * Ns; Ns *= 8;
      cf32* tmp = src; src = dst; dst = tmp;
    }
    if (REM != 0) {
      if (REM == 2) stft_pass<4>(p, t, s, g, tid, Ns, s.tw + off, src, dst);
      else stft_pass<2>(p, t, s, g, tid, Ns, s.tw + off, src, dst);
      __syncthreads();
      cf32* tmp = src; src = dst; dst = tmp;
    }
    // src = Z (half-size FFT), dst free -> X
    bwd_phase_spectrum(p, t, b, g, tid, src, dst);
    __syncthreads();
    bwd_phase_mel(q, t, s, b, g, tid, nthr);
    __syncthreads();
    bwd_phase_gx(p, t, b, g, tid, dst);
    __syncthreads();
    bwd_phase_pack(p, t, g, tid, dst, src);                 // packed spectrum back into `src`
    __syncthreads();
    // ---- inverse half-size FFT: same radix plan, conjugated twiddles, first pass without twiddles
    Ns = 1; off = 0;
#pragma unroll
    for (int ps = 0; ps < NP8; ++ps) {
      bwd_pass_inv<8>(p, t, g, tid, Ns, s.tw + off, src, dst);
      __syncthreads();
      if (ps > 0) off += 7 * Ns;
      Ns *= 8;
      cf32* tmp = src; src = dst; dst = tmp;
    }
    if (REM != 0) {
      // twiddle block offset of the last pass: after NP8 radix-8 passes (the first has none)
      int off2 = 0, ns2 = 8;
      for (int ps = 1; ps < NP8; ++ps) { off2 += 7 * ns2; ns2 *= 8; }
      if (REM == 2) bwd_pass_inv<4>(p, t, g, tid, Ns, s.tw + off2, src, dst);
      else bwd_pass_inv<2>(p, t, g, tid, Ns, s.tw + off2, src, dst);
      __syncthreads();
      cf32* tmp = src; src = dst; dst = tmp;
    }
    for (int slot = 0; slot < p.G; ++slot) {
      bwd_phase_accumulate(p, t, s, b, g, slot, tid, nthr, src);
      __syncthreads();
    }
  }
  bwd_phase_scatter(q, t, b, tid, nthr);
}
#endif  // __CUDACC__

}  // namespace roar
