// K4 beta-binomial alignment prior, K5 pitch-statistics partials, K6 FilterbankFeatures normalise/mask.
#pragma once
#include "common.cuh"
#include "../../include/roar_sup.h"

namespace roar {

// ------------------------------------------------------------------------------------ K4 prior
// Replaces beta_binomial_prior_distribution (tts/parts/utils/tts_dataset_utils.py:128-149):
//   P[y-1, k] = exp( lnC(n,k) + lnB(k+a, n-k+b) - lnB(a,b) ),  n=N-1, a=s*y, b=s*(M+1-y), y=1..M, k=0..N-1
// The reference evaluates this in FLOAT32 (torch.special.gammaln on float32 tensors, float32 adds), and
// its rounding noise (~1e-3 relative in P) decides the row arg-max near every mode crossover, so the
// kernel reproduces that arithmetic instead of the exact value: with the scaling factor s == 1 every
// gammaln argument is a positive integer, torch's float32 gammaln of an integer is the correctly rounded
// float32 of ln((i-1)!) (checked for i < 5000 in tests/), taken here from a float64 log-factorial table
// built once per handle, and the nine terms are combined with the reference's own association:
//   logcomb = (G(n+1) - G(k+1)) - G(n-k+1);  logbeta(u, v) = (G(u) + G(v)) - G(u+v);
//   s = (logcomb + logbeta(k+a, n-k+b)) - logbeta(a, b);   P = exp(s).
// Output is write-bound: 4*M*N bytes.
struct PriorParams {
  const int32_t* text_len;
  const int32_t* mel_len;
  const int64_t* out_off;
  float* out;
  const double* lf;       // ln(i!) for i < lf_n
  int32_t lf_n;
  int32_t utt_base;       // utterance index of blockIdx.y == 0
  int32_t rows_per_cta;
  double scaling;
};

HD double prior_lf(const PriorParams& p, int i) { return i < p.lf_n ? p.lf[i] : lgamma((double)i + 1.0); }
// float32 gammaln(i) for integer i >= 1, correctly rounded
HD float prior_g32(const PriorParams& p, int i) { return (float)prior_lf(p, i - 1); }
HD float prior_add32(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
HD float prior_sub32(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
// float32 log-probability exactly as the reference forms it (s == 1)
HD float prior_logp32(const PriorParams& p, int N, int M, int y, int k) {
  const int n = N - 1;
  const float lc = prior_sub32(prior_sub32(prior_g32(p, n + 1), prior_g32(p, k + 1)), prior_g32(p, n - k + 1));
  const float lb1 = prior_sub32(prior_add32(prior_g32(p, k + y), prior_g32(p, n - k + M + 1 - y)), prior_g32(p, n + M + 1));
  const float lb2 = prior_sub32(prior_add32(prior_g32(p, y), prior_g32(p, M + 1 - y)), prior_g32(p, M + 1));
  return prior_sub32(prior_add32(lc, lb1), lb2);
}
HD float prior_value_int(const PriorParams& p, int N, int M, int y, int k) {
  return (float)exp((double)prior_logp32(p, N, M, y, k));
}
HD float prior_value_real(const PriorParams& p, int N, int M, int y, int k) {
  const double n = N - 1, a = p.scaling * y, b = p.scaling * (M + 1 - y);
  const double lc = lgamma(n + 1) - lgamma(k + 1.0) - lgamma(n - k + 1);
  const double lb1 = lgamma(k + a) + lgamma(n - k + b) - lgamma(n + a + b);
  const double lb2 = lgamma(a) + lgamma(b) - lgamma(a + b);
  return (float)exp(lc + lb1 - lb2);
}

// BetaBinomialInterpolator.__call__(w = mel_len, h = text_len) (tts_dataset_utils.py:69-92): the exact prior at
// sizes rounded to multiples of (round_mel, round_text) -- with the reference's swapped roles,
// bank(bw, bh) = prior(phoneme_count = bw, mel_count = bh), transposed -- resampled to [w, h] by
// scipy.ndimage.zoom(order=1): output index o reads input coordinate o * (in - 1) / (out - 1), linear.
// The LRU bank of the reference is a CPU cache; here the four taps are evaluated on the fly.
HD int prior_round(int val, int to) {
  // max(1, int(np.round((val + 1) / to))) * to ; np.round = round half to even
  const double q = (double)(val + 1) / (double)to;
  const int r = (int)rint(q);
  return (r < 1 ? 1 : r) * to;
}
HD float prior_interp_value(const PriorParams& p, int w, int h, int bw, int bh, int i, int j) {
  // scipy's NI_ZoomShift arithmetic, reproduced literally: zoom = (in - 1) / (out - 1) as one double
  // division, coordinate = zoom * index; mode="constant" turns a coordinate that rounding pushed past
  // in - 1 into cval = 0 (the reference inherits this: e.g. the last row of a 75 x 15 prior is zero).
  const double zx = w > 1 ? (double)(bw - 1) / (double)(w - 1) : 1.0;
  const double zy = h > 1 ? (double)(bh - 1) / (double)(h - 1) : 1.0;
  const double x = zx * (double)i, y = zy * (double)j;
  if (x > (double)(bw - 1) || y > (double)(bh - 1)) return 0.f;
  const int i0 = (int)floor(x), j0 = (int)floor(y);
  const int i1 = i0 + 1 < bw ? i0 + 1 : bw - 1, j1 = j0 + 1 < bh ? j0 + 1 : bh - 1;
  const double fx = x - i0, fy = y - j0;
  // base[i'][j'] = prior(N = bw, M = bh)[j'][i']  (mel index j', token index i')
  const double b00 = prior_value_int(p, bw, bh, j0 + 1, i0), b01 = prior_value_int(p, bw, bh, j1 + 1, i0);
  const double b10 = prior_value_int(p, bw, bh, j0 + 1, i1), b11 = prior_value_int(p, bw, bh, j1 + 1, i1);
  const double a0 = b00 + (b01 - b00) * fy, a1 = b10 + (b11 - b10) * fy;
  return (float)(a0 + (a1 - a0) * fx);
}

#ifdef __CUDACC__
struct PriorInterpParams { PriorParams base; int32_t round_mel, round_text; };
__global__ void __launch_bounds__(256) k_align_prior_interp(const PriorInterpParams q) {
  const PriorParams& p = q.base;
  const int utt = p.utt_base + blockIdx.y;
  const int h = p.text_len[utt], w = p.mel_len[utt];
  const int r0 = blockIdx.x * p.rows_per_cta;
  if (r0 >= w || h <= 0) return;
  const int r1 = r0 + p.rows_per_cta < w ? r0 + p.rows_per_cta : w;
  const int bw = prior_round(w, q.round_mel), bh = prior_round(h, q.round_text);
  float* out = p.out + p.out_off[utt];
  const int64_t e0 = (int64_t)r0 * h, e1 = (int64_t)r1 * h;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int i = (int)(e / h), j = (int)(e - (int64_t)i * h);
    out[e] = prior_interp_value(p, w, h, bw, bh, i, j);
  }
}

// The float32 log-probability of prior_logp32 splits into a column term lc(k), a row term lb2(y) and the mixed
// term lb1(k, y); each is a self-contained sub-expression of the reference's formula, so evaluating lc once per
// column and lb2 once per row (shared memory) and only lb1 per element gives bit-identical values with a third of
// the table look-ups and conversions.  Elements are walked with incremental (row, column) indices: no division.
constexpr int PRIOR_MAX_N = 2048;      // columns staged in shared memory (longer texts: the per-element path)
__global__ void __launch_bounds__(256) k_align_prior(const PriorParams p) {
  __shared__ float s_lc[PRIOR_MAX_N];
  __shared__ float s_lb2[64];
  const int utt = p.utt_base + blockIdx.y;
  const int N = p.text_len[utt], M = p.mel_len[utt];
  const int r0 = blockIdx.x * p.rows_per_cta;
  if (r0 >= M || N <= 0) return;
  const int r1 = r0 + p.rows_per_cta < M ? r0 + p.rows_per_cta : M;
  float* out = p.out + p.out_off[utt];
  const int64_t e0 = (int64_t)r0 * N, e1 = (int64_t)r1 * N;
  const bool integer = p.scaling == 1.0;
  if (!integer || N > PRIOR_MAX_N || p.rows_per_cta > 64) {
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
      const int m = (int)(e / N), k = (int)(e - (int64_t)m * N);
      out[e] = integer ? prior_value_int(p, N, M, m + 1, k) : prior_value_real(p, N, M, m + 1, k);
    }
    return;
  }
  const int n = N - 1;
  for (int k = threadIdx.x; k < N; k += blockDim.x)
    s_lc[k] = prior_sub32(prior_sub32(prior_g32(p, n + 1), prior_g32(p, k + 1)), prior_g32(p, n - k + 1));
  for (int r = threadIdx.x; r < r1 - r0; r += blockDim.x) {
    const int y = r0 + r + 1;
    s_lb2[r] = prior_sub32(prior_add32(prior_g32(p, y), prior_g32(p, M + 1 - y)), prior_g32(p, M + 1));
  }
  __syncthreads();
  const float g_nm = prior_g32(p, n + M + 1);
  // thread's first element and its stride in (row, column) form
  int row = (int)threadIdx.x / N, k = (int)threadIdx.x - row * N;
  const int drow = (int)blockDim.x / N, dk = (int)blockDim.x - drow * N;
  const int rows = r1 - r0;
  while (row < rows) {
    const int y = r0 + row + 1;
    const float lb1 = prior_sub32(prior_add32(prior_g32(p, k + y), prior_g32(p, n - k + M + 1 - y)), g_nm);
    const float lp = prior_sub32(prior_add32(s_lc[k], lb1), s_lb2[row]);
    out[e0 + (int64_t)row * N + k] = (float)exp((double)lp);
    row += drow; k += dk;
    if (k >= N) { k -= N; ++row; }
  }
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------ K7 silence trim
// librosa.effects.trim(samples, top_db, ref, frame_length, hop_length) as called by AudioSegment
// (asr/parts/preprocessing/segment.py:76-88): frame RMS (center=True, zero padding), dB relative to
// the reference (np.max of the RMS by default), first / last frame above -top_db ->
// [start, end) = [f0 * hop, min(L, (f1 + 1) * hop)).  One CTA per utterance, one warp per frame, the
// frame powers kept in shared memory (float64 sums of float32 squares).
struct TrimParams {
  const float* audio;
  const int64_t* sample_off;
  const int32_t* sample_len;
  int32_t n_utts, frame_length, hop_length, max_frames;
  double top_db, ref_value;     // ref_value <= 0: np.max over the utterance's frames
  int64_t* start;
  int64_t* end;
};
HD bool trim_is_loud(double power, double ref_power, double top_db) {
  const double amin = 1e-10;     // amin ** 2 with amin = 1e-5 (amplitude_to_db)
  const double db = 10.0 * log10(power > amin ? power : amin) - 10.0 * log10(ref_power > amin ? ref_power : amin);
  return db > -top_db;
}
// mean square of frame f (zero-padded by frame_length / 2 on both sides), lanes stride the samples
HD double trim_frame_power_part(const float* y, int L, int frame_length, int hop, int f, int lane, int nl) {
  const int64_t s0 = (int64_t)f * hop - frame_length / 2;
  double acc = 0.0;
  for (int i = lane; i < frame_length; i += nl) {
    const int64_t q = s0 + i;
    if (q >= 0 && q < L) { const double v = y[q]; acc += v * v; }
  }
  return acc;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(256) k_trim(const TrimParams p) {
  extern __shared__ __align__(16) double t_pow[];      // [max_frames]
  __shared__ double s_wmax[8];
  __shared__ int s_first[8], s_last[8];
  const int utt = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int L = p.sample_len[utt];
  const float* y = p.audio + p.sample_off[utt];
  const int T = 1 + L / p.hop_length;
  double wmax = 0.0;
  for (int f = warp; f < T; f += nwarp) {
    double acc = trim_frame_power_part(y, L, p.frame_length, p.hop_length, f, lane, 32);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const double pw = acc / (double)p.frame_length;
    if (lane == 0) t_pow[f] = pw;
    wmax = pw > wmax ? pw : wmax;
  }
  if (lane == 0) s_wmax[warp] = wmax;
  __syncthreads();
  double ref_power = p.ref_value * p.ref_value;
  if (p.ref_value <= 0.0) {
    ref_power = 0.0;
    for (int w = 0; w < nwarp; ++w) ref_power = s_wmax[w] > ref_power ? s_wmax[w] : ref_power;
    // ref = np.max(rms): (sqrt(max power)) ** 2 in the reference; sqrt then square in float64 here
    const double r = sqrt(ref_power);
    ref_power = r * r;
  }
  int first = 0x7fffffff, last = -1;
  for (int f = tid; f < T; f += blockDim.x) {
    // the reference squares the float32 RMS again: rms = sqrt(power) (float32), power' = rms ** 2
    const float rms32 = sqrtf((float)t_pow[f]);
    const double pw = (double)rms32 * (double)rms32;
    if (trim_is_loud(pw, ref_power, p.top_db)) { first = f < first ? f : first; last = f > last ? f : last; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int a = __shfl_xor_sync(0xffffffffu, first, o), b = __shfl_xor_sync(0xffffffffu, last, o);
    first = a < first ? a : first; last = b > last ? b : last;
  }
  if (lane == 0) { s_first[warp] = first; s_last[warp] = last; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 0; w < nwarp; ++w) { first = s_first[w] < first ? s_first[w] : first; last = s_last[w] > last ? s_last[w] : last; }
    int64_t st = 0, en = 0;
    if (last >= 0) {
      st = (int64_t)first * p.hop_length;
      en = (int64_t)(last + 1) * p.hop_length;
      if (en > L) en = L;
    }
    p.start[utt] = st; p.end[utt] = en;
  }
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------ K8 polyphase resampler
// Rational-rate FIR resampling of a packed batch: upsample by `up`, low-pass, downsample by `down`, with the
// arithmetic of scipy.signal.resample_poly / upfirdn (zero extension at the ends):
//   out[n] = sum_k taps[(n + n_pre_remove) * down - n_pre_pad - k * up] * x[k]
// over the k whose tap index falls inside [0, n_taps).  `taps` is the host-built low-pass (Kaiser-windowed sinc,
// already multiplied by `up`).  Replaces the resampling step of AudioSegment.__init__
// (asr/parts/preprocessing/segment.py:68-75), where the reference calls librosa.core.resample (soxr_hq): a
// different filter design of the same band limit -- parity unpinned, see oracle/resample.py.
struct ResampleParams {
  const float* in;
  const int64_t* in_off;
  const int32_t* in_len;
  const int64_t* out_off;
  const int32_t* out_len;
  float* out;
  const float* taps;
  int32_t n_utts, n_taps, up, down, n_pre_pad, n_pre_remove;
};
HD float resample_one(const ResampleParams& p, const float* x, int L, int n) {
  const int64_t c = ((int64_t)n + p.n_pre_remove) * p.down - p.n_pre_pad;
  // k*up <= c  and  c - k*up <= n_taps - 1
  int64_t k_hi = c >= 0 ? c / p.up : -1;
  int64_t lo_num = c - (p.n_taps - 1);
  int64_t k_lo = lo_num <= 0 ? 0 : (lo_num + p.up - 1) / p.up;
  if (k_hi > L - 1) k_hi = L - 1;
  double acc = 0.0;
  for (int64_t k = k_lo; k <= k_hi; ++k) acc += (double)p.taps[c - k * p.up] * (double)x[k];
  return (float)acc;
}

#ifdef __CUDACC__
// grid.y = utterance, grid.x * blockDim.x covers the longest output
__global__ void __launch_bounds__(256) k_resample(const ResampleParams p, int32_t utt_base) {
  const int utt = utt_base + blockIdx.y;
  const int n_out = p.out_len[utt], L = p.in_len[utt];
  const float* x = p.in + p.in_off[utt];
  float* y = p.out + p.out_off[utt];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_out; n += gridDim.x * blockDim.x) y[n] = resample_one(p, x, L, n);
}

// ------------------------------------------------------------------------------------ K0 16-bit PCM ingest
// int16 -> float32 / 2^15 (exact).  Grid-stride, 8 samples per thread per trip: one 16-byte load, two
// 16-byte stores when both pointers are 16-byte aligned; scalar head/tail otherwise.
__global__ void __launch_bounds__(256) k_pcm16_to_f32(const int16_t* __restrict__ pcm, int64_t n, float* __restrict__ out) {
  const float sc = 1.0f / 32768.0f;
  const bool vec = ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int64_t n8 = vec ? n / 8 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(pcm) + i);
    float4 a, b;
    a.x = (float)(short)(v.x & 0xffff) * sc; a.y = (float)(short)(v.x >> 16) * sc;
    a.z = (float)(short)(v.y & 0xffff) * sc; a.w = (float)(short)(v.y >> 16) * sc;
    b.x = (float)(short)(v.z & 0xffff) * sc; b.y = (float)(short)(v.z >> 16) * sc;
    b.z = (float)(short)(v.w & 0xffff) * sc; b.w = (float)(short)(v.w >> 16) * sc;
    reinterpret_cast<float4*>(out)[2 * i] = a;
    reinterpret_cast<float4*>(out)[2 * i + 1] = b;
  }
  for (int64_t i = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)pcm[i] * sc;
}

// Small metadata upload (offsets, lengths, frame prefix sums): the SMs read the pinned host block directly
// (unified addressing) and write it to HBM.  A few KB per call; unlike cudaMemcpyAsync it does not queue
// behind the multi-hundred-MB audio copies on the host-to-device copy engine, so the kernels that need the
// metadata are not held up by them.  Both pointers 16-byte aligned, n16 = number of 16-byte words.
__global__ void __launch_bounds__(256) k_upload_small(const uint4* __restrict__ host_src, uint4* __restrict__ dst, int64_t n16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = host_src[i];
}

// ------------------------------------------------------------------------------------ K5 stats
// get_pitch_stats (scripts/dataset_processing/tts/extract_sup_data.py:8-13): mean / unbiased std /
// min / max over pitch != 0.  Produces float64 partials (sum, sumsq, count, min, max) that the ranks
// all-reduce.  min/max use the unsigned-integer order of positive doubles.
__device__ __forceinline__ void stats_merge(double* out5, double s, double q, double c, double mn, double mx) {
  if (c > 0) {
    atomicAdd(&out5[0], s);
    atomicAdd(&out5[1], q);
    atomicAdd(&out5[2], c);
    atomicMin((unsigned long long*)&out5[3], (unsigned long long)__double_as_longlong(mn));
    atomicMax((unsigned long long*)&out5[4], (unsigned long long)__double_as_longlong(mx));
  }
}
__device__ __forceinline__ void stats_warp_reduce(double& s, double& q, double& c, double& mn, double& mx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
}
__global__ void k_stats_init(double* out, int n_groups) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_groups) {
    out[i * 5 + 0] = 0; out[i * 5 + 1] = 0; out[i * 5 + 2] = 0;
    out[i * 5 + 3] = __longlong_as_double(0x7FF0000000000000LL);  // +inf
    out[i * 5 + 4] = 0;
  }
}
__global__ void __launch_bounds__(256) k_pitch_partials(const float* f0, int64_t n, double* out5) {
  double s = 0, q = 0, c = 0, mn = __longlong_as_double(0x7FF0000000000000LL), mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = f0[i];
    if (v != 0.f) { const double d = v; s += d; q += d * d; c += 1; mn = fmin(mn, d); mx = fmax(mx, d); }
  }
  stats_warp_reduce(s, q, c, mn, mx);
  __shared__ double sh[5][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = q; sh[2][warp] = c; sh[3][warp] = mn; sh[4][warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? sh[0][lane] : 0; q = lane < nw ? sh[1][lane] : 0; c = lane < nw ? sh[2][lane] : 0;
    mn = lane < nw ? sh[3][lane] : __longlong_as_double(0x7FF0000000000000LL);
    mx = lane < nw ? sh[4][lane] : 0;
    stats_warp_reduce(s, q, c, mn, mx);
    if (lane == 0) stats_merge(out5, s, q, c, mn, mx);
  }
}
// per-utterance groups (speaker ids): one warp per utterance, one merge per utterance
__global__ void __launch_bounds__(256) k_pitch_partials_grouped(const float* f0, const int64_t* frame_off,
                                                                const int32_t* group, int32_t n_utts,
                                                                int32_t n_groups, double* out) {
  const int utt = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (utt >= n_utts) return;
  const int g = group[utt];
  double s = 0, q = 0, c = 0, mn = __longlong_as_double(0x7FF0000000000000LL), mx = 0;
  for (int64_t i = frame_off[utt] + lane; i < frame_off[utt + 1]; i += 32) {
    const float v = f0[i];
    if (v != 0.f) { const double d = v; s += d; q += d * d; c += 1; mn = fmin(mn, d); mx = fmax(mx, d); }
  }
  stats_warp_reduce(s, q, c, mn, mx);
  if (lane == 0 && g >= 0 && g < n_groups) stats_merge(out + (size_t)g * 5, s, q, c, mn, mx);
}

// ------------------------------------------------------------------------------------ K6 normalise
// normalize_batch + masked_fill + pad (features.py:25-61, 444-460) on out[B, n_mels, Tpad], in place.
struct NormParams {
  float* x;
  const int64_t* seq_len;   // frames valid per utterance
  int32_t B, n_mels, T_full, Tpad, mode;
  float pad_value;
};
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}
// grid = (n_mels, B) for per_feature / none, (1, B) for all_features
__global__ void __launch_bounds__(256) k_fbank_normalize(const NormParams p) {
  __shared__ double sh[8];
  const int b = blockIdx.y;
  int64_t n = p.seq_len[b];
  if (n > p.T_full) n = p.T_full;
  if (n < 0) n = 0;
  const int rows = p.mode == ROAR_NORM_ALL_FEATURES ? p.n_mels : 1;
  float* base = p.x + ((size_t)b * p.n_mels + (p.mode == ROAR_NORM_ALL_FEATURES ? 0 : blockIdx.x)) * p.Tpad;
  double mean = 0, inv = 1;
  if (p.mode != ROAR_NORM_NONE) {
    double s = 0;
    for (int r = 0; r < rows; ++r)
      for (int64_t t = threadIdx.x; t < n; t += blockDim.x) s += base[(size_t)r * p.Tpad + t];
    const double cnt = (double)n * rows;
    mean = block_sum(s, sh) / cnt;
    double q = 0;
    for (int r = 0; r < rows; ++r)
      for (int64_t t = threadIdx.x; t < n; t += blockDim.x) { const double d = base[(size_t)r * p.Tpad + t] - mean; q += d * d; }
    const double var = block_sum(q, sh) / (cnt - 1.0);
    // reference: float32 mean/std, std += 1e-5
    const float stdf = (float)sqrt(var) + 1e-5f;
    inv = 1.0;
    const float meanf = (float)mean;
    for (int r = 0; r < rows; ++r)
      for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
        float* px = base + (size_t)r * p.Tpad + t;
        *px = (*px - meanf) / stdf;
      }
  }
  (void)inv;
  for (int r = 0; r < rows; ++r)
    for (int64_t t = n + threadIdx.x; t < p.Tpad; t += blockDim.x) base[(size_t)r * p.Tpad + t] = p.pad_value;
}
#endif  // __CUDACC__

}  // namespace roar
