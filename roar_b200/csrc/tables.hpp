// Host-side table construction (float64) for libroar_sup: window, twiddles, sparse mel rows,
// pYIN threshold/boltzmann tables and the banded log-transition rows of the pitch HMM.
// Pure C++ (no CUDA): also compiled into the CPU test harness.
//
// Arithmetic being restated (reference = AshwinSankar17/Roar, paths relative to its root):
//   window        torch.hann_window(win_length, periodic=False) & co, zero-padded centred by
//                 torch.stft              roar/collections/tts/data/dataset.py:324-333
//   mel rows      librosa.filters.mel     dataset.py:305-314; asr/parts/preprocessing/features.py:297-308
//   pYIN tables   librosa.pyin defaults   dataset.py:696-703 (librosa 0.10.x core/pitch.py, sequence.py)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../include/roar_sup.h"
#include "fft.cuh"

namespace roar {

static const double kPi = 3.14159265358979323846;
static const double kTiny64 = 2.2250738585072014e-308;  // np.finfo(np.float64).tiny

// ---------------------------------------------------------------------------------- config
struct Geometry {
  int n_fft, win, hop, n_bins, n_mels, M;  // M = n_fft/2 complex points
  // pyin
  int pf, pw, ph;              // frame, win, hop
  int min_period, max_period;  // lags used: min_period..max_period
  int n_lags;                  // max_period - min_period + 1
  int npb, nbps;               // pitch bins, bins per semitone
  int tw, hw;                  // transition width / half width
  int kmax;                    // max candidates per frame
  int n_thr;
};

inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

inline std::string validate(const roar_sup_config& c) {
  if (c.struct_size != (int32_t)sizeof(roar_sup_config)) return "roar_sup_config.struct_size mismatch (ABI)";
  if (c.sample_rate <= 0) return "sample_rate must be > 0";
  if (!is_pow2(c.n_fft) || c.n_fft < 64 || c.n_fft > 4096) return "n_fft must be a power of two in [64, 4096]";
  if (c.win_length <= 0 || c.win_length > c.n_fft) return "win_length must be in (0, n_fft]";
  if (c.hop_length <= 0) return "hop_length must be > 0";
  if (c.n_mels <= 0 || c.n_mels > 512) return "n_mels must be in (0, 512]";
  if (c.window < 0 || c.window > 4) return "unknown window";
  if (c.exact_pad && (c.hop_length % 2 == 1)) return "exact_pad requires an even hop_length";
  if (c.pitch_fmin <= 0 || c.pitch_fmax <= c.pitch_fmin) return "need 0 < pitch_fmin < pitch_fmax";
  // pyin_frame_length == 0: a mel-only handle (FilterbankFeatures) -- no pYIN tables, roar_sup_pyin is rejected
  if (c.pyin_frame_length != 0 &&
      (!is_pow2(c.pyin_frame_length) || c.pyin_frame_length < 64 || c.pyin_frame_length > 4096))
    return "pyin_frame_length must be 0 (no pYIN) or a power of two in [64, 4096]";
  if (c.n_thresholds <= 0 || c.n_thresholds > 256) return "n_thresholds must be in (0, 256]";
  return "";
}

inline Geometry geometry(const roar_sup_config& c) {
  Geometry g;
  g.n_fft = c.n_fft; g.win = c.win_length; g.hop = c.hop_length;
  g.n_bins = c.n_fft / 2 + 1; g.n_mels = c.n_mels; g.M = c.n_fft / 2;
  g.pf = c.pyin_frame_length;
  if (g.pf == 0) {   // mel-only handle
    g.pw = g.ph = g.min_period = g.max_period = g.n_lags = g.npb = g.nbps = g.tw = g.hw = g.kmax = 0;
    g.n_thr = c.n_thresholds;
    return g;
  }
  g.pw = c.pyin_win_length > 0 ? c.pyin_win_length : g.pf / 2;
  g.ph = c.pyin_hop_length > 0 ? c.pyin_hop_length : g.pf / 4;
  g.min_period = (int)std::floor((double)c.sample_rate / c.pitch_fmax);
  int mp = (int)std::ceil((double)c.sample_rate / c.pitch_fmin);
  g.max_period = mp < g.pf - g.pw - 1 ? mp : g.pf - g.pw - 1;
  g.n_lags = g.max_period - g.min_period + 1;
  g.nbps = (int)std::ceil(1.0 / c.resolution);
  g.npb = (int)std::floor(12.0 * g.nbps * std::log2(c.pitch_fmax / c.pitch_fmin)) + 1;
  // Python round() = round-half-even on the float64 product
  double r = c.max_transition_rate * 12.0 * g.ph / c.sample_rate;
  int msf = (int)std::nearbyint(r);
  g.tw = msf * g.nbps + 1;
  g.hw = g.tw / 2;
  g.kmax = (g.n_lags + 1) / 2 + 1;
  g.n_thr = c.n_thresholds;
  return g;
}

// ---------------------------------------------------------------------------------- window
inline std::vector<float> make_window(const roar_sup_config& c) {
  const int N = c.win_length;
  std::vector<double> w(N, 1.0);
  if (N > 1) {
    for (int n = 0; n < N; ++n) {
      double x = (double)n / (N - 1);
      switch (c.window) {
        case ROAR_WIN_HANN:     w[n] = 0.5 - 0.5 * std::cos(2 * kPi * x); break;
        case ROAR_WIN_HAMMING:  w[n] = 0.54 - 0.46 * std::cos(2 * kPi * x); break;
        case ROAR_WIN_BLACKMAN: w[n] = 0.42 - 0.5 * std::cos(2 * kPi * x) + 0.08 * std::cos(4 * kPi * x); break;
        case ROAR_WIN_BARTLETT: w[n] = 1.0 - std::fabs(2.0 * x - 1.0); break;
        default: w[n] = 1.0;
      }
    }
  }
  std::vector<float> out(c.n_fft, 0.f);
  const int left = (c.n_fft - N) / 2;  // torch.stft pads the window centred
  for (int n = 0; n < N; ++n) out[left + n] = (float)w[n];
  return out;
}

// ---------------------------------------------------------------------------------- mel
inline double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
inline double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// float32 [n_mels, n_bins], same rounding sequence as librosa (float32 store, then *= float64 norm)
inline std::vector<float> make_mel_filterbank(const roar_sup_config& c) {
  const int n_mels = c.n_mels, n_bins = c.n_fft / 2 + 1;
  const double fmax = c.fmax > 0 ? c.fmax : c.sample_rate / 2.0;
  std::vector<double> mel_f(n_mels + 2);
  const double m0 = hz_to_mel(c.fmin), m1 = hz_to_mel(fmax);
  // np.linspace: start + i*step, last element forced to stop
  const double step = (m1 - m0) / (n_mels + 1);
  for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(i == n_mels + 1 ? m1 : m0 + i * step);
  std::vector<float> w((size_t)n_mels * n_bins, 0.f);
  // np.fft.rfftfreq(n, d=1/sr): (arange * (1/(n*d))) with val = 1.0/(n*d)
  const double val = 1.0 / (c.n_fft * (1.0 / c.sample_rate));
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < n_bins; ++k) {
      const double fk = k * val;
      const double lower = -(mel_f[i] - fk) / fd0;
      const double upper = (mel_f[i + 2] - fk) / fd1;
      double v = lower < upper ? lower : upper;
      if (!(v > 0)) v = 0;
      float f = (float)v;
      if (c.mel_norm) f = (float)((double)f * enorm);
      w[(size_t)i * n_bins + k] = f;
    }
  }
  return w;
}

// banded (start, count) representation + packed weights; rows are contiguous triangles
struct MelRows {
  std::vector<int32_t> start, count, offset;
  std::vector<float> weights;
  int max_count = 0;
  int last_bin = 0;  // highest bin with any weight
};
inline MelRows make_mel_rows(const std::vector<float>& fb, int n_mels, int n_bins) {
  MelRows r;
  for (int i = 0; i < n_mels; ++i) {
    int a = -1, b = -1;
    for (int k = 0; k < n_bins; ++k)
      if (fb[(size_t)i * n_bins + k] != 0.f) { if (a < 0) a = k; b = k; }
    if (a < 0) { a = 0; b = -1; }
    r.start.push_back(a);
    r.count.push_back(b - a + 1);
    r.offset.push_back((int32_t)r.weights.size());
    for (int k = a; k <= b; ++k) r.weights.push_back(fb[(size_t)i * n_bins + k]);  // interior zeros kept
    if (b - a + 1 > r.max_count) r.max_count = b - a + 1;
    if (b > r.last_bin) r.last_bin = b;
  }
  return r;
}

// ---------------------------------------------------------------------------------- twiddles
template <class T2, class T>
inline std::vector<T2> make_twiddles(int n, int count, double sign = -1.0) {
  std::vector<T2> t(count);
  for (int k = 0; k < count; ++k) {
    double a = sign * 2.0 * kPi * k / n;
    t[k].x = (T)std::cos(a);
    t[k].y = (T)std::sin(a);
  }
  return t;
}

// per-pass Stockham twiddles, layout of fft.cuh's FftPlan: pass p >= 1 holds [(r-1)*Ns + k] =
// exp(-2*pi*i*r*k/(Ns*R)), computed directly in float64
template <class T2, class T>
inline std::vector<T2> make_pass_twiddles(int M) {
  FftPlan plan = make_plan(M);
  std::vector<T2> t(plan.tw_total > 0 ? plan.tw_total : 1);
  for (int p = 1; p < plan.n_pass; ++p) {
    const int R = plan.radix[p], Ns = plan.ns[p];
    for (int r = 1; r < R; ++r)
      for (int k = 0; k < Ns; ++k) {
        double a = -2.0 * kPi * (double)r * (double)k / ((double)Ns * (double)R);
        t[plan.tw_off[p] + (r - 1) * Ns + k].x = (T)std::cos(a);
        t[plan.tw_off[p] + (r - 1) * Ns + k].y = (T)std::sin(a);
      }
  }
  return t;
}

// ---------------------------------------------------------------------------------- numpy sums
// np.add.reduce on a contiguous float64 vector (pairwise, 8 accumulators, block 128)
inline double np_pairwise_sum(const double* a, long n) {
  if (n < 8) {
    double res = 0.;
    for (long i = 0; i < n; ++i) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    long i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  } else {
    long n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
  }
}

// ---------------------------------------------------------------------------------- pYIN tables
struct PyinTables {
  std::vector<double> thresholds;   // [n_thr+1]  np.linspace(0, 1, n_thr+1)
  std::vector<double> beta_probs;   // [n_thr]    diff of the beta(a,b) cdf at the thresholds
  std::vector<double> beta_cum;     // [n_thr+1]  np.sum(beta_probs[:k])
  std::vector<double> boltz_exp;    // [kmax+1]   exp(-lambda*k)
  std::vector<double> boltz_fact;   // [kmax+1]   (1-exp(-lambda))/(1-exp(-lambda*N)), [0] unused
  std::vector<double> freqs;        // [npb]      fmin * 2^(b/(12*nbps))
  // banded log-transition rows: lt[(row*tw + d)*2 + {0:same voicing, 1:switch}] for source bin
  // with row id `row`, destination offset d-hw
  std::vector<double> lt_rows;
  std::vector<uint16_t> row_id;     // [npb]
  int n_rows = 0;
  double lt0 = 0;                   // log(0 + tiny): every out-of-band transition
  double lt_max = 0;                // largest entry of lt_rows
  int flat_ok = 0;                  // centre entry of the uniform row is its strict maximum by > 1e-3 (k_viterbi.cuh rule 10)
  double twin_gap = -1e300;         // smallest (same - switch) entry difference minus a rounding margin (k_viterbi.cuh rule 8); -1e300: rule off
  // interior rows (source bins hw .. npb-1-hw) differ only by the rounding of the row sum (a last-place
  // unit in a few entries).  fl(V + lt) does not see that difference once |V| >= 2^uniform_emin (see
  // make_uniform_row): then every interior source may use `lt_uniform` (the same-voicing entries of one
  // representative row), which the kernel reads as an immediate constant operand.
  std::vector<double> lt_uniform;   // [tw]; empty when tw != the fast path's width or npb too small
  double uniform_vmax = -1e308;     // the uniform row is valid while vmax <= uniform_vmax
  double li_voiced = 0, li_unvoiced = 0;  // log(p_init + tiny)
};

// regularised incomplete beta for integer a,b: I_x(a,b) = sum_{j=a}^{a+b-1} C(a+b-1, j) x^j (1-x)^(a+b-1-j)
inline double beta_cdf_int(double x, int a, int b) {
  if (x <= 0) return 0.0;
  if (x >= 1) return 1.0;
  const int n = a + b - 1;
  // sum the complementary (shorter, better conditioned for small x) side: 1 - sum_{j<a}
  double s = 0;
  double lx = std::log(x), l1x = std::log1p(-x);
  for (int j = 0; j < a; ++j) {
    double lc = std::lgamma(n + 1.0) - std::lgamma(j + 1.0) - std::lgamma(n - j + 1.0);
    s += std::exp(lc + j * lx + (n - j) * l1x);
  }
  return 1.0 - s;
}

inline std::vector<double> triangle_window(int width) {  // scipy.signal.windows.triang(sym=True)
  std::vector<double> w(width);
  if (width % 2 == 0) {
    for (int i = 0; i < width / 2; ++i) { w[i] = (2.0 * (i + 1) - 1.0) / width; w[width - 1 - i] = w[i]; }
  } else {
    for (int i = 0; i < (width + 1) / 2; ++i) { w[i] = 2.0 * (i + 1) / (width + 1.0); w[width - 1 - i] = w[i]; }
  }
  return w;
}

// fl(V + a) == fl(V + b) for EVERY double V with |V| >= 2^e (a, b < 0 much smaller in magnitude) iff
// no multiple of 2^(e-53) lies in the closed interval [min(a,b), max(a,b)]: V is a multiple of its own
// ulp u >= 2^(e-52); the sum rounds to the grid u (same binade: boundaries at odd multiples of u/2) or
// 2u (next binade: boundaries at odd multiples of u), i.e. the decision boundaries, shifted by -V, are
// multiples of u/2 >= 2^(e-53); coarser grids are subsets of finer ones, so the finest e decides.
inline int uniform_min_exponent(double a, double b) {
  if (a == b) return -1074;
  const double lo = a < b ? a : b, hi = a < b ? b : a;
  for (int e = 1; e < 1000; ++e) {
    const double sl = std::ldexp(lo, 53 - e), sh = std::ldexp(hi, 53 - e);   // exact scalings
    if (std::ceil(sl) > std::floor(sh)) return e;
  }
  return 1000;
}

inline void make_uniform_row(PyinTables& t, const Geometry& g) {
  const int n = g.npb, tw = g.tw, hw = g.hw;
  if (n < 4 * hw + 2) return;
  const int base = t.row_id[n / 2];
  int emin = -1074;
  for (int i = hw; i <= n - 1 - hw; ++i) {
    const int r = t.row_id[i];
    if (r == base) continue;
    for (int d = 0; d < tw; ++d) {
      const int e = uniform_min_exponent(t.lt_rows[((size_t)base * tw + d) * 2], t.lt_rows[((size_t)r * tw + d) * 2]);
      if (e > emin) emin = e;
    }
  }
  if (emin >= 1000) return;
  t.lt_uniform.resize(tw);
  for (int d = 0; d < tw; ++d) t.lt_uniform[d] = t.lt_rows[((size_t)base * tw + d) * 2];
  t.uniform_vmax = emin <= -1074 ? 0.0 : -std::ldexp(1.0, emin);
  t.flat_ok = 1;
  for (int d = 0; d < tw; ++d)
    if (d != hw && !(t.lt_uniform[hw] - t.lt_uniform[d] > 1e-3)) t.flat_ok = 0;
}

inline PyinTables make_pyin_tables(const roar_sup_config& c, const Geometry& g) {
  PyinTables t;
  const int nt = g.n_thr;
  t.thresholds.resize(nt + 1);
  const double step = 1.0 / nt;
  for (int i = 0; i <= nt; ++i) t.thresholds[i] = i == nt ? 1.0 : i * step;
  std::vector<double> cdf(nt + 1);
  const bool int_ab = c.beta_a == std::floor(c.beta_a) && c.beta_b == std::floor(c.beta_b) && c.beta_a >= 1 && c.beta_b >= 1;
  for (int i = 0; i <= nt; ++i) {
    if (int_ab) {
      cdf[i] = beta_cdf_int(t.thresholds[i], (int)c.beta_a, (int)c.beta_b);
    } else {
      // midpoint-rule quadrature of the beta density (non-default parameters only)
      const int steps = 20000;
      double acc = 0, x1 = t.thresholds[i];
      for (int s = 0; s < steps; ++s) {
        double x = (s + 0.5) * x1 / steps;
        acc += std::pow(x, c.beta_a - 1) * std::pow(1 - x, c.beta_b - 1);
      }
      double B = std::exp(std::lgamma(c.beta_a) + std::lgamma(c.beta_b) - std::lgamma(c.beta_a + c.beta_b));
      cdf[i] = acc * x1 / steps / B;
    }
  }
  t.beta_probs.resize(nt);
  for (int i = 0; i < nt; ++i) t.beta_probs[i] = cdf[i + 1] - cdf[i];
  t.beta_cum.resize(nt + 1);
  for (int k = 0; k <= nt; ++k) t.beta_cum[k] = np_pairwise_sum(t.beta_probs.data(), k);
  const double lam = c.boltzmann_parameter;
  t.boltz_exp.resize(g.kmax + 1);
  t.boltz_fact.resize(g.kmax + 1);
  for (int k = 0; k <= g.kmax; ++k) {
    t.boltz_exp[k] = std::exp(-lam * k);
    t.boltz_fact[k] = k == 0 ? 0.0 : (1 - std::exp(-lam)) / (1 - std::exp(-lam * k));
  }
  t.freqs.resize(g.npb);
  for (int b = 0; b < g.npb; ++b) t.freqs[b] = c.pitch_fmin * std::pow(2.0, (double)b / (12.0 * g.nbps));

  // ---- transition_local(npb, tw, "triangle", wrap=False), row-normalised with numpy's pairwise sum
  const int n = g.npb, tw = g.tw, hw = g.hw;
  std::vector<double> win = triangle_window(tw);
  std::vector<double> row(n);
  const double p_same = 1.0 - c.switch_prob;          // transition_loop(2, 1 - switch_prob)
  const double p_switch = (1.0 - p_same) / (2 - 1);
  t.lt0 = std::log(0.0 + kTiny64);
  std::map<std::string, int> seen;
  t.row_id.resize(n);
  std::vector<double> banded(2 * tw);
  for (int i = 0; i < n; ++i) {
    std::fill(row.begin(), row.end(), 0.0);
    for (int d = 0; d < tw; ++d) {
      int j = i + d - hw;
      if (j >= 0 && j < n) row[j] = win[d];
    }
    const double s = np_pairwise_sum(row.data(), n);
    for (int d = 0; d < tw; ++d) {
      int j = i + d - hw;
      if (j >= 0 && j < n) {
        double p = row[j] / s;
        banded[2 * d + 0] = std::log(p_same * p + kTiny64);
        banded[2 * d + 1] = std::log(p_switch * p + kTiny64);
      } else {
        banded[2 * d + 0] = t.lt0;
        banded[2 * d + 1] = t.lt0;
      }
    }
    std::string key((const char*)banded.data(), banded.size() * sizeof(double));
    auto it = seen.find(key);
    if (it == seen.end()) {
      int id = (int)seen.size();
      seen.emplace(key, id);
      t.lt_rows.insert(t.lt_rows.end(), banded.begin(), banded.end());
      t.row_id[i] = (uint16_t)id;
    } else {
      t.row_id[i] = (uint16_t)it->second;
    }
  }
  t.n_rows = (int)seen.size();
  t.lt_max = t.lt0;
  for (double v : t.lt_rows) if (v > t.lt_max) t.lt_max = v;
  {   // (lt0, lt0) marks a destination outside the bin range: never offered
    double gap = 1e300;
    for (size_t i = 0; i + 1 < t.lt_rows.size(); i += 2)
      if (t.lt_rows[i] > t.lt0) gap = std::min(gap, t.lt_rows[i] - t.lt_rows[i + 1]);
    t.twin_gap = (gap > 1.0 && gap < 1e299) ? gap - 1e-3 : -1e300;
  }
  t.li_voiced = std::log(0.0 + kTiny64);
  t.li_unvoiced = std::log(1.0 / n + kTiny64);
  make_uniform_row(t, g);
  return t;
}

// dense [2npb, 2npb] log-transition matrix from the banded rows (tests only)
inline void dense_log_transition(const PyinTables& t, const Geometry& g, double* out) {
  const int n = g.npb, S = 2 * n;
  for (long i = 0; i < (long)S * S; ++i) out[i] = t.lt0;
  for (int bs = 0; bs < 2; ++bs)
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < g.tw; ++d) {
        int j = i + d - g.hw;
        if (j < 0 || j >= n) continue;
        const double* e = &t.lt_rows[((size_t)t.row_id[i] * g.tw + d) * 2];
        for (int bd = 0; bd < 2; ++bd)
          out[(size_t)(bs * n + i) * S + (bd * n + j)] = e[bs == bd ? 0 : 1];
      }
}

}  // namespace roar
