// libroar_sup.so -- C ABI (include/roar_sup.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/roar_sup.h"
#include "common.cuh"
#include "fft.cuh"
#include "host_io.hpp"
#include "k_misc.cuh"
#include "k_pyin_front.cuh"
#include "k_stft_mel.cuh"
#include "k_stft_bwd.cuh"
#include "k_viterbi.cuh"
#include "tables.hpp"

using namespace roar;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ROAR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));         \
  } while (0)

struct roar_sup_handle {
  roar_sup_config cfg;
  Geometry g;
  int device = 0;
  int has_pyin = 1;            // 0: cfg.pyin_frame_length == 0 (mel-only handle, no pYIN tables)
  int sm_count = 148;
  size_t max_smem = 0;
  std::vector<void*> allocs;
  // K1
  float* d_window = nullptr; cf32* d_tw = nullptr; cf32* d_tw_post = nullptr;
  int32_t *d_mel_start = nullptr, *d_mel_count = nullptr, *d_mel_offset = nullptr;
  float* d_mel_w = nullptr; int mel_nw = 0;
  int stft_FT = 16, stft_P = 64, stft_G = 4, stft_span = 0; size_t stft_smem = 0;
  int use_tma = 1;
  // K2
  double *d_thr = nullptr, *d_beta = nullptr, *d_beta_cum = nullptr, *d_bexp = nullptr, *d_bfact = nullptr;
  int pyin_FT = 15, pyin_BL = 0, pyin_nb = 0, pyin_ngroups = 0, pyin_ylen = 0; size_t cmnd_smem = 0, prob_smem = 0, energy_smem = 0;
  // K3
  double* d_lt_rows = nullptr; uint16_t* d_row_id = nullptr; double* d_freqs = nullptr;
  int n_rows = 0; double lt0 = 0, lt_max = 0, li_v = 0, li_u = 0;
  double uniform_vmax = -1e308; double ltu[VIT_TW] = {0}; double twin_gap = 0; int flat_ok = 0;
  int vit_threads = 0; size_t vit_smem = 0; int lt_in_smem = 1;
  int vit_fast = 0; size_t vit3_smem = 0;
  // K4
  double* d_lf = nullptr; int lf_n = 0;
  // optional per-kernel timing (diagnostics; not thread-safe)
  int profiling = 0;
  struct Pending { int id; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  double prof_ms[ROAR_K_COUNT] = {0};
  int64_t prof_n[ROAR_K_COUNT] = {0};
};

static cudaEvent_t prof_event(roar_sup_handle* h) {
  if (!h->pool.empty()) { cudaEvent_t e = h->pool.back(); h->pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
// LAUNCH(h, id, stream, kernel<<<...>>>(...)) -- brackets the launch with events when profiling is on
#define LAUNCH(h, kid, st, ...)                                             \
  do {                                                                     \
    if ((h)->profiling) {                                                  \
      roar_sup_handle::Pending pe_; pe_.id = (kid);                        \
      pe_.a = prof_event(h); pe_.b = prof_event(h);                        \
      cudaEventRecord(pe_.a, st);                                          \
      __VA_ARGS__;                                                         \
      cudaEventRecord(pe_.b, st);                                          \
      (h)->pending.push_back(pe_);                                         \
    } else {                                                               \
      __VA_ARGS__;                                                         \
    }                                                                      \
  } while (0)

template <class T> static int upload(roar_sup_handle* h, const std::vector<T>& v, T** out) {
  void* p = nullptr;
  size_t bytes = v.size() * sizeof(T);
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(ROAR_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  h->allocs.push_back(p);
  if (!v.empty()) {
    e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(ROAR_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
  }
  *out = (T*)p;
  return 0;
}

static void launch_stft(const roar_sup_handle* h, unsigned grid, cudaStream_t st, const StftParams& p);
static StftParams stft_params_base(const roar_sup_handle* h) {
  StftParams p;
  memset(&p, 0, sizeof(p));
  const Geometry& g = h->g;
  p.n_fft = g.n_fft; p.hop = g.hop; p.M = g.M; p.n_bins = g.n_bins; p.n_mels = g.n_mels;
  p.FT = h->stft_FT; p.span = h->stft_span; p.P = h->stft_P; p.G = h->stft_G;
  p.pad_left = h->cfg.exact_pad ? (g.n_fft - g.hop) / 2 : g.n_fft / 2;
  p.floor_ = (float)h->cfg.spec_floor; p.mag_power = (float)h->cfg.mag_power;
  p.log_guard = (float)h->cfg.log_guard; p.preemph = (float)h->cfg.preemph;
  p.log_mode = h->cfg.log_mode; p.has_preemph = h->cfg.has_preemph; p.use_tma = h->use_tma;
  p.preemph_after_pad = h->cfg.exact_pad; p.energy_mode = h->cfg.energy_mode;
  p.window = h->d_window; p.tw = h->d_tw; p.tw_post = h->d_tw_post;
  p.mel_start = h->d_mel_start; p.mel_count = h->d_mel_count; p.mel_offset = h->d_mel_offset;
  p.mel_w = h->d_mel_w; p.mel_nw = h->mel_nw;
  p.tw_total = fft_tw_total(g.M);
  p.part_per_slot = h->stft_P >= 32 && h->stft_P % 32 == 0 ? h->stft_P / 32 : 0;
  return p;
}

static PyinParams pyin_params_base(const roar_sup_handle* h) {
  PyinParams p;
  memset(&p, 0, sizeof(p));
  const Geometry& g = h->g;
  p.F = g.pf; p.W = g.pw; p.hop = g.ph;
  p.min_period = g.min_period; p.max_period = g.max_period; p.n_lags = g.n_lags;
  p.FT = h->pyin_FT; p.BL = h->pyin_BL; p.nb = h->pyin_nb; p.n_groups = h->pyin_ngroups; p.ylen = h->pyin_ylen;
  p.npb = g.npb; p.nbps = g.nbps; p.kmax = g.kmax; p.n_thr = g.n_thr;
  p.sr = h->cfg.sample_rate; p.fmin = h->cfg.pitch_fmin; p.no_trough_prob = h->cfg.no_trough_prob;
  p.thresholds = h->d_thr; p.beta_probs = h->d_beta; p.beta_cum = h->d_beta_cum;
  p.boltz_exp = h->d_bexp; p.boltz_fact = h->d_bfact;
  return p;
}

static void launch_stft_bwd(unsigned grid, size_t smem, cudaStream_t st, const StftBwdParams& q) {
  switch (q.f.M) {
    case 32: k_stft_mel_bwd<5><<<grid, 256, smem, st>>>(q); break;
    case 64: k_stft_mel_bwd<6><<<grid, 256, smem, st>>>(q); break;
    case 128: k_stft_mel_bwd<7><<<grid, 256, smem, st>>>(q); break;
    case 256: k_stft_mel_bwd<8><<<grid, 256, smem, st>>>(q); break;
    case 512: k_stft_mel_bwd<9><<<grid, 256, smem, st>>>(q); break;
    case 1024: k_stft_mel_bwd<10><<<grid, 256, smem, st>>>(q); break;
    default: k_stft_mel_bwd<11><<<grid, 256, smem, st>>>(q); break;
  }
}

static void launch_stft(const roar_sup_handle* h, unsigned grid, cudaStream_t st, const StftParams& p) {
  switch (p.M) {
    case 32: k_stft_mel<5><<<grid, 256, h->stft_smem, st>>>(p); break;
    case 64: k_stft_mel<6><<<grid, 256, h->stft_smem, st>>>(p); break;
    case 128: k_stft_mel<7><<<grid, 256, h->stft_smem, st>>>(p); break;
    case 256: k_stft_mel<8><<<grid, 256, h->stft_smem, st>>>(p); break;
    case 512: k_stft_mel<9><<<grid, 256, h->stft_smem, st>>>(p); break;
    case 1024: k_stft_mel<10><<<grid, 256, h->stft_smem, st>>>(p); break;
    default: k_stft_mel<11><<<grid, 256, h->stft_smem, st>>>(p); break;
  }
}

extern "C" {

int roar_sup_abi_version(void) { return ROAR_SUP_ABI_VERSION; }
const char* roar_sup_last_error(void) { return g_err.c_str(); }

void roar_sup_config_default(roar_sup_config* c) {
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(*c);
  c->sample_rate = 22050; c->n_fft = 1024; c->win_length = 1024; c->hop_length = 256;
  c->window = ROAR_WIN_HANN; c->n_mels = 80; c->mel_norm = 1; c->fmin = 0.0; c->fmax = 8000.0;
  c->spec_floor = 1e-9; c->mag_power = 1.0; c->log_mode = ROAR_LOG_CLAMP; c->exact_pad = 0;
  c->log_guard = 1.17549435e-38; c->has_preemph = 0; c->normalize = ROAR_NORM_NONE; c->preemph = 0.97;
  c->pad_value = 0.0; c->pad_to = 0;
  c->pitch_fmin = 65.40639132514966; c->pitch_fmax = 2093.004522404789;
  c->pyin_frame_length = 1024; c->pyin_win_length = 0; c->pyin_hop_length = 0; c->n_thresholds = 100;
  c->beta_a = 2; c->beta_b = 18; c->boltzmann_parameter = 2; c->resolution = 0.1;
  c->max_transition_rate = 35.92; c->switch_prob = 0.01; c->no_trough_prob = 0.01;
}

int roar_sup_host_mel_filterbank(const roar_sup_config* cfg, float* out) {
  std::string v = validate(*cfg);
  if (!v.empty()) return fail(ROAR_ERR_INVALID_ARG, v);
  std::vector<float> fb = make_mel_filterbank(*cfg);
  memcpy(out, fb.data(), fb.size() * sizeof(float));
  return 0;
}
int roar_sup_host_window(const roar_sup_config* cfg, float* out) {
  std::string v = validate(*cfg);
  if (!v.empty()) return fail(ROAR_ERR_INVALID_ARG, v);
  std::vector<float> w = make_window(*cfg);
  memcpy(out, w.data(), w.size() * sizeof(float));
  return 0;
}
int roar_sup_host_pyin_log_transition(const roar_sup_config* cfg, double* out, int64_t n) {
  std::string v = validate(*cfg);
  if (!v.empty()) return fail(ROAR_ERR_INVALID_ARG, v);
  if (cfg->pyin_frame_length == 0) return fail(ROAR_ERR_INVALID_ARG, "this configuration has no pYIN (pyin_frame_length == 0)");
  Geometry g = geometry(*cfg);
  if (n != (int64_t)4 * g.npb * g.npb) return fail(ROAR_ERR_INVALID_ARG, "log-transition buffer must hold (2*npb)^2 doubles");
  PyinTables t = make_pyin_tables(*cfg, g);
  dense_log_transition(t, g, out);
  return 0;
}
int roar_sup_host_pyin_beta_probs(const roar_sup_config* cfg, double* out) {
  std::string v = validate(*cfg);
  if (!v.empty()) return fail(ROAR_ERR_INVALID_ARG, v);
  if (cfg->pyin_frame_length == 0) return fail(ROAR_ERR_INVALID_ARG, "this configuration has no pYIN (pyin_frame_length == 0)");
  Geometry g = geometry(*cfg);
  PyinTables t = make_pyin_tables(*cfg, g);
  memcpy(out, t.beta_probs.data(), t.beta_probs.size() * sizeof(double));
  return 0;
}

// Builds the tables of `h` (already allocated).  Every failure returns to roar_sup_create, which destroys
// the half-built handle and restores the caller's current device.
static int create_impl(roar_sup_handle* h, const roar_sup_config* cfg, int device, const cudaDeviceProp& prop) {
  h->cfg = *cfg; h->g = geometry(*cfg); h->device = device;
  h->has_pyin = cfg->pyin_frame_length > 0 ? 1 : 0;
  h->sm_count = prop.multiProcessorCount;
  // The opt-in limit is a per-FUNCTION attribute shared by every handle in the process, so each kernel is
  // opened up to the device maximum once; a launch still asks only for what its own geometry needs.
  h->max_smem = prop.sharedMemPerBlockOptin;
  const Geometry& g = h->g;
  if (h->has_pyin) {
    if (g.pf < 256) return fail(ROAR_ERR_UNSUPPORTED, "pyin_frame_length < 256 is not supported");
    if (g.ph % 4 != 0) return fail(ROAR_ERR_UNSUPPORTED, "pyin hop length must be a multiple of 4");
    if (g.npb > 1024 || g.npb < 2 * g.hw + 2) return fail(ROAR_ERR_UNSUPPORTED, "pitch-bin count out of range");
    if (g.min_period < 1 || g.n_lags < 3 || g.max_period >= g.pf - g.pw) return fail(ROAR_ERR_INVALID_ARG, "pyin period range empty");
  }
  const char* env_tma = getenv("ROAR_SUP_NO_TMA");
  h->use_tma = (env_tma && env_tma[0] == '1') ? 0 : 1;
  int rc = 0;
#define UP(vec, field) if ((rc = upload(h, vec, &h->field)) != 0) return rc;
  // ---- K1 tables
  {
    std::vector<float> win = make_window(*cfg);
    std::vector<float> fb = make_mel_filterbank(*cfg);
    MelRows mr = make_mel_rows(fb, g.n_mels, g.n_bins);
    std::vector<cf32> tw = make_pass_twiddles<cf32, float>(g.M);
    std::vector<cf32> twp = make_twiddles<cf32, float>(g.n_fft, g.M + 1);
    UP(win, d_window) UP(tw, d_tw) UP(twp, d_tw_post)
    UP(mr.start, d_mel_start) UP(mr.count, d_mel_count) UP(mr.offset, d_mel_offset) UP(mr.weights, d_mel_w)
    h->mel_nw = (int)mr.weights.size();
    h->stft_P = g.M / 8 < 1 ? 1 : g.M / 8;
    h->stft_G = 256 / h->stft_P < 1 ? 1 : 256 / h->stft_P;
    h->stft_FT = g.n_fft <= 1024 ? 16 : 8;
    if (h->stft_FT < h->stft_G) h->stft_FT = h->stft_G;
    h->stft_span = (h->stft_FT - 1) * g.hop + g.n_fft;
    StftParams sp = stft_params_base(h);
    h->stft_smem = stft_smem_carve(sp, 256, nullptr, nullptr);
    if (h->stft_smem > h->max_smem) return fail(ROAR_ERR_UNSUPPORTED, "STFT tile does not fit in shared memory");
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel_bwd<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel_bwd<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel_bwd<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel_bwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel_bwd<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel_bwd<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_stft_mel_bwd<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
  }
  CUDA_TRY(cudaFuncSetAttribute(k_trim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem - 2048));   // it also has static arrays
  // ---- K2 / K3 tables
  if (h->has_pyin) {
    PyinTables t = make_pyin_tables(*cfg, g);
    std::vector<double> lt_up = t.lt_rows;              // + one all-zero row for the fast Viterbi's sentinel sources
    lt_up.resize(lt_up.size() + 2 * (size_t)g.tw, 0.0);
    UP(t.thresholds, d_thr) UP(t.beta_probs, d_beta) UP(t.beta_cum, d_beta_cum)
    UP(t.boltz_exp, d_bexp) UP(t.boltz_fact, d_bfact) UP(lt_up, d_lt_rows) UP(t.row_id, d_row_id)
    UP(t.freqs, d_freqs)
    h->n_rows = t.n_rows; h->lt0 = t.lt0; h->lt_max = t.lt_max; h->li_v = t.li_voiced; h->li_u = t.li_unvoiced;
    h->twin_gap = t.twin_gap;
    { const char* env_f = getenv("ROAR_SUP_NO_FLAT"); h->flat_ok = (env_f && env_f[0] == '1') ? 0 : t.flat_ok; }
    if (g.tw == VIT_TW && (int)t.lt_uniform.size() == VIT_TW) {
      h->uniform_vmax = t.uniform_vmax;
      for (int d = 0; d < VIT_TW; ++d) h->ltu[d] = t.lt_uniform[d];
    }
    // K2a tiling: frames per tile chosen for (i) two CTAs per SM, (ii) little block overhead
    // (a tile of ft frames computes ft + nb - 1 blocks), (iii) a chunk count that fills the 8 warps
    cmnd_blocking(g.pw, g.ph, &h->pyin_BL, &h->pyin_nb);
    h->pyin_ngroups = (g.max_period + 1 + ACF_R - 1) / ACF_R;
    {
      const size_t budget = (h->max_smem + 1024) / 2 - 1024;   // per-SM shared memory / 2, minus the per-CTA reserve
      int best = 0; double best_score = -1.0;
      for (int pass = 0; pass < 2 && best == 0; ++pass) {
        for (int ft = 1; ft <= 32; ++ft) {
          h->pyin_FT = ft;
          h->pyin_ylen = cmnd_ylen(ft, g.pf, g.ph, h->pyin_BL, h->pyin_nb, h->pyin_ngroups);
          PyinParams pp = pyin_params_base(h);
          const size_t need = cmnd_smem_carve(pp, nullptr, nullptr);
          if (need > (pass == 0 ? budget : h->max_smem)) break;
          const int chunks = (ft + h->pyin_nb - 1) * cmnd_cpb(pp);
          const double score = (double)ft / (ft + h->pyin_nb - 1) * chunks / (8.0 * ((chunks + 7) / 8));
          if (score > best_score) { best_score = score; best = ft; }
        }
      }
      if (best == 0) return fail(ROAR_ERR_UNSUPPORTED, "pYIN tile does not fit in shared memory");
      h->pyin_FT = best;
      h->pyin_ylen = cmnd_ylen(best, g.pf, g.ph, h->pyin_BL, h->pyin_nb, h->pyin_ngroups);
    }
    PyinParams pp = pyin_params_base(h);
    h->cmnd_smem = cmnd_smem_carve(pp, nullptr, nullptr);
    h->prob_smem = prob_smem_carve(pp, nullptr, nullptr) * 8 + sizeof(double) * (2 * (g.n_thr + 2) + 2 * (g.kmax + 2));
    if (h->cmnd_smem > h->max_smem || h->prob_smem > h->max_smem) return fail(ROAR_ERR_UNSUPPORTED, "pYIN tile does not fit in shared memory");
    CUDA_TRY(cudaFuncSetAttribute(k_pyin_cmnd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    h->energy_smem = sizeof(float) * (size_t)(epad(energy_span(pp), g.ph) + 4);
    if (h->energy_smem > h->max_smem) return fail(ROAR_ERR_UNSUPPORTED, "pYIN energy tile does not fit in shared memory");
    CUDA_TRY(cudaFuncSetAttribute(k_pyin_energy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_pyin_probs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    h->vit_threads = (g.npb + 31) / 32 * 32;
    if (h->vit_threads < g.kmax) return fail(ROAR_ERR_UNSUPPORTED, "kmax exceeds Viterbi block size");
    size_t base = sizeof(cf64) * 2 * g.npb + sizeof(double) * 2 * g.npb + sizeof(double) * 64 + sizeof(int) * 64 +
                  ((sizeof(int32_t) * g.npb + 15) & ~(size_t)15) + 128;
    size_t ltb = sizeof(double) * 2 * (size_t)t.n_rows * g.tw;
    h->lt_in_smem = base + ltb <= h->max_smem ? 1 : 0;
    h->vit_smem = base + (h->lt_in_smem ? ltb : 0);
    CUDA_TRY(cudaFuncSetAttribute(k_pyin_viterbi<true, 640>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_pyin_viterbi<false, 640>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_pyin_viterbi<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(k_pyin_viterbi<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
    // fast path: the reference geometry (transition width 51), row ids in 6 bits
    h->vit3_smem = sizeof(Vit3Shared);
    // ROAR_SUP_VITERBI=generic forces the any-geometry kernel (tests compare the two)
    const char* env_v = getenv("ROAR_SUP_VITERBI");
    h->vit_fast = (g.tw == VIT_TW && t.n_rows + 1 <= VIT_ROWS_MAX && g.npb <= VIT_NPB_MAX && g.kmax <= VIT_KMAX_MAX &&
                   h->vit3_smem <= h->max_smem && !(env_v && env_v[0] == 'g')) ? 1 : 0;
    if (h->vit_fast)
      CUDA_TRY(cudaFuncSetAttribute(k_pyin_viterbi51, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->max_smem));
  }
  // ---- K4 log-factorial table
  {
    h->lf_n = 1 << 16;
    std::vector<double> lf(h->lf_n);
    for (int i = 0; i < h->lf_n; ++i) lf[i] = std::lgamma((double)i + 1.0);
    UP(lf, d_lf)
  }
#undef UP
  return 0;
}

int roar_sup_create(const roar_sup_config* cfg, int device, roar_sup_handle** out) {
  if (!cfg || !out) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  std::string v = validate(*cfg);
  if (!v.empty()) return fail(ROAR_ERR_INVALID_ARG, v);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(ROAR_ERR_NO_DEVICE, "no CUDA device: libroar_sup has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(ROAR_ERR_INVALID_ARG, "bad device index");
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  cudaDeviceProp prop;
  int rc = 0;
  roar_sup_handle* h = nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    rc = fail(ROAR_ERR_CUDA, "cannot select the CUDA device");
  } else if (prop.major < 10) {
    rc = fail(ROAR_ERR_UNSUPPORTED, "libroar_sup is built for sm_100a (Blackwell) only");
  } else {
    h = new roar_sup_handle();
    h->device = device;
    rc = create_impl(h, cfg, device, prop);
    if (rc != 0) {
      const std::string keep = g_err;     // roar_sup_destroy does not touch it, but stay explicit
      roar_sup_destroy(h);
      g_err = keep;
      h = nullptr;
    }
  }
  if (prev_device >= 0 && prev_device != device) cudaSetDevice(prev_device);   // leave the caller's device as it was
  *out = h;
  return rc;
}

int roar_sup_debug_counters(uint64_t* out, int32_t n, int reset) {
  if (!out || n <= 0) return fail(ROAR_ERR_INVALID_ARG, "bad argument");
  unsigned long long host[VIT_STATS_N] = {0};
#ifdef ROAR_VIT_STATS
  CUDA_TRY(cudaMemcpyFromSymbol(host, g_vit_stats, sizeof(host)));
  if (reset) { unsigned long long z[VIT_STATS_N] = {0}; CUDA_TRY(cudaMemcpyToSymbol(g_vit_stats, z, sizeof(z))); }
#else
  (void)reset;
#endif
  for (int i = 0; i < n; ++i) out[i] = i < VIT_STATS_N ? host[i] : 0;
  return 0;
}

int roar_sup_set_profiling(roar_sup_handle* h, int on) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  h->profiling = on ? 1 : 0;
  return 0;
}
int roar_sup_profile_read(roar_sup_handle* h, double* ms_out, int64_t* count_out, int reset) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  for (auto& pe : h->pending) {
    cudaEventSynchronize(pe.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) { h->prof_ms[pe.id] += ms; h->prof_n[pe.id] += 1; }
    h->pool.push_back(pe.a); h->pool.push_back(pe.b);
  }
  h->pending.clear();
  for (int i = 0; i < ROAR_K_COUNT; ++i) {
    if (ms_out) ms_out[i] = h->prof_ms[i];
    if (count_out) count_out[i] = h->prof_n[i];
    if (reset) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  }
  return 0;
}

void roar_sup_destroy(roar_sup_handle* h) {
  if (!h) return;
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  cudaSetDevice(h->device);
  for (auto& pe : h->pending) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
  for (auto e : h->pool) cudaEventDestroy(e);
  for (void* p : h->allocs) cudaFree(p);
  if (prev_device >= 0 && prev_device != h->device) cudaSetDevice(prev_device);
  delete h;
}

int64_t roar_sup_num_frames(const roar_sup_handle* h, int64_t L) {
  if (h->cfg.exact_pad) {
    const int64_t p = (h->g.n_fft - h->g.hop) / 2;
    return (L + 2 * p - h->g.n_fft) / h->g.hop + 1;
  }
  return 1 + L / h->g.hop;
}
int64_t roar_sup_pyin_num_frames(const roar_sup_handle* h, int64_t L) { return h->has_pyin ? 1 + L / h->g.ph : 0; }

int roar_sup_pyin_geometry(const roar_sup_handle* h, int32_t out8[8]) {
  const Geometry& g = h->g;
  out8[0] = g.min_period; out8[1] = g.max_period; out8[2] = g.npb; out8[3] = g.tw;
  out8[4] = g.ph; out8[5] = g.pw; out8[6] = g.kmax; out8[7] = h->n_rows;
  return 0;
}

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

struct PyinWs {
  int32_t *tile_off, *etile_off, *order, *hist, *last_state, *n_cand, *tile_map;
  double *big, *cand_lp, *lp_unv;
  float* energy;
  uint16_t* cand_bin;
  size_t total;
};
static PyinWs pyin_ws_layout(const roar_sup_handle* h, int32_t n_utts, int64_t frames, int32_t max_T, unsigned char* base) {
  const Geometry& g = h->g;
  PyinWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) { unsigned char* p = base ? base + o : nullptr; o += a256(bytes); return p; };
  w.tile_off = (int32_t*)take(sizeof(int32_t) * (n_utts + 1));
  w.order = (int32_t*)take(sizeof(int32_t) * (n_utts + 1));
  w.etile_off = (int32_t*)take(sizeof(int32_t) * (n_utts + 1));
  w.hist = (int32_t*)take(sizeof(int32_t) * ((size_t)max_T + 2));
  w.last_state = (int32_t*)take(sizeof(int32_t) * (n_utts + 1));
  w.n_cand = (int32_t*)take(sizeof(int32_t) * (frames + 1));
  w.tile_map = (int32_t*)take(sizeof(int32_t) * (size_t)(frames / h->pyin_FT + n_utts + 1));
  w.lp_unv = (double*)take(sizeof(double) * (frames + 1));
  w.cand_lp = (double*)take(sizeof(double) * (size_t)frames * g.kmax);
  w.cand_bin = (uint16_t*)take(sizeof(uint16_t) * (size_t)frames * g.kmax);
  w.energy = (float*)take(sizeof(float) * (size_t)frames * (g.max_period + 1));
  size_t a = sizeof(double) * (size_t)frames * g.n_lags;
  size_t b = sizeof(uint16_t) * (size_t)frames * 2 * g.npb;
  w.big = (double*)take(a > b ? a : b);
  w.total = o;
  return w;
}

// roar_fbank_forward / roar_fbank_backward: three int64 and two int32 arrays of B + 1 entries
static size_t fbank_ws_bytes(int32_t B) {
  const size_t n = (size_t)(B > 0 ? B : 0) + 1;
  return 3 * a256(sizeof(int64_t) * n) + 2 * a256(sizeof(int32_t) * n) + 256;
}

size_t roar_fbank_workspace_bytes(const roar_sup_handle* h, int32_t B) {
  (void)h;
  return fbank_ws_bytes(B);
}

size_t roar_sup_workspace_bytes(const roar_sup_handle* h, int32_t n_utts, int64_t total_samples,
                                int64_t total_pyin_frames) {
  (void)total_samples;
  // the maximum over every entry point's layout, so one buffer of this size serves them all
  size_t need = fbank_ws_bytes(n_utts);
  const size_t lm = 2 * a256(sizeof(int32_t) * ((size_t)(n_utts > 0 ? n_utts : 0) + 1)) + 256;   // roar_sup_logmel_energy: tile prefix sums
  if (lm > need) need = lm;
  if (h->has_pyin) {
    // max_T is bounded by the frame total; the histogram is sized for the worst case
    int64_t max_T = total_pyin_frames < (1 << 20) ? total_pyin_frames : (1 << 20);
    PyinWs w = pyin_ws_layout(h, n_utts, total_pyin_frames, (int32_t)max_T, nullptr);
    if (w.total + 256 > need) need = w.total + 256;
  }
  return need;
}

int roar_sup_logmel_energy(roar_sup_handle* h, const float* d_audio, const int64_t* d_sample_off,
                           const int32_t* d_sample_len, int32_t n_utts, const int64_t* d_frame_off,
                           int64_t total_frames, float* d_logmel, float* d_energy, void* d_ws,
                           size_t ws_bytes, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n_utts <= 0 || total_frames <= 0) return 0;   // empty batch: nothing to do (buffers may be null)
  if (!d_audio || !d_sample_off || !d_sample_len || !d_frame_off) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if (ws_bytes < a256(sizeof(int32_t) * (size_t)(n_utts + 1)) || !d_ws) return fail(ROAR_ERR_WORKSPACE, "workspace too small for roar_sup_logmel_energy");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* tile_off = (int32_t*)(((uintptr_t)d_ws + 255) & ~(uintptr_t)255);
  StftParams p = stft_params_base(h);
  p.audio = d_audio; p.sample_off = d_sample_off; p.sample_len = d_sample_len; p.frame_off = d_frame_off;
  p.tile_off = tile_off; p.n_utts = n_utts; p.logmel = d_logmel; p.energy = d_energy;
  LAUNCH(h, ROAR_K_TILE_OFFSETS, st, k_tile_offsets<<<1, 1024, 0, st>>>(d_frame_off, n_utts, p.FT, tile_off));
  const int64_t max_tiles = total_frames / p.FT + n_utts;
  {   // tile -> utterance map behind the prefix sums, when the caller's workspace has room for it
    const size_t off_bytes = a256(sizeof(int32_t) * (size_t)(n_utts + 1));
    const size_t lead = (size_t)((unsigned char*)tile_off - (unsigned char*)d_ws);
    if (lead + off_bytes + sizeof(int32_t) * (size_t)(max_tiles + 1) <= ws_bytes) {
      int32_t* map = (int32_t*)((unsigned char*)tile_off + off_bytes);
      k_tile_map<<<(unsigned)((max_tiles + 255) / 256), 256, 0, st>>>(tile_off, n_utts, map);
      p.tile_map = map;
    }
  }
  LAUNCH(h, ROAR_K_STFT_MEL, st, launch_stft(h, (unsigned)max_tiles, st, p));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int roar_sup_pyin(roar_sup_handle* h, const float* d_audio, const int64_t* d_sample_off,
                  const int32_t* d_sample_len, int32_t n_utts, const int64_t* d_frame_off,
                  int64_t total_frames, int32_t max_T, float* d_f0, float* d_vflag, float* d_vprob,
                  void* d_ws, size_t ws_bytes, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (!h->has_pyin) return fail(ROAR_ERR_UNSUPPORTED, "handle was created without pYIN (pyin_frame_length == 0)");
  if (n_utts <= 0 || total_frames <= 0) return 0;
  if (!d_audio || !d_sample_off || !d_sample_len || !d_frame_off || !d_f0 || !d_vflag || !d_vprob)
    return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if (max_T <= 0 || max_T > (1 << 20)) return fail(ROAR_ERR_INVALID_ARG, "max_frames_per_utt out of range");
  const Geometry& g = h->g;
  unsigned char* base = (unsigned char*)(((uintptr_t)d_ws + 255) & ~(uintptr_t)255);
  PyinWs w = pyin_ws_layout(h, n_utts, total_frames, max_T, base);
  if (!d_ws || w.total + 256 > ws_bytes) return fail(ROAR_ERR_WORKSPACE, "workspace too small for roar_sup_pyin");
  cudaStream_t st = (cudaStream_t)stream;

  PyinParams p = pyin_params_base(h);
  p.audio = d_audio; p.sample_off = d_sample_off; p.sample_len = d_sample_len; p.frame_off = d_frame_off;
  p.tile_off = w.tile_off; p.n_utts = n_utts; p.cmnd = w.big; p.energy = w.energy; p.etile_off = w.etile_off; p.cand_bin = w.cand_bin; p.cand_lp = w.cand_lp;
  p.n_cand = w.n_cand; p.lp_unvoiced = w.lp_unv; p.voiced_prob = d_vprob; p.total_frames = total_frames;
  // the tile -> utterance map region serves the energy kernel first (its tiles are the fewer), then K2a
  const int64_t max_etiles = total_frames / ENERGY_FT + n_utts;
  LAUNCH(h, ROAR_K_TILE_OFFSETS, st, k_tile_offsets<<<1, 1024, 0, st>>>(d_frame_off, n_utts, ENERGY_FT, w.etile_off));
  k_tile_map<<<(unsigned)((max_etiles + 255) / 256), 256, 0, st>>>(w.etile_off, n_utts, w.tile_map);
  p.etile_map = w.tile_map;
  LAUNCH(h, ROAR_K_PYIN_ENERGY, st, k_pyin_energy<<<(unsigned)max_etiles, ENERGY_THREADS, h->energy_smem, st>>>(p));
  p.etile_map = nullptr;
  LAUNCH(h, ROAR_K_TILE_OFFSETS, st, k_tile_offsets<<<1, 1024, 0, st>>>(d_frame_off, n_utts, p.FT, w.tile_off));
  const int64_t max_tiles = total_frames / p.FT + n_utts;
  k_tile_map<<<(unsigned)((max_tiles + 255) / 256), 256, 0, st>>>(w.tile_off, n_utts, w.tile_map);
  p.tile_map = w.tile_map;
  LAUNCH(h, ROAR_K_PYIN_CMND, st, k_pyin_cmnd<<<(unsigned)max_tiles, CMND_THREADS, h->cmnd_smem, st>>>(p));
  int64_t pb = (total_frames + 7) / 8;
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (pb > cap) pb = cap;
  LAUNCH(h, ROAR_K_PYIN_PROBS, st, k_pyin_probs<<<(unsigned)pb, 256, h->prob_smem, st>>>(p));

  // utterances longest-first
  CUDA_TRY(cudaMemsetAsync(w.hist, 0, sizeof(int32_t) * ((size_t)max_T + 2), st));
  const int nb = (n_utts + 255) / 256;
  LAUNCH(h, ROAR_K_LEN_SORT, st, k_len_hist<<<nb, 256, 0, st>>>(d_frame_off, n_utts, max_T, w.hist));
  LAUNCH(h, ROAR_K_LEN_SORT, st, k_len_scan<<<1, 1024, 0, st>>>(w.hist, max_T + 1));
  LAUNCH(h, ROAR_K_LEN_SORT, st, k_len_scatter<<<nb, 256, 0, st>>>(d_frame_off, n_utts, max_T, w.hist, w.order));

  VitParams v;
  memset(&v, 0, sizeof(v));
  v.frame_off = d_frame_off; v.order = w.order; v.n_utts = n_utts;
  v.npb = g.npb; v.tw = g.tw; v.hw = g.hw; v.kmax = g.kmax; v.n_rows = h->n_rows;
  v.lt_rows = h->d_lt_rows; v.row_id = h->d_row_id; v.lt0 = h->lt0; v.li_voiced = h->li_v; v.li_unvoiced = h->li_u;
  v.cand_bin = w.cand_bin; v.cand_lp = w.cand_lp; v.n_cand = w.n_cand; v.lp_unvoiced = w.lp_unv;
  v.ptr = (uint16_t*)w.big; v.last_state = w.last_state; v.freqs = h->d_freqs; v.f0 = d_f0; v.voiced_flag = d_vflag;
  v.lt_in_smem = h->lt_in_smem; v.lt_max = h->lt_max;
  v.uniform_vmax = h->uniform_vmax; v.twin_gap = h->twin_gap; v.flat_ok = h->flat_ok;
  for (int d = 0; d < VIT_TW; ++d) v.ltu[d] = h->ltu[d];
  v.ptr_stride = 2 * g.npb; v.ptr_uoff = g.npb;
  if (h->vit_fast) {
    LAUNCH(h, ROAR_K_VITERBI, st, (k_pyin_viterbi51<<<n_utts, h->vit_threads, h->vit3_smem, st>>>(v)));
  } else if (h->vit_threads <= 640) {
    if (h->lt_in_smem) LAUNCH(h, ROAR_K_VITERBI, st, (k_pyin_viterbi<true, 640><<<n_utts, h->vit_threads, h->vit_smem, st>>>(v)));
    else LAUNCH(h, ROAR_K_VITERBI, st, (k_pyin_viterbi<false, 640><<<n_utts, h->vit_threads, h->vit_smem, st>>>(v)));
  } else {
    if (h->lt_in_smem) LAUNCH(h, ROAR_K_VITERBI, st, (k_pyin_viterbi<true, 1024><<<n_utts, h->vit_threads, h->vit_smem, st>>>(v)));
    else LAUNCH(h, ROAR_K_VITERBI, st, (k_pyin_viterbi<false, 1024><<<n_utts, h->vit_threads, h->vit_smem, st>>>(v)));
  }
  LAUNCH(h, ROAR_K_BACKTRACK, st, k_pyin_backtrack<<<(n_utts + 127) / 128, 128, 0, st>>>(v));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int roar_sup_align_prior(roar_sup_handle* h, const int32_t* d_text_len, const int32_t* d_mel_len,
                         int32_t n_utts, const int64_t* d_out_off, int32_t max_mel_len,
                         double scaling_factor, float* d_prior, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n_utts <= 0 || max_mel_len <= 0) return 0;
  if (!d_text_len || !d_mel_len || !d_out_off || !d_prior) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  PriorParams p;
  p.text_len = d_text_len; p.mel_len = d_mel_len; p.out_off = d_out_off; p.out = d_prior;
  p.lf = h->d_lf; p.lf_n = h->lf_n; p.rows_per_cta = 64; p.scaling = scaling_factor;
  const unsigned gx = (max_mel_len + p.rows_per_cta - 1) / p.rows_per_cta;
  for (int32_t u0 = 0; u0 < n_utts; u0 += 65535) {
    p.utt_base = u0;
    const unsigned gy = n_utts - u0 < 65535 ? n_utts - u0 : 65535;
    LAUNCH(h, ROAR_K_PRIOR, st, k_align_prior<<<dim3(gx, gy), 256, 0, st>>>(p));
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int roar_sup_align_prior_interp(roar_sup_handle* h, const int32_t* d_text_len, const int32_t* d_mel_len,
                                int32_t n_utts, const int64_t* d_out_off, int32_t max_mel_len,
                                int32_t round_mel_len_to, int32_t round_text_len_to, float* d_prior,
                                void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n_utts <= 0 || max_mel_len <= 0) return 0;
  if (!d_text_len || !d_mel_len || !d_out_off || !d_prior) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if (round_mel_len_to < 1 || round_text_len_to < 1) return fail(ROAR_ERR_INVALID_ARG, "rounding steps must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  PriorInterpParams q;
  q.base.text_len = d_text_len; q.base.mel_len = d_mel_len; q.base.out_off = d_out_off; q.base.out = d_prior;
  q.base.lf = h->d_lf; q.base.lf_n = h->lf_n; q.base.rows_per_cta = 64; q.base.scaling = 1.0;
  q.round_mel = round_mel_len_to; q.round_text = round_text_len_to;
  const unsigned gx = (max_mel_len + q.base.rows_per_cta - 1) / q.base.rows_per_cta;
  for (int32_t u0 = 0; u0 < n_utts; u0 += 65535) {
    q.base.utt_base = u0;
    const unsigned gy = n_utts - u0 < 65535 ? n_utts - u0 : 65535;
    LAUNCH(h, ROAR_K_PRIOR, st, k_align_prior_interp<<<dim3(gx, gy), 256, 0, st>>>(q));
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// 16-bit PCM -> float32 in [-1, 1): x / 2^15, exact (AudioSegment._convert_samples_to_float32,
// asr/parts/preprocessing/segment.py:140-153).  Moves the wav decoder's integer -> float step onto the GPU so
// the host -> device copy carries 2 bytes per sample.
int roar_sup_pcm16_to_f32(roar_sup_handle* h, const int16_t* d_pcm, int64_t n_samples, float* d_audio, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n_samples <= 0) return 0;
  if (!d_pcm || !d_audio) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  const int64_t per_cta = 256 * 8;
  int64_t nb = (n_samples + per_cta - 1) / per_cta;
  const int64_t cap = (int64_t)h->sm_count * 16;
  if (nb > cap) nb = cap;
  LAUNCH(h, ROAR_K_PCM16, (cudaStream_t)stream, k_pcm16_to_f32<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(d_pcm, n_samples, d_audio));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int roar_sup_resample(roar_sup_handle* h, const float* d_in, const int64_t* d_in_off, const int32_t* d_in_len, int32_t n_utts,
                      int32_t max_out_len, int32_t up, int32_t down, const float* d_taps, int32_t n_taps, int32_t n_pre_pad,
                      int32_t n_pre_remove, float* d_out, const int64_t* d_out_off, const int32_t* d_out_len, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n_utts <= 0 || max_out_len <= 0) return 0;
  if (!d_in || !d_in_off || !d_in_len || !d_taps || !d_out || !d_out_off || !d_out_len) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if (up < 1 || down < 1 || n_taps < 1) return fail(ROAR_ERR_INVALID_ARG, "resample: up, down and n_taps must be >= 1");
  ResampleParams p;
  p.in = d_in; p.in_off = d_in_off; p.in_len = d_in_len; p.out_off = d_out_off; p.out_len = d_out_len; p.out = d_out;
  p.taps = d_taps; p.n_utts = n_utts; p.n_taps = n_taps; p.up = up; p.down = down; p.n_pre_pad = n_pre_pad; p.n_pre_remove = n_pre_remove;
  unsigned gx = (unsigned)((max_out_len + 255) / 256);
  if (gx > 1024) gx = 1024;
  for (int32_t u0 = 0; u0 < n_utts; u0 += 65535) {
    const unsigned gy = n_utts - u0 < 65535 ? n_utts - u0 : 65535;
    k_resample<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(p, u0);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// Metadata upload by kernel (see k_upload_small): `host_pinned` is page-locked host memory (cudaHostAlloc /
// cudaHostRegister; a PyTorch pinned tensor), read by the SMs through its device alias.
int roar_sup_upload(roar_sup_handle* h, const void* host_pinned, void* d_dst, size_t bytes, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (bytes == 0) return 0;
  if (!host_pinned || !d_dst) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if ((bytes & 15) || ((uintptr_t)host_pinned & 15) || ((uintptr_t)d_dst & 15))
    return fail(ROAR_ERR_INVALID_ARG, "roar_sup_upload needs 16-byte aligned pointers and size");
  void* alias = nullptr;
  if (cudaHostGetDevicePointer(&alias, const_cast<void*>(host_pinned), 0) != cudaSuccess) {
    cudaGetLastError();      // not page-locked / not mapped: take the copy engine
    CUDA_TRY(cudaMemcpyAsync(d_dst, host_pinned, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
  }
  const int64_t n16 = (int64_t)(bytes >> 4);
  int64_t nb = (n16 + 255) / 256;
  if (nb > 64) nb = 64;
  k_upload_small<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>((const uint4*)alias, (uint4*)d_dst, n16);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int roar_sup_trim(roar_sup_handle* h, const float* d_audio, const int64_t* d_sample_off,
                  const int32_t* d_sample_len, int32_t n_utts, int32_t max_samples_per_utt, double top_db,
                  double ref_value, int32_t frame_length, int32_t hop_length, int64_t* d_start, int64_t* d_end,
                  void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n_utts <= 0) return 0;
  if (!d_audio || !d_sample_off || !d_sample_len || !d_start || !d_end) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if (frame_length < 1 || hop_length < 1) return fail(ROAR_ERR_INVALID_ARG, "trim frame / hop length must be positive");
  const int64_t max_frames = 1 + (int64_t)max_samples_per_utt / hop_length;
  if (max_frames > 24576) return fail(ROAR_ERR_UNSUPPORTED, "utterance too long for roar_sup_trim (more than 24576 trim frames)");
  TrimParams p;
  p.audio = d_audio; p.sample_off = d_sample_off; p.sample_len = d_sample_len; p.n_utts = n_utts;
  p.frame_length = frame_length; p.hop_length = hop_length; p.max_frames = (int32_t)max_frames;
  p.top_db = top_db; p.ref_value = ref_value; p.start = d_start; p.end = d_end;
  const size_t smem = sizeof(double) * (size_t)max_frames;
  k_trim<<<n_utts, 256, smem, (cudaStream_t)stream>>>(p);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int roar_sup_pitch_partials_init(roar_sup_handle* h, double* d_out, int32_t n_groups, void* stream) {
  if (!h || !d_out || n_groups <= 0) return fail(ROAR_ERR_INVALID_ARG, "bad argument");
  k_stats_init<<<(n_groups + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_out, n_groups);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int roar_sup_pitch_partials(roar_sup_handle* h, const float* d_f0, int64_t n, double* d_out5, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n <= 0) return 0;
  if (!d_f0 || !d_out5) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  int64_t nb = (n + 256 * 8 - 1) / (256 * 8);
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (nb > cap) nb = cap;
  LAUNCH(h, ROAR_K_STATS, (cudaStream_t)stream, k_pitch_partials<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(d_f0, n, d_out5));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int roar_sup_pitch_partials_grouped(roar_sup_handle* h, const float* d_f0, const int64_t* d_frame_off,
                                    const int32_t* d_group, int32_t n_utts, int32_t n_groups, double* d_out,
                                    void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (n_utts <= 0) return 0;
  if (!d_f0 || !d_frame_off || !d_group || !d_out) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  const int64_t threads = (int64_t)n_utts * 32;
  k_pitch_partials_grouped<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_f0, d_frame_off, d_group, n_utts, n_groups, d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int64_t roar_fbank_out_frames(const roar_sup_handle* h, int64_t Lmax) {
  int64_t T = roar_sup_num_frames(h, Lmax);
  const int pt = h->cfg.pad_to;
  if (pt > 0 && T % pt != 0) T += pt - T % pt;
  return T;
}

__global__ void k_fbank_setup(const int64_t* len, int32_t B, int64_t Lmax, int64_t T_full, int32_t n_fft,
                              int32_t hop, int32_t pad_amount, int64_t* sample_off, int32_t* sample_len,
                              int64_t* frame_off, int64_t* out_len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    sample_off[i] = (int64_t)i * Lmax;
    sample_len[i] = (int32_t)Lmax;
    int64_t num = len[i] + pad_amount - n_fft;
    int64_t q = num >= 0 ? num / hop : -((-num + hop - 1) / hop);   // floor division
    out_len[i] = q + 1;
  }
  if (i <= B) frame_off[i] = (int64_t)i * T_full;
}

int roar_fbank_forward(roar_sup_handle* h, const float* d_x, const int64_t* d_len, int32_t B, int64_t Lmax,
                       float* d_out, int64_t* d_out_len, void* d_ws, size_t ws_bytes, void* stream) {
  if (!h || !d_x || !d_len || !d_out || !d_out_len) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if (B <= 0) return 0;
  const Geometry& g = h->g;
  const int64_t pad = h->cfg.exact_pad ? (g.n_fft - g.hop) / 2 : g.n_fft / 2;
  if (Lmax <= pad) return fail(ROAR_ERR_INVALID_ARG, "input shorter than the reflect padding");
  const int64_t T_full = roar_sup_num_frames(h, Lmax), Tpad = roar_fbank_out_frames(h, Lmax);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* base = (unsigned char*)(((uintptr_t)d_ws + 255) & ~(uintptr_t)255);
  size_t o = 0;
  auto take = [&](size_t bytes) { unsigned char* p = base + o; o += a256(bytes); return p; };
  int64_t* sample_off = (int64_t*)take(sizeof(int64_t) * (B + 1));
  int32_t* sample_len = (int32_t*)take(sizeof(int32_t) * (B + 1));
  int64_t* frame_off = (int64_t*)take(sizeof(int64_t) * (B + 1));
  int32_t* tile_off = (int32_t*)take(sizeof(int32_t) * (B + 1));
  if (!d_ws || o + 256 > ws_bytes) return fail(ROAR_ERR_WORKSPACE, "workspace too small for roar_fbank_forward");
  k_fbank_setup<<<(B + 256) / 256, 256, 0, st>>>(d_len, B, Lmax, T_full, g.n_fft, g.hop, (int32_t)(2 * pad),
                                                 sample_off, sample_len, frame_off, d_out_len);
  StftParams p = stft_params_base(h);
  p.audio = d_x; p.sample_off = sample_off; p.sample_len = sample_len; p.frame_off = frame_off;
  p.tile_off = tile_off; p.n_utts = B; p.logmel = d_out; p.energy = nullptr;
  p.out_utt_stride = (int64_t)g.n_mels * Tpad; p.out_row_stride = Tpad;
  LAUNCH(h, ROAR_K_TILE_OFFSETS, st, k_tile_offsets<<<1, 1024, 0, st>>>(frame_off, B, p.FT, tile_off));
  const int64_t max_tiles = (T_full * B) / p.FT + B;
  LAUNCH(h, ROAR_K_STFT_MEL, st, launch_stft(h, (unsigned)max_tiles, st, p));
  NormParams np;
  np.x = d_out; np.seq_len = d_out_len; np.B = B; np.n_mels = g.n_mels; np.T_full = (int32_t)T_full;
  np.Tpad = (int32_t)Tpad; np.mode = h->cfg.normalize; np.pad_value = (float)h->cfg.pad_value;
  dim3 grid(np.mode == ROAR_NORM_ALL_FEATURES ? 1 : g.n_mels, B);
  LAUNCH(h, ROAR_K_FBANK_NORM, st, k_fbank_normalize<<<grid, 256, 0, st>>>(np));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int roar_fbank_backward(roar_sup_handle* h, const float* d_x, const int64_t* d_len, int32_t B, int64_t Lmax,
                        const float* d_grad_out, float* d_grad_x, void* d_ws, size_t ws_bytes, void* stream) {
  if (!h) return fail(ROAR_ERR_INVALID_ARG, "null handle");
  if (B <= 0) return 0;
  if (!d_x || !d_len || !d_grad_out || !d_grad_x) return fail(ROAR_ERR_INVALID_ARG, "null argument");
  if (h->cfg.has_preemph) return fail(ROAR_ERR_UNSUPPORTED, "roar_fbank_backward: pre-emphasis is not differentiated (the reference's grad configs use preemph=None)");
  if (h->cfg.normalize != ROAR_NORM_NONE) return fail(ROAR_ERR_UNSUPPORTED, "roar_fbank_backward: normalize must be None (as in the reference's grad configs)");
  const Geometry& g = h->g;
  const int64_t pad = h->cfg.exact_pad ? (g.n_fft - g.hop) / 2 : g.n_fft / 2;
  if (Lmax <= pad) return fail(ROAR_ERR_INVALID_ARG, "input shorter than the reflect padding");
  const int64_t T_full = roar_sup_num_frames(h, Lmax), Tpad = roar_fbank_out_frames(h, Lmax);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* base = (unsigned char*)(((uintptr_t)d_ws + 255) & ~(uintptr_t)255);
  size_t o = 0;
  auto take = [&](size_t bytes) { unsigned char* p = base + o; o += a256(bytes); return p; };
  int64_t* sample_off = (int64_t*)take(sizeof(int64_t) * (B + 1));
  int32_t* sample_len = (int32_t*)take(sizeof(int32_t) * (B + 1));
  int64_t* frame_off = (int64_t*)take(sizeof(int64_t) * (B + 1));
  int32_t* tile_off = (int32_t*)take(sizeof(int32_t) * (B + 1));
  int64_t* valid = (int64_t*)take(sizeof(int64_t) * (B + 1));
  if (!d_ws || o + 256 > ws_bytes) return fail(ROAR_ERR_WORKSPACE, "workspace too small for roar_fbank_backward");
  k_fbank_setup<<<(B + 256) / 256, 256, 0, st>>>(d_len, B, Lmax, T_full, g.n_fft, g.hop, (int32_t)(2 * pad),
                                                 sample_off, sample_len, frame_off, valid);
  CUDA_TRY(cudaMemsetAsync(d_grad_x, 0, sizeof(float) * (size_t)B * (size_t)Lmax, st));
  StftBwdParams q;
  memset(&q, 0, sizeof(q));
  q.f = stft_params_base(h);
  q.f.use_tma = 0;
  q.f.audio = d_x; q.f.sample_off = sample_off; q.f.sample_len = sample_len; q.f.frame_off = frame_off;
  q.f.tile_off = tile_off; q.f.n_utts = B;
  q.f.out_utt_stride = (int64_t)g.n_mels * Tpad; q.f.out_row_stride = Tpad;
  q.grad_out = d_grad_out; q.valid_len = valid; q.grad_audio = d_grad_x;
  k_tile_offsets<<<1, 1024, 0, st>>>(frame_off, B, q.f.FT, tile_off);
  const size_t smem = h->stft_smem + stft_bwd_extra_carve(q.f, nullptr, nullptr);
  if (smem > h->max_smem) return fail(ROAR_ERR_UNSUPPORTED, "backward tile does not fit in shared memory");
  const int64_t max_tiles = (T_full * B) / q.f.FT + B;
  launch_stft_bwd((unsigned)max_tiles, smem, st, q);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- host I/O
int roar_sup_wav_probe_batch(const char* const* paths, int32_t n, roar_wav_info* out, int32_t n_threads) {
  if (n <= 0) return 0;
  if (!paths || !out) { fail(ROAR_ERR_INVALID_ARG, "null argument"); return n; }
  std::atomic<int> bad(0);
  std::string first_err;
  std::atomic<int> have_err(0);
  roar_io::parallel_for(n, n_threads, [&](int i) {
    std::string e;
    if (roar_io::wav_probe(paths[i], &out[i], &e) != 0) {
      bad.fetch_add(1);
      if (have_err.exchange(1) == 0) first_err = e;
    }
  });
  if (bad.load()) fail(ROAR_ERR_INVALID_ARG, first_err);
  return bad.load();
}

int roar_sup_wav_read_batch(const char* const* paths, const roar_wav_info* info, int32_t n, const int64_t* first_frame,
                            const int64_t* n_frames, int32_t channel, int32_t as_pcm16, void* dst, const int64_t* dst_off,
                            int32_t n_threads) {
  if (n <= 0) return 0;
  if (!paths || !info || !first_frame || !n_frames || !dst || !dst_off) { fail(ROAR_ERR_INVALID_ARG, "null argument"); return n; }
  std::atomic<int> bad(0), have_err(0);
  std::string first_err;
  roar_io::parallel_for(n, n_threads, [&](int i) {
    std::string e;
    const int rc = as_pcm16 ? roar_io::wav_read_pcm16(paths[i], info[i], first_frame[i], n_frames[i], (int16_t*)dst + dst_off[i], &e)
                            : roar_io::wav_read_f32(paths[i], info[i], first_frame[i], n_frames[i], channel, (float*)dst + dst_off[i], &e);
    if (rc != 0) {
      bad.fetch_add(1);
      if (have_err.exchange(1) == 0) first_err = e;
    }
  });
  if (bad.load()) fail(ROAR_ERR_INVALID_ARG, first_err);
  return bad.load();
}

int roar_sup_pt_write_batch(const float* base, int32_t n_files, const int64_t* elem_off, const int32_t* rank,
                            const int64_t* shape3, const char* const* paths, int32_t n_threads) {
  if (n_files <= 0) return 0;
  if (!base || !elem_off || !rank || !shape3 || !paths) { fail(ROAR_ERR_INVALID_ARG, "null argument"); return n_files; }
  // Small files of one directory are written by one thread: concurrent create/rename in the same directory
  // serialise on the directory lock (measured: 8 threads creating 2 KB files in one ext4 directory are 7x slower
  // than one thread), while different directories (one per sup-data type) proceed in parallel.  A directory of
  // large files (log_mel, ~180 KB each: checksum + copy dominate) is split over several threads by cost.
  std::vector<std::vector<int>> groups;
  {
    auto numel = [&](int i) { int64_t n = 1; for (int d = 0; d < rank[i]; ++d) n *= shape3[3 * (size_t)i + d]; return n; };
    auto cost = [&](int i) { return 20.0 + 4.0 * (double)numel(i) / 1000.0; };     // us: fixed + ~1 GB/s
    std::vector<std::pair<std::string, int>> key(n_files);
    double total = 0;
    for (int i = 0; i < n_files; ++i) {
      const char* sl = strrchr(paths[i], '/');
      key[i] = {sl ? std::string(paths[i], sl - paths[i]) : std::string("."), i};
      total += cost(i);
    }
    std::stable_sort(key.begin(), key.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    const double target = total / (n_threads > 0 ? n_threads : 1);
    for (int a = 0; a < n_files;) {
      int b = a;
      double c = 0, bytes = 0;
      while (b < n_files && key[b].first == key[a].first) { c += cost(key[b].second); bytes += 4.0 * (double)numel(key[b].second); ++b; }
      int parts = 1;
      if (bytes / (b - a) >= 16384.0 && target > 0) { parts = (int)(c / target + 0.5); if (parts < 1) parts = 1; if (parts > b - a) parts = b - a; }
      for (int q = 0; q < parts; ++q) {
        const int lo = a + (int)((int64_t)(b - a) * q / parts), hi = a + (int)((int64_t)(b - a) * (q + 1) / parts);
        groups.emplace_back();
        for (int i = lo; i < hi; ++i) groups.back().push_back(key[i].second);
      }
      a = b;
    }
    std::stable_sort(groups.begin(), groups.end(), [](const auto& x, const auto& y) { return x.size() > y.size(); });
  }
  std::atomic<int> bad(0), have_err(0);
  std::string first_err;
  roar_io::parallel_for((int)groups.size(), n_threads, [&](int g) {
    for (int i : groups[g]) {
      std::string e;
      if (roar_io::pt_write_f32(paths[i], base + elem_off[i], shape3 + 3 * (size_t)i, rank[i], &e) != 0) {
        bad.fetch_add(1);
        if (have_err.exchange(1) == 0) first_err = e;
      }
    }
  });
  if (bad.load()) fail(ROAR_ERR_INVALID_ARG, first_err);
  return bad.load();
}

}  // extern "C"
