// TEST INFRASTRUCTURE ONLY -- never loaded by the product (roar_b200/).
//
// Runs the per-thread code of the CUDA kernels (the HD functions of roar_b200/csrc/*.cuh, the same
// source the GPU compiles) on the CPU: the grid becomes loops over tiles, each __syncthreads()
// phase becomes a loop over thread ids.  Lets `pytest -m "not gpu"` check kernel logic against the
// oracle in a container without a GPU.  Device-only parts (TMA staging, warp shuffles) are
// replaced by their plain equivalents here and are covered by the `-m gpu` tests.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../roar_b200/csrc/common.cuh"
#include "../../roar_b200/csrc/fft.cuh"
#include "../../roar_b200/csrc/tables.hpp"
#include "../../roar_b200/csrc/k_stft_mel.cuh"
#include "../../roar_b200/csrc/k_stft_bwd.cuh"
#include "../../roar_b200/csrc/k_pyin_front.cuh"
#include "../../roar_b200/csrc/k_viterbi.cuh"
#include "../../roar_b200/csrc/k_misc.cuh"

using namespace roar;

extern "C" {

// ---------------------------------------------------------------------------- K1
int emu_logmel_energy(const roar_sup_config* cfg, const float* audio, int64_t L, float* logmel, float* energy) {
  if (!validate(*cfg).empty()) return -1;
  Geometry g = geometry(*cfg);
  std::vector<float> win = make_window(*cfg);
  std::vector<float> fb = make_mel_filterbank(*cfg);
  MelRows mr = make_mel_rows(fb, g.n_mels, g.n_bins);
  std::vector<cf32> tw = make_pass_twiddles<cf32, float>(g.M);
  std::vector<cf32> twp = make_twiddles<cf32, float>(g.n_fft, g.M + 1);
  const int NT = 256;
  StftParams p;
  memset(&p, 0, sizeof(p));
  p.n_fft = g.n_fft; p.hop = g.hop; p.M = g.M; p.n_bins = g.n_bins; p.n_mels = g.n_mels;
  p.P = g.M / 8; p.G = NT / p.P < 1 ? 1 : NT / p.P;
  p.FT = g.n_fft <= 1024 ? 16 : 8; if (p.FT < p.G) p.FT = p.G;
  p.span = (p.FT - 1) * g.hop + g.n_fft;
  p.pad_left = cfg->exact_pad ? (g.n_fft - g.hop) / 2 : g.n_fft / 2;
  p.floor_ = (float)cfg->spec_floor; p.mag_power = (float)cfg->mag_power; p.log_guard = (float)cfg->log_guard;
  p.preemph = (float)cfg->preemph; p.log_mode = cfg->log_mode; p.has_preemph = cfg->has_preemph;
  p.preemph_after_pad = cfg->exact_pad; p.energy_mode = cfg->energy_mode;
  p.window = win.data(); p.tw = tw.data(); p.tw_post = twp.data();
  p.mel_start = mr.start.data(); p.mel_count = mr.count.data(); p.mel_offset = mr.offset.data();
  p.mel_w = mr.weights.data(); p.mel_nw = (int)mr.weights.size();
  int64_t T = cfg->exact_pad ? (L + 2 * p.pad_left - g.n_fft) / g.hop + 1 : 1 + L / g.hop;
  int64_t sample_off[1] = {0}; int32_t sample_len[1] = {(int32_t)L};
  int64_t frame_off[2] = {0, T};
  int32_t n_tiles = (int32_t)((T + p.FT - 1) / p.FT);
  int32_t tile_off[2] = {0, n_tiles};
  p.audio = audio; p.sample_off = sample_off; p.sample_len = sample_len; p.frame_off = frame_off;
  p.tile_off = tile_off; p.n_utts = 1; p.logmel = logmel; p.energy = energy;
  std::vector<unsigned char> smem(stft_smem_carve(p, NT, nullptr, nullptr) + 64);
  for (int tile = 0; tile < n_tiles + 1; ++tile) {
    StftSmem s;
    stft_smem_carve(p, NT, smem.data(), &s);
    StftTile t;
    if (!stft_locate(p, tile, &t)) continue;
    const int hi = (t.nf - 1) * p.hop + p.n_fft;
    for (int tid = 0; tid < NT; ++tid) { stft_phase_tables(p, s, tid, NT); stft_phase_audio(p, t, s, tid, NT, 0, hi); }
    const int n_groups = (t.nf + p.G - 1) / p.G;
    for (int gi = 0; gi < n_groups; ++gi) {
      for (int tid = 0; tid < NT; ++tid) stft_first_pass<8>(p, t, s, gi, tid);
      const FftPlan plan = make_plan(p.M);
      int Ns = plan.radix[0];
      const cf32* src = s.bufA; cf32* dst = s.bufB;
      for (int ps = 1; ps < plan.n_pass; ++ps) {
        for (int tid = 0; tid < NT; ++tid)
          stft_pass_any(plan.radix[ps], p, t, s, gi, tid, Ns, p.tw + plan.tw_off[ps], src, dst);
        Ns *= plan.radix[ps];
        const cf32* tmp = src; src = dst; dst = const_cast<cf32*>(tmp);
      }
      float* spec = (float*)dst;
      for (int tid = 0; tid < NT; ++tid) stft_phase_post(p, t, s, gi, tid, src, spec);
      for (int tid = 0; tid < NT; ++tid) stft_phase_mel(p, t, s, gi, tid, NT, spec);
    }
    for (int tid = 0; tid < NT; ++tid) stft_phase_store(p, t, s, tid, NT);
  }
  return 0;
}

// ---------------------------------------------------------------------------- K2 + K3
// stage outputs (any may be null): cmnd [T, n_lags]; states [T]
int emu_pyin(const roar_sup_config* cfg, const float* audio, int64_t L, float* f0, float* vflag, float* vprob,
             double* cmnd_out, int32_t* states_out) {
  if (!validate(*cfg).empty()) return -1;
  Geometry g = geometry(*cfg);
  PyinTables tb = make_pyin_tables(*cfg, g);
  const int NT = 256;
  PyinParams p;
  memset(&p, 0, sizeof(p));
  p.F = g.pf; p.W = g.pw; p.hop = g.ph;
  p.min_period = g.min_period; p.max_period = g.max_period; p.n_lags = g.n_lags;
  p.FT = g.pf <= 1024 ? 15 : 6;
  cmnd_blocking(g.pw, g.ph, &p.BL, &p.nb);
  p.n_groups = (g.max_period + 1 + ACF_R - 1) / ACF_R;
  p.ylen = cmnd_ylen(p.FT, g.pf, g.ph, p.BL, p.nb, p.n_groups);
  p.npb = g.npb; p.nbps = g.nbps; p.kmax = g.kmax; p.n_thr = g.n_thr;
  p.sr = cfg->sample_rate; p.fmin = cfg->pitch_fmin; p.no_trough_prob = cfg->no_trough_prob;
  p.thresholds = tb.thresholds.data(); p.beta_probs = tb.beta_probs.data();
  p.beta_cum = tb.beta_cum.data(); p.boltz_exp = tb.boltz_exp.data(); p.boltz_fact = tb.boltz_fact.data();
  const int64_t T = 1 + L / g.ph;
  int64_t sample_off[1] = {0}; int32_t sample_len[1] = {(int32_t)L};
  int64_t frame_off[2] = {0, T};
  int32_t n_tiles = (int32_t)((T + p.FT - 1) / p.FT);
  int32_t tile_off[2] = {0, n_tiles};
  std::vector<double> cmnd((size_t)T * g.n_lags), cand_lp((size_t)T * g.kmax), lp_unv(T);
  std::vector<uint16_t> cand_bin((size_t)T * g.kmax);
  std::vector<int32_t> n_cand(T);
  p.audio = audio; p.sample_off = sample_off; p.sample_len = sample_len; p.frame_off = frame_off;
  p.tile_off = tile_off; p.n_utts = 1; p.cmnd = cmnd.data(); p.cand_bin = cand_bin.data();
  p.cand_lp = cand_lp.data(); p.n_cand = n_cand.data(); p.lp_unvoiced = lp_unv.data(); p.voiced_prob = vprob;
  p.total_frames = T;
  // ---- K2a-0 energy, K2a
  std::vector<float> en((size_t)T * (g.max_period + 1));
  p.energy = en.data();
  {
    std::vector<float> ys((size_t)epad(energy_span(p), g.ph) + 4);
    for (int64_t t0 = 0; t0 < T; t0 += ENERGY_FT) {
      const int nf = (int)(T - t0 < ENERGY_FT ? T - t0 : ENERGY_FT);
      const int n = (nf - 1) * g.ph + g.pw + g.max_period + 1;
      for (int l = 0; l < 32; ++l) pyin_energy_stage(p, audio, (int)L, t0 * g.ph - g.pf / 2, n, ys.data(), l, 32);
      for (int l = 0; l < nf; ++l) pyin_energy_frame(p, ys.data(), l, en.data() + t0 + l, (size_t)T);
    }
  }
  std::vector<unsigned char> smem(cmnd_smem_carve(p, nullptr, nullptr) + 64);
  for (int tile = 0; tile < n_tiles; ++tile) {
    CmndSmem s;
    cmnd_smem_carve(p, smem.data(), &s);
    PyinTile t;
    if (!pyin_locate(p, tile, &t)) continue;
    for (int tid = 0; tid < NT; ++tid) cmnd_phase_load(p, t, s, tid, NT);
    const int n_units = cmnd_n_blocks(p, t.nf) * p.n_groups;
    for (int u = 0; u < n_units; ++u) cmnd_acf_unit(p, s, u / p.n_groups, u % p.n_groups);
    for (int tid = 0; tid < NT; ++tid) cmnd_phase_diff(p, t, s, tid, NT);
    for (int f = 0; f < t.nf; ++f) {
      const int slot = f % CMND_SLOTS;
      for (int l = 0; l < 32; ++l) cmnd_phase_scan1(p, s, f, slot, l);
      for (int l = 0; l < 32; ++l) cmnd_phase_scan2(p, s, slot, l);
      for (int l = 0; l < 32; ++l) cmnd_phase_emit(p, t, s, f, slot, l);
    }
  }
  if (cmnd_out) memcpy(cmnd_out, cmnd.data(), cmnd.size() * sizeof(double));
  // ---- K2b
  {
    std::vector<unsigned char> sm(prob_smem_carve(p, nullptr, nullptr) + 64);
    ProbSmem s;
    prob_smem_carve(p, sm.data(), &s);
    for (int64_t fr = 0; fr < T; ++fr) {
      for (int l = 0; l < 32; ++l) prob_phase0(p, s, fr, l);
      for (int l = 0; l < 32; ++l) prob_phase1(p, s, l);
      for (int l = 0; l < 32; ++l) prob_phase2(p, s, l);
      for (int l = 0; l < 32; ++l) prob_phase3(p, s, l, tb.thresholds.data());
      for (int l = 0; l < 32; ++l) prob_phase4(p, s, l);
      {
        const int R = s.cnt[32];
        const int na = prob_compact_active(p, s);
        for (int r = 0; r < R; ++r) if ((int)s.cr[r] >= p.n_thr) prob_trough_finish(p, s, r, 0.0);
        for (int a = 0; a < na; ++a) prob_trough_finish(p, s, s.sorted[a], prob_active_sum(p, s, na, a));
      }
      for (int l = 0; l < 32; ++l) prob_phase6a(p, s, l);
      for (int l = 0; l < 32; ++l) prob_phase6b(p, s, fr, l);
    }
  }
  // ---- K3 (same per-bin step functions; the block reductions are plain scans here)
  if (getenv("ROAR_EMU_VERBOSE")) {
    long tot = 0, mx = 0, over = 0, over16=0; double mlp = 0;
    for (int64_t t = 0; t < T; ++t) { tot += n_cand[t]; if (n_cand[t] > mx) mx = n_cand[t]; if (n_cand[t] > 32) ++over; if (n_cand[t] > 16) ++over16; if (lp_unv[t] < mlp) mlp = lp_unv[t]; }
    fprintf(stderr, "cand: T %ld mean %.2f max %ld frames>32 %ld >16 %ld min lp_u %.3f\n", (long)T, (double)tot / T, mx, over, over16, mlp);
  }
  VitParams v;
  memset(&v, 0, sizeof(v));
  v.npb = g.npb; v.tw = g.tw; v.hw = g.hw; v.kmax = g.kmax; v.n_rows = tb.n_rows;
  v.lt0 = tb.lt0; v.li_voiced = tb.li_voiced; v.li_unvoiced = tb.li_unvoiced;
  const int npb = g.npb;
  std::vector<cf64> V(2 * (size_t)npb);
  std::vector<double> lpv(npb, tb.lt0);
  std::vector<uint16_t> ptr((size_t)T * 2 * npb);
  std::vector<int32_t> row_ofs(npb);
  for (int i = 0; i < npb; ++i) row_ofs[i] = (int32_t)tb.row_id[i] * g.tw;
  auto argmax = [&](const cf64* Vc, int* k, double* m) {
    double best = -1e308; int bk = 0x7fffffff;
    for (int b = 0; b < 2; ++b)
      for (int j = 0; j < npb; ++j) {
        const double val = b ? Vc[j].y : Vc[j].x;
        if (val > best) { best = val; bk = b * npb + j; }
      }
    *k = bk; *m = best;
  };
  for (int c = 0; c < n_cand[0]; ++c) lpv[cand_bin[c]] = cand_lp[c];
  for (int j = 0; j < npb; ++j) { V[j].x = lpv[j] + tb.li_voiced; V[j].y = lp_unv[0] + tb.li_unvoiced; }
  for (int c = 0; c < n_cand[0]; ++c) lpv[cand_bin[c]] = tb.lt0;
  const char* env_v = getenv("ROAR_SUP_VITERBI");
  const bool geom_ok = g.tw == VIT_TW && tb.n_rows + 1 <= 64 && npb <= 608;
  const int fast = !geom_ok || (env_v && env_v[0] == 'g') ? 0 : 1;
  if (fast == 1) {
    // fast path (k_pyin_viterbi51): padded V rows, live lists, dominance skipping
    const int VP = npb + 2 * VIT_HW;
    std::vector<double> Vv(2 * (size_t)VP, VIT_NEG), Vu(2 * (size_t)VP, VIT_NEG);
    std::vector<cf64> lt2((size_t)(tb.n_rows + 1) * VIT_TW);
    for (size_t i = 0; i < (size_t)tb.n_rows * VIT_TW; ++i) { lt2[i].x = tb.lt_rows[2 * i]; lt2[i].y = tb.lt_rows[2 * i + 1]; }
    for (size_t i = (size_t)tb.n_rows * VIT_TW; i < lt2.size(); ++i) { lt2[i].x = 0; lt2[i].y = 0; }
    std::vector<VitLive> Lv(3 * VIT_LIST_MAX), Lu(3 * VIT_LIST_MAX);
    int cnt[3][2] = {{VIT_LIST_MAX + 1, VIT_LIST_MAX + 1}, {0, 0}, {0, 0}};
    std::vector<uint8_t> rid((size_t)npb + 2 * VIT_HW);
    vit_pad_rows(tb.row_id.data(), npb, tb.n_rows, rid.data());
    for (int j = 0; j < npb; ++j) { Vv[VIT_HW + j] = V[j].x; Vu[VIT_HW + j] = V[j].y; }
    long skipped = 0, listed = 0, n_sparse = 0, n_uniform = 0, u_over = 0, n_doa = 0, n_flat = 0;
    bool sparse_prev = false;
    double prev_vmax = 0.0;
    for (int64_t t = 1; t < T; ++t) {
      const int rp = (int)((t - 1) & 1), wp = (int)(t & 1), rl = (int)((t - 1) % 3), wl = (int)(t % 3), zl = (int)((t + 1) % 3);
      Vit3Step c;
      c.Vv = Vv.data() + (size_t)rp * VP; c.Vu = Vu.data() + (size_t)rp * VP;
      c.vmax = -1e308; c.kstar = 0x7fffffff; c.vvmax = -1e308;
      for (int b = 0; b < 2; ++b)
        for (int j = 0; j < npb; ++j) {
          const double val = b ? c.Vu[VIT_HW + j] : c.Vv[VIT_HW + j];
          if (val > c.vmax) { c.vmax = val; c.kstar = b * npb + j; }
          if (!b && val > c.vvmax) c.vvmax = val;
        }
      c.Lv = Lv.data() + (size_t)rl * VIT_LIST_MAX; c.Lu = Lu.data() + (size_t)rl * VIT_LIST_MAX;
      c.nv = cnt[rl][0]; c.nu = cnt[rl][1];
      c.lt2 = lt2.data(); c.lt0 = tb.lt0; c.lt_max = tb.lt_max; c.npb = npb;
      cnt[zl][0] = cnt[zl][1] = 0;
      if (c.nv <= VIT_LIST_MAX) ++listed; else ++skipped;
      const double lb = vit3_lower_bound(c, tb.row_id.data(), lp_unv[t], cand_bin.data() + (size_t)t * g.kmax,
                                         cand_lp.data() + (size_t)t * g.kmax, n_cand[t], 0, 1);
      const double thr = lb + tb.lt0;
      const double lmg = getenv("ROAR_EMU_WEAK") ? 0.0 : tb.lt_max;   // liveness margin (see k_viterbi.cuh)
      if (c.nu > VIT_LIST_MAX) ++u_over;
      for (int q = 0; q < n_cand[t]; ++q) lpv[cand_bin[(size_t)t * g.kmax + q]] = cand_lp[(size_t)t * g.kmax + q];
      const bool sparse = lp_unv[t] >= tb.lt0 + VIT_SPARSE_MARGIN;
      // dead-on-arrival dense step (k_viterbi.cuh): every state without a candidate observation is dead the
      // moment it is created, only the candidate bins are evaluated
      const double top = c.vmax + tb.lt_max;
      const bool doa = !sparse && !getenv("ROAR_EMU_NO_DOA") && ((tb.lt0 + top) + tb.lt_max) < thr && ((lp_unv[t] + top) + tb.lt_max) < thr;
      if (doa) ++n_doa;
      if (sparse || doa) {
        if (sparse) ++n_sparse;
        // candidate bins (one warp each on the device; here the "lanes" are reduced by vit_offer)
        for (int q = 0; q < n_cand[t]; ++q) {
          const int b = cand_bin[(size_t)t * g.kmax + q];
          VitBest2 x; x.b = VIT_NEG; x.a = 0x7fffffff;
          for (int lane = 0; lane < 32; ++lane) {
            const VitBest2 y = vit4_cand_partial(c, tb.row_id.data(), b, lane, 32);
            vit_offer(x, y.b, y.a);
          }
          const double cv = cand_lp[(size_t)t * g.kmax + q] + x.b;
          Vv[(size_t)wp * VP + VIT_HW + b] = cv;
          ptr[(size_t)t * 2 * npb + b] = (uint16_t)x.a;
          if (cv + lmg >= thr) { if (cnt[wl][0] < VIT_LIST_MAX) { VitLive e; e.v = cv; e.kb = b; e.row = tb.row_id[b]; Lv[(size_t)wl * VIT_LIST_MAX + cnt[wl][0]] = e; } ++cnt[wl][0]; }
        }
        c.vvmax = 0.0;   // placeholder, replaced per bin below
      }
      for (int j = 0; j < npb; ++j) {
        double nv, nu; int av, au;
        if (sparse_prev) {
          c.vvmax = 0.0;  // the kernel keeps no per-segment voiced maxima after a sparse step
        } else {   // the kernel's per-warp bound: voiced maximum over the 32-bin segments w-1, w, w+1
          const int w = j / 32, nw = (npb + 31) / 32;
          double seg = -1e308;
          for (int ww = (w > 0 ? w - 1 : 0); ww <= (w + 1 < nw ? w + 1 : nw - 1); ++ww)
            for (int q = ww * 32; q < ww * 32 + 32 && q < npb; ++q) if (c.Vv[VIT_HW + q] > seg) seg = c.Vv[VIT_HW + q];
          c.vvmax = seg;
        }
        if (doa) {
          if (lpv[j] == tb.lt0) Vv[(size_t)wp * VP + VIT_HW + j] = VIT_NEG;
          Vu[(size_t)wp * VP + VIT_HW + j] = VIT_NEG;
          continue;
        }
        if (sparse) {
          const int w = j / 32;
          const bool uni = (int)tb.lt_uniform.size() == VIT_TW && prev_vmax <= tb.uniform_vmax && 32 * w >= 2 * VIT_HW &&
                           32 * w + 31 + 2 * VIT_HW <= npb - 1;
          if (uni) ++n_uniform;
          if (c.nv > VIT_LIST_MAX) c.vvmax = 0.0;
          // the kernel's per-warp filter of the voiced live list (reach of the warp's 32 bins, twin dominance)
          c.lv_mask = 0xffffffffu;
          if (c.nv <= VIT_LIST_MAX && !getenv("ROAR_EMU_NO_TWIN")) {
            c.lv_mask = 0;
            for (int e = 0; e < c.nv; ++e) {
              const VitLive& le = c.Lv[e];
              bool need = (unsigned)(le.kb - (32 * w - VIT_HW)) <= (unsigned)(31 + 2 * VIT_HW);
              if (need && c.nu > VIT_LIST_MAX) need = !(le.v - c.Vu[VIT_HW + le.kb] < tb.twin_gap);
              if (need) c.lv_mask |= 1u << e;
            }
          }
          // flat-segment rule: the three 32-bin segments in reach hold one unvoiced value
          c.u_flat = false;
          if (uni && tb.flat_ok && !getenv("ROAR_EMU_NO_FLAT")) {
            const double x0 = c.Vu[VIT_HW + 32 * w];
            bool flat = x0 >= -1e12;
            for (int q2 = 32 * (w - 1); flat && q2 < 32 * (w + 2); ++q2)
              if (q2 >= 0 && q2 < npb && c.Vu[VIT_HW + q2] != x0) flat = false;
            c.u_flat = flat; c.u_flat_val = x0;
            if (flat && c.nu > VIT_LIST_MAX) ++n_flat;
          }
          const VitBest2 bu = vit4_unvoiced_scan(c, j, rid.data() + j, uni ? tb.lt_uniform.data() : nullptr);
          vit4_unvoiced_finish(bu, npb, tb.lt0, c.vmax, c.kstar, j, lp_unv[t], &nu, &au);
          if (lpv[j] == tb.lt0) Vv[(size_t)wp * VP + VIT_HW + j] = VIT_NEG;
          Vu[(size_t)wp * VP + VIT_HW + j] = nu;
          ptr[(size_t)t * 2 * npb + npb + j] = (uint16_t)au;
        } else {
          vit3_step_bin(c, j, rid.data() + j, lpv[j], lp_unv[t], &nv, &nu, &av, &au);
          Vv[(size_t)wp * VP + VIT_HW + j] = nv; Vu[(size_t)wp * VP + VIT_HW + j] = nu;
          ptr[(size_t)t * 2 * npb + j] = (uint16_t)av; ptr[(size_t)t * 2 * npb + npb + j] = (uint16_t)au;
          if (nv + lmg >= thr) { if (cnt[wl][0] < VIT_LIST_MAX) { VitLive e; e.v = nv; e.kb = j; e.row = tb.row_id[j]; Lv[(size_t)wl * VIT_LIST_MAX + cnt[wl][0]] = e; } ++cnt[wl][0]; }
        }
        if (nu + lmg >= thr) { if (cnt[wl][1] < VIT_LIST_MAX) { VitLive e; e.v = nu; e.kb = j; e.row = tb.row_id[j]; Lu[(size_t)wl * VIT_LIST_MAX + cnt[wl][1]] = e; } ++cnt[wl][1]; }
      }
      sparse_prev = sparse || doa;
      prev_vmax = c.vmax;
      for (int q = 0; q < n_cand[t]; ++q) lpv[cand_bin[(size_t)t * g.kmax + q]] = tb.lt0;
    }
    const int lp_ = (int)((T - 1) & 1);
    for (int j = 0; j < npb; ++j) { V[(size_t)lp_ * npb + j].x = Vv[(size_t)lp_ * VP + VIT_HW + j]; V[(size_t)lp_ * npb + j].y = Vu[(size_t)lp_ * VP + VIT_HW + j]; }
    if (getenv("ROAR_EMU_VERBOSE")) fprintf(stderr, "viterbi fast: %ld list steps, %ld overflow steps, %ld sparse steps, %ld uniform bin-steps (uniform_vmax %g), %ld unvoiced-overflow steps, %ld dead-on-arrival dense steps of %ld, %ld flat bin-scans\n", listed, skipped, n_sparse, n_uniform, tb.uniform_vmax, u_over, n_doa, (long)T - 1, n_flat);
  } else {
  for (int64_t t = 1; t < T; ++t) {
    const cf64* Vc = V.data() + (size_t)((t - 1) & 1) * npb;
    cf64* Vn = V.data() + (size_t)(t & 1) * npb;
    int kstar; double vmax;
    argmax(Vc, &kstar, &vmax);
    for (int c = 0; c < n_cand[t]; ++c) lpv[cand_bin[(size_t)t * g.kmax + c]] = cand_lp[(size_t)t * g.kmax + c];
    for (int j = 0; j < npb; ++j)
      vit_step_bin(v, j, Vc, tb.lt_rows.data(), row_ofs.data(), lpv[j], lp_unv[t], kstar, vmax, &Vn[j],
                   ptr.data() + (size_t)t * 2 * npb);
    for (int c = 0; c < n_cand[t]; ++c) lpv[cand_bin[(size_t)t * g.kmax + c]] = tb.lt0;
  }
  }
  int s; double m;
  argmax(V.data() + (size_t)((T - 1) & 1) * npb, &s, &m);
  for (int64_t t = T - 1; t >= 0; --t) {
    const bool voiced = s < npb;
    if (states_out) states_out[t] = s;
    f0[t] = voiced ? (float)tb.freqs[s] : 0.f;
    vflag[t] = voiced ? 1.f : 0.f;
    if (t > 0) s = ptr[(size_t)t * 2 * npb + s];
  }
  return 0;
}

// ---------------------------------------------------------------------------- K4
// Brute-force check of tables.hpp make_uniform_row: for every interior source row, every band offset
// and `n_rand` doubles V <= uniform_vmax per binade (random mantissas plus the binade's corner cases),
// fl(V + lt_row) must equal fl(V + lt_uniform).
// out[0] = mismatches, out[1] = differing table entries, out[2] = -uniform_vmax, out[3] = sums compared.
int emu_uniform_row_check(const roar_sup_config* cfg, int32_t n_rand, double* out) {
  if (!validate(*cfg).empty()) return -1;
  Geometry g = geometry(*cfg);
  PyinTables tb = make_pyin_tables(*cfg, g);
  out[0] = out[1] = out[2] = out[3] = 0;
  if ((int)tb.lt_uniform.size() != g.tw) return 1;
  out[2] = -tb.uniform_vmax;
  uint64_t rs = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; };
  long bad = 0, differing = 0, compared = 0;
  for (int i = g.hw; i <= g.npb - 1 - g.hw; ++i) {
    const int r = tb.row_id[i];
    for (int d = 0; d < g.tw; ++d) {
      const double a = tb.lt_rows[((size_t)r * g.tw + d) * 2], b = tb.lt_uniform[d];
      if (a == b) continue;
      ++differing;
      int e0; std::frexp(tb.uniform_vmax == 0.0 ? -1.0 : tb.uniform_vmax, &e0);   // |vmax| = 2^(e0-1)
      for (int e = e0 - 1; e < 40; ++e) {
        const double u = std::ldexp(1.0, e - 52);
        for (int k = 0; k < n_rand; ++k) {
          // random mantissa, plus the structural cases: even / odd multiple of u at the bottom of the
          // binade (sum stays inside) and at its top (sum crosses into the next binade)
          double V = -std::ldexp(1.0 + (double)(rnd() >> 12) * std::ldexp(1.0, -52), e);
          if (k == 0) V = -std::ldexp(1.0, e);
          if (k == 1) V = -(std::ldexp(1.0, e) + u);
          if (k == 2) V = -(std::ldexp(1.0, e + 1) - u);
          if (k == 3) V = -(std::ldexp(1.0, e + 1) - 2 * u);
          if (V > tb.uniform_vmax) continue;
          volatile double s1 = V + a, s2 = V + b;
          ++compared;
          if (s1 != s2) ++bad;
        }
      }
    }
  }
  out[0] = (double)bad; out[1] = (double)differing; out[3] = (double)compared;
  return 0;
}

int emu_prior(int32_t N, int32_t M, double scaling, float* out) {
  std::vector<double> lf(4096);
  for (size_t i = 0; i < lf.size(); ++i) lf[i] = std::lgamma((double)i + 1.0);
  PriorParams p;
  memset(&p, 0, sizeof(p));
  p.lf = lf.data(); p.lf_n = (int)lf.size(); p.scaling = scaling;
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < N; ++k)
      out[(size_t)m * N + k] = scaling == 1.0 ? prior_value_int(p, N, M, m + 1, k) : prior_value_real(p, N, M, m + 1, k);
  return 0;
}

// ---------------------------------------------------------------------------- K1b backward (one utterance, dense)
// x [Lmax] with `len` valid samples; grad_out [n_mels, T] (T = frames of Lmax); grad_x [Lmax]
int emu_fbank_backward(const roar_sup_config* cfg, const float* x, int64_t Lmax, int64_t len, const float* grad_out,
                       float* grad_x) {
  if (!validate(*cfg).empty()) return -1;
  Geometry g = geometry(*cfg);
  std::vector<float> win = make_window(*cfg);
  std::vector<float> fb = make_mel_filterbank(*cfg);
  MelRows mr = make_mel_rows(fb, g.n_mels, g.n_bins);
  std::vector<cf32> tw = make_pass_twiddles<cf32, float>(g.M);
  std::vector<cf32> twp = make_twiddles<cf32, float>(g.n_fft, g.M + 1);
  const int NT = 256;
  StftBwdParams q;
  memset(&q, 0, sizeof(q));
  StftParams& p = q.f;
  p.n_fft = g.n_fft; p.hop = g.hop; p.M = g.M; p.n_bins = g.n_bins; p.n_mels = g.n_mels;
  p.P = g.M / 8; p.G = NT / p.P < 1 ? 1 : NT / p.P;
  p.FT = g.n_fft <= 1024 ? 16 : 8; if (p.FT < p.G) p.FT = p.G;
  p.span = (p.FT - 1) * g.hop + g.n_fft;
  p.pad_left = cfg->exact_pad ? (g.n_fft - g.hop) / 2 : g.n_fft / 2;
  p.floor_ = (float)cfg->spec_floor; p.mag_power = (float)cfg->mag_power; p.log_guard = (float)cfg->log_guard;
  p.log_mode = cfg->log_mode; p.has_preemph = 0;
  p.window = win.data(); p.tw = tw.data(); p.tw_post = twp.data();
  p.mel_start = mr.start.data(); p.mel_count = mr.count.data(); p.mel_offset = mr.offset.data();
  p.mel_w = mr.weights.data(); p.mel_nw = (int)mr.weights.size();
  const int64_t T = cfg->exact_pad ? (Lmax + 2 * p.pad_left - g.n_fft) / g.hop + 1 : 1 + Lmax / g.hop;
  int64_t num = len + 2 * p.pad_left - g.n_fft;
  int64_t valid[1] = {(num >= 0 ? num / g.hop : -((-num + g.hop - 1) / g.hop)) + 1};
  int64_t sample_off[1] = {0}; int32_t sample_len[1] = {(int32_t)Lmax};
  int64_t frame_off[2] = {0, T};
  int32_t n_tiles = (int32_t)((T + p.FT - 1) / p.FT);
  int32_t tile_off[2] = {0, n_tiles};
  p.audio = x; p.sample_off = sample_off; p.sample_len = sample_len; p.frame_off = frame_off;
  p.tile_off = tile_off; p.n_utts = 1;
  p.out_utt_stride = (int64_t)g.n_mels * T; p.out_row_stride = T;
  q.grad_out = grad_out; q.valid_len = valid; q.grad_audio = grad_x;
  for (int64_t i = 0; i < Lmax; ++i) grad_x[i] = 0.f;
  const size_t fwd_bytes = stft_smem_carve(p, NT, nullptr, nullptr);
  std::vector<unsigned char> smem(fwd_bytes + stft_bwd_extra_carve(p, nullptr, nullptr) + 64);
  const FftPlan plan = make_plan(p.M);
  for (int tile = 0; tile < n_tiles; ++tile) {
    StftSmem s;
    stft_smem_carve(p, NT, smem.data(), &s);
    StftBwdSmem b;
    stft_bwd_extra_carve(p, smem.data() + fwd_bytes, &b);
    StftTile t;
    if (!stft_locate(p, tile, &t)) continue;
    const int hi = (t.nf - 1) * p.hop + p.n_fft;
    for (int tid = 0; tid < NT; ++tid) { stft_phase_tables(p, s, tid, NT); stft_phase_audio(p, t, s, tid, NT, 0, hi); }
    for (int i = 0; i < p.span; ++i) b.gspan[i] = 0.f;
    const int n_groups = (t.nf + p.G - 1) / p.G;
    for (int gi = 0; gi < n_groups; ++gi) {
      for (int tid = 0; tid < NT; ++tid) stft_first_pass<8>(p, t, s, gi, tid);
      int Ns = plan.radix[0];
      cf32* src = s.bufA; cf32* dst = s.bufB;
      for (int ps = 1; ps < plan.n_pass; ++ps) {
        for (int tid = 0; tid < NT; ++tid) stft_pass_any(plan.radix[ps], p, t, s, gi, tid, Ns, p.tw + plan.tw_off[ps], src, dst);
        Ns *= plan.radix[ps];
        cf32* tmp = src; src = dst; dst = tmp;
      }
      for (int tid = 0; tid < NT; ++tid) bwd_phase_spectrum(p, t, b, gi, tid, src, dst);
      for (int tid = 0; tid < NT; ++tid) bwd_phase_mel(q, t, s, b, gi, tid, NT);
      for (int tid = 0; tid < NT; ++tid) bwd_phase_gx(p, t, b, gi, tid, dst);
      for (int tid = 0; tid < NT; ++tid) bwd_phase_pack(p, t, gi, tid, dst, src);
      Ns = 1;
      for (int ps = 0; ps < plan.n_pass; ++ps) {
        for (int tid = 0; tid < NT; ++tid) bwd_pass_inv_any(plan.radix[ps], p, t, gi, tid, Ns, p.tw + plan.tw_off[ps], src, dst);
        Ns *= plan.radix[ps];
        cf32* tmp = src; src = dst; dst = tmp;
      }
      for (int slot = 0; slot < p.G; ++slot)
        for (int tid = 0; tid < NT; ++tid) bwd_phase_accumulate(p, t, s, b, gi, slot, tid, NT, src);
    }
    for (int tid = 0; tid < NT; ++tid) bwd_phase_scatter(q, t, b, tid, NT);
  }
  return 0;
}

int emu_trim(const float* y, int64_t L, double top_db, double ref_value, int32_t frame_length, int32_t hop,
             int64_t* out2) {
  const int T = (int)(1 + L / hop);
  std::vector<double> pw(T);
  double mx = 0.0;
  for (int f = 0; f < T; ++f) {
    double acc = 0.0;
    for (int l = 0; l < 32; ++l) acc += trim_frame_power_part(y, (int)L, frame_length, hop, f, l, 32);
    pw[f] = acc / frame_length;
    if (pw[f] > mx) mx = pw[f];
  }
  double ref_power = ref_value * ref_value;
  if (ref_value <= 0.0) { const double r = std::sqrt(mx); ref_power = r * r; }
  int first = 0x7fffffff, last = -1;
  for (int f = 0; f < T; ++f) {
    const float rms32 = sqrtf((float)pw[f]);
    if (trim_is_loud((double)rms32 * (double)rms32, ref_power, top_db)) { if (f < first) first = f; if (f > last) last = f; }
  }
  out2[0] = 0; out2[1] = 0;
  if (last >= 0) { out2[0] = (int64_t)first * hop; out2[1] = (int64_t)(last + 1) * hop; if (out2[1] > L) out2[1] = L; }
  return 0;
}

int emu_prior_interp(int32_t text_len, int32_t mel_len, int32_t round_mel, int32_t round_text, float* out) {
  std::vector<double> lf(8192);
  for (size_t i = 0; i < lf.size(); ++i) lf[i] = std::lgamma((double)i + 1.0);
  PriorParams p;
  memset(&p, 0, sizeof(p));
  p.lf = lf.data(); p.lf_n = (int)lf.size(); p.scaling = 1.0;
  const int w = mel_len, h = text_len, bw = prior_round(w, round_mel), bh = prior_round(h, round_text);
  for (int i = 0; i < w; ++i)
    for (int j = 0; j < h; ++j) out[(size_t)i * h + j] = prior_interp_value(p, w, h, bw, bh, i, j);
  return 0;
}

}  // extern "C"
