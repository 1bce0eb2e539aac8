// K3  pYIN Viterbi over the 2*npb-state pitch/voicing HMM, one CTA per utterance.
//
// Replaces librosa.sequence.viterbi (dense 2npb x 2npb max-plus per step on the CPU) as used by
// librosa.pyin, called from roar/collections/tts/data/dataset.py:696-703.
//
// Exactly the dense recursion  V[t,j] = lp[t,j] + max_k (V[t-1,k] + lt[k,j])  with first-index
// argmax, evaluated without the dense matrix:
//   * in-band predecessors (|bin(k)-bin(j)| <= hw, both voicing blocks) use the banded float64
//     log-transition rows built on the host (de-duplicated bitwise, staged in shared memory);
//   * every out-of-band transition is the constant lt0 = log(0 + tiny).  If the first global argmax
//     k* of V[t-1] is in band for j, no out-of-band k can win (V[k]+lt0 <= V[k*]+lt0 < V[k*]+lt[k*,j]);
//     otherwise the best out-of-band predecessor IS k*.  So one block-wide (max, first index)
//     reduction per step covers them.
// Thread j owns pitch bin j in both voicing blocks: 4 independent running maxima
// (voiced/unvoiced source x voiced/unvoiced destination).  Observations are sparse: a frame has a
// handful of candidate bins (log-prob list from K2b), every other voiced bin is log(tiny).
// Back-pointers (uint16) go to HBM scratch; K3b walks them backwards, one thread per utterance, and
// emits f0 / voiced flag.
#pragma once
#include "common.cuh"

namespace roar {

constexpr int VIT_TW = 51;   // fast-path transition width (k_pyin_viterbi51)
constexpr int VIT_HW = 25;

struct VitParams {
  const int64_t* frame_off;     // [n_utts+1]
  const int32_t* order;         // [n_utts] utterances by decreasing length
  int32_t n_utts;
  int32_t npb, tw, hw, kmax, n_rows;
  const double* lt_rows;        // [n_rows][tw][2]  (same voicing, switch)
  const uint16_t* row_id;       // [npb]
  double lt0, li_voiced, li_unvoiced;
  const uint16_t* cand_bin;     // [frames][kmax]
  const double* cand_lp;
  const int32_t* n_cand;
  const double* lp_unvoiced;
  uint16_t* ptr;                // [frames][2*npb]
  int32_t* last_state;          // [n_utts]
  const double* freqs;          // [npb]
  float* f0;                    // [frames]
  float* voiced_flag;           // [frames]
  int32_t lt_in_smem;
  int32_t ptr_stride, ptr_uoff;   // back-pointer row layout: [voiced | unvoiced at ptr_uoff], row stride
  double lt_max;                // largest banded table entry
  // same-voicing entries of the representative interior row: kernel-parameter (constant bank) operands
  // of the uniform band scan, valid while vmax <= uniform_vmax (tables.hpp make_uniform_row)
  double uniform_vmax;
  double ltu[VIT_TW];
  // smallest (same - switch) difference over the banded table entries (~ ln 99), minus a rounding margin: a voiced
  // source whose value is below its unvoiced twin's by less than this never beats the twin at an unvoiced destination
  double twin_gap;
  // 1 when the centre entry of the uniform row is its strict maximum by > 1e-3 (tables.hpp): a warp whose sources in
  // reach all hold the same unvoiced value X >= -1e12 then knows its band-scan result without scanning (rule 10)
  int32_t flat_ok;
};

// One DP step for pitch bin j.  V: previous values as (voiced, unvoiced) pairs; returns the new pair,
// writes the two back-pointers.  (kstar, vmax) = first global argmax of the previous values.
// `row_ofs[i]` = row_id[i] * tw (entry offset of source bin i's row in the banded table).
HD void vit_step_bin(const VitParams& p, int j, const cf64* V, const double* lt, const int32_t* row_ofs,
                     double lp_v, double lp_u, int kstar, double vmax, cf64* vnew, uint16_t* ptr_row) {
  const int lo = j - p.hw < 0 ? 0 : j - p.hw;
  const int hi = j + p.hw > p.npb - 1 ? p.npb - 1 : j + p.hw;
  double b00 = -1e308, b10 = -1e308, b01 = -1e308, b11 = -1e308;   // src block -> dst block
  int a00 = 0, a10 = 0, a01 = 0, a11 = 0;
  const cf64* lt2 = reinterpret_cast<const cf64*>(lt) + (j + p.hw);
  for (int i = lo; i <= hi; ++i) {
    const cf64 v = V[i];
    const cf64 e = lt2[row_ofs[i] - i];
    const double ls = e.x, lc = e.y;
    const double s00 = v.x + ls, s10 = v.y + lc, s01 = v.x + lc, s11 = v.y + ls;
    if (s00 > b00) { b00 = s00; a00 = i; }
    if (s10 > b10) { b10 = s10; a10 = i; }
    if (s01 > b01) { b01 = s01; a01 = i; }
    if (s11 > b11) { b11 = s11; a11 = i; }
  }
  // voiced destination: voiced sources come first in state order
  double bv = b00; int av = a00;
  if (b10 > bv) { bv = b10; av = p.npb + a10; }
  double bu = b01; int au = a01;
  if (b11 > bu) { bu = b11; au = p.npb + a11; }
  const int ks = kstar >= p.npb ? kstar - p.npb : kstar;
  const int dist = ks > j ? ks - j : j - ks;
  if (dist > p.hw) {
    const double so = vmax + p.lt0;
    if (so > bv || (so == bv && kstar < av)) { bv = so; av = kstar; }
    if (so > bu || (so == bu && kstar < au)) { bu = so; au = kstar; }
  }
  vnew->x = lp_v + bv;
  vnew->y = lp_u + bu;
  ptr_row[j] = (uint16_t)av;
  ptr_row[p.npb + j] = (uint16_t)au;
}

#ifdef __CUDACC__
__device__ __forceinline__ void vit_warp_argmax(double& v, int& k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int ok = __shfl_xor_sync(0xffffffffu, k, o);
    if (ov > v || (ov == v && ok < k)) { v = ov; k = ok; }
  }
}

// blockDim.x = npb rounded up to a warp multiple
template <bool LT_SMEM, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_pyin_viterbi(const VitParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int utt = p.order[blockIdx.x];
  const int64_t f0 = p.frame_off[utt];
  const int T = (int)(p.frame_off[utt + 1] - f0);
  if (T <= 0) return;
  // carve
  size_t o = 0;
  cf64* Vb = (cf64*)(smem_raw + o);          o += sizeof(cf64) * 2 * (size_t)p.npb;
  double* lpv = (double*)(smem_raw + o);     o += sizeof(double) * 2 * (size_t)p.npb;
  double* wv = (double*)(smem_raw + o);      o += sizeof(double) * 64;   // [2][32] double-buffered
  int* wk = (int*)(smem_raw + o);            o += sizeof(int) * 64;
  int32_t* rid = (int32_t*)(smem_raw + o);   o += (sizeof(int32_t) * (size_t)p.npb + 15) & ~(size_t)15;
  double* lts = (double*)(smem_raw + o);
  if (LT_SMEM) {
    const int n = p.n_rows * p.tw * 2;
    for (int i = tid; i < n; i += nthr) lts[i] = p.lt_rows[i];
  }
  const double* lt = LT_SMEM ? lts : p.lt_rows;
  for (int i = tid; i < p.npb; i += nthr) { rid[i] = (int32_t)p.row_id[i] * p.tw; lpv[i] = p.lt0; lpv[p.npb + i] = p.lt0; }
  __syncthreads();
  // frame 0 observations
  {
    const int nc = p.n_cand[f0];
    if (tid < nc) lpv[p.cand_bin[(size_t)f0 * p.kmax + tid]] = p.cand_lp[(size_t)f0 * p.kmax + tid];
  }
  __syncthreads();
  const int j = tid;
  double bestv = -1e308; int bestk = 0x7fffffff;
  if (j < p.npb) {
    cf64 v;
    v.x = lpv[j] + p.li_voiced;
    v.y = p.lp_unvoiced[f0] + p.li_unvoiced;
    lpv[j] = p.lt0;
    Vb[j] = v;
    bestv = v.x; bestk = j;
    if (v.y > bestv) { bestv = v.y; bestk = p.npb + j; }
  }
  if (T > 1) {
    const int nc = p.n_cand[f0 + 1];
    if (tid < nc) lpv[p.npb + p.cand_bin[(size_t)(f0 + 1) * p.kmax + tid]] = p.cand_lp[(size_t)(f0 + 1) * p.kmax + tid];
  }
  vit_warp_argmax(bestv, bestk);
  if (lane == 0) { wv[warp] = bestv; wk[warp] = bestk; }
  __syncthreads();

  for (int t = 1; t < T; ++t) {
    // block-wide first argmax of V[t-1]; per-warp partials are double-buffered on step parity so one
    // barrier per step suffices
    const int rb = ((t - 1) & 1) * 32, wb = (t & 1) * 32;
    double vmax = lane < nwarp ? wv[rb + lane] : -1e308;
    int kstar = lane < nwarp ? wk[rb + lane] : 0x7fffffff;
    vit_warp_argmax(vmax, kstar);
    const cf64* Vc = Vb + (size_t)((t - 1) & 1) * p.npb;
    cf64* Vn = Vb + (size_t)(t & 1) * p.npb;
    double* lpc = lpv + (size_t)(t & 1) * p.npb;
    double* lpn = lpv + (size_t)((t + 1) & 1) * p.npb;
    // prefetch next frame's sparse observations (consumed after the band loop)
    int nc_next = 0; unsigned nb_bin = 0; double nb_lp = 0.0;
    if (t + 1 < T) {
      nc_next = p.n_cand[f0 + t + 1];
      if (tid < p.kmax) {
        nb_bin = p.cand_bin[(size_t)(f0 + t + 1) * p.kmax + tid];
        nb_lp = p.cand_lp[(size_t)(f0 + t + 1) * p.kmax + tid];
      }
    }
    const double lp_u = p.lp_unvoiced[f0 + t];
    bestv = -1e308; bestk = 0x7fffffff;
    if (j < p.npb) {
      cf64 vn;
      vit_step_bin(p, j, Vc, lt, rid, lpc[j], lp_u, kstar, vmax, &vn, p.ptr + (size_t)(f0 + t) * (2 * p.npb));
      lpc[j] = p.lt0;
      Vn[j] = vn;
      bestv = vn.x; bestk = j;
      if (vn.y > bestv) { bestv = vn.y; bestk = p.npb + j; }
    }
    if (tid < nc_next) lpn[nb_bin] = nb_lp;
    vit_warp_argmax(bestv, bestk);
    if (lane == 0) { wv[wb + warp] = bestv; wk[wb + warp] = bestk; }
    __syncthreads();
  }
  if (warp == 0) {
    const int rb = ((T - 1) & 1) * 32;
    double vmax = lane < nwarp ? wv[rb + lane] : -1e308;
    int kstar = lane < nwarp ? wk[rb + lane] : 0x7fffffff;
    vit_warp_argmax(vmax, kstar);
    if (lane == 0) p.last_state[utt] = kstar;
  }
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// K3 fast path (transition width 51, the reference's geometry at hop = frame/4 and sr >= 22.05 kHz).
//
// Same recursion, same float64 adds, same first-index arg-max -- but sources that provably cannot win
// are not visited:
//   * DEAD sources.  Every destination j receives an offer >= fl(vmax + lt0) from the global arg-max
//     k* (in band: fl(vmax + lt[k*,j]) with lt >= lt0; out of band: fl(vmax + lt0)).  A source s with
//     fl(V[s] + lt_max) < fl(vmax + lt0) (lt_max = largest banded entry) offers fl(V[s] + lt) <=
//     fl(V[s] + lt_max) in band (rounding is monotone), strictly less: it never wins, never ties (out of
//     band only k* matters).  While step t computes V[t,.], each thread compares fl(new value + lt_max) with
//     thr' = fl(LB + lt0), LB <= vmax_t being the value of the best one-step continuation of k*_{t-1}
//     (known before the step), and appends the survivors to a per-voicing LIVE LIST.  thr' <= the true
//     threshold, so the list is a superset of the live states.  Step t+1 walks the list when it holds
//     <= VIT_LIST_MAX entries (voiced stretches: a handful of candidates), otherwise scans the band.
//   * DOMINATED voiced sources.  When the voiced list overflowed, voiced offers are bounded by
//     Wc = fl(vvmax + lt_max) (vvmax = exact max over voiced V[t-1,.], lt_max = largest table entry).
//     If the best offer so far (unvoiced band scan + out-of-band) is > Wc for both destination
//     voicings, the voiced band scan is skipped (unvoiced stretches); otherwise it runs.
//   * IRRELEVANT voiced states.  Bin j is a candidate of frame t when K2b listed it (non-zero
//     observation); every other voiced state has lp = lt0 = log(tiny).  Its value X = fl(lt0 + bv)
//     sits far below its unvoiced twin X' = fl(lp_u + bu): the offers that built bv and bu come from
//     the same sources through the (same, switch) pair of one table slot, so bv - bu <= ln 99, hence
//     X - X' <= lt0 - lp_u + ln 99.  As sources at t+1 the two states again use one table slot, so the
//     twin's offer is larger by >= lp_u - lt0 - 2 ln 99 for EVERY destination: strictly larger whenever
//     lp_u >= lt0 + 16 (true unless voiced_prob clipped to exactly 1).  Such a voiced state never wins,
//     never ties, is never the arg-max: its value and back-pointer cannot reach the decoded path.  In
//     those ("sparse") steps the bin threads compute the unvoiced destination only (a band scan that
//     reads one table component), the voiced row is filled with the sentinel, and the few candidate
//     bins (4 per frame on average) are evaluated one per warp, lanes striding the sources.  Frames
//     with lp_u < lt0 + 16 take the dense step (both voicings for every bin).
//   * DEAD SEGMENTS.  When the unvoiced live list overflowed the band is scanned -- but a whole warp may still
//     have nothing live in reach: per 32-bin segment the maximum of the new unvoiced values is kept, and a warp
//     whose three segments in reach (w-1, w, w+1) all lie below the threshold thr' of the step that created them
//     sees only dead sources there.  It skips the unvoiced band scan (sparse and dense steps, and the candidate
//     evaluation of bins in such a segment): in voiced stretches only the bins within 25 per frame of the track
//     are revived after a frame whose voiced probability clipped to 1 killed the unvoiced layer.
// All rules only skip sources that lose strictly, so the decoded path is bit-identical to the dense
// recursion (checked against it in tests/); values and back-pointers are identical for every state
// that is not irrelevant in the sense above.
//
// The band scan is fully unrolled (51 sources): V rows are padded by hw sentinels (-1e308) on both
// sides and a zero row is appended to the table for them, row ids of a thread's 51 sources are packed
// 5 x 6 bit per register once per utterance, so one scan step is LDS.64 + bit-field + LDS.128 + 2 DADD
// + 2 compare/select.
constexpr int VIT_LIST_MAX = 32;
constexpr int VIT_CHAINS = 1;
constexpr double VIT_NEG = -1e308;

struct alignas(16) VitLive { double v; int32_t kb; int32_t row; };

struct VitBest2 { double b; int a; };
HD void vit_offer(VitBest2& x, double s, int idx) {
  if (s > x.b || (s == x.b && idx < x.a)) { x.b = s; x.a = idx; }
}

// padded per-source row ids: pad[i + hw] = row of source bin i, `zero_row` for the hw sentinel sources on both sides
HD void vit_pad_rows(const uint16_t* row_id, int npb, int zero_row, uint8_t* pad) {
  for (int i = 0; i < npb + 2 * VIT_HW; ++i) {
    const int b = i - VIT_HW;
    pad[i] = (uint8_t)((b >= 0 && b < npb) ? row_id[b] : zero_row);
  }
}

// band scan over one source voicing.  Vp = &Vsrc_padded[j] (source i = j-hw+d sits at Vp[d]).
// same/swit: best offers for the destination of the same / the other voicing; d_* = winning d.
HD void vit_band_scan(const double* Vp, const cf64* lt2, const uint8_t* rowp, double* same_b, int* same_d,
                      double* swit_b, int* swit_d) {
  // VIT_CHAINS independent running maxima over consecutive source ranges, merged in source order with
  // a strict compare: same first-index arg-max as one chain, more instruction-level parallelism
  constexpr int NCH = VIT_CHAINS, PER = (VIT_TW + NCH - 1) / NCH;
  double b0[NCH], b1[NCH];
  int d0[NCH], d1[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { b0[c] = VIT_NEG; b1[c] = VIT_NEG; d0[c] = 0; d1[c] = 0; }
#pragma unroll
  for (int q = 0; q < PER; ++q) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int d = c * PER + q;
      if (d < VIT_TW) {
        const double v = Vp[d];
        const uint32_t row = rowp[d];
        const cf64 e = lt2[row * VIT_TW + (2 * VIT_HW - d)];
        const double s0 = v + e.x, s1 = v + e.y;
        if (s0 > b0[c]) { b0[c] = s0; d0[c] = d; }
        if (s1 > b1[c]) { b1[c] = s1; d1[c] = d; }
      }
    }
  }
#pragma unroll
  for (int c = 1; c < NCH; ++c) {
    if (b0[c] > b0[0]) { b0[0] = b0[c]; d0[0] = d0[c]; }
    if (b1[c] > b1[0]) { b1[0] = b1[c]; d1[0] = d1[c]; }
  }
  *same_b = b0[0]; *same_d = d0[0]; *swit_b = b1[0]; *swit_d = d1[0];
}

// one live-list entry against destination bin j: `same` gets e.v + ls, `swit` gets e.v + lc
HD void vit_list_offer(const VitLive& e, int state_base, const cf64* lt2, int j, VitBest2& same, VitBest2& swit) {
  const int dd = j - e.kb + VIT_HW;
  if ((unsigned)dd <= 2u * VIT_HW) {
    const cf64 l = lt2[e.row * VIT_TW + dd];
    vit_offer(same, e.v + l.x, state_base + e.kb);
    vit_offer(swit, e.v + l.y, state_base + e.kb);
  }
}

struct Vit3Step {
  // sources (time t-1)
  const double* Vv;       // padded rows: element i at [i + hw]
  const double* Vu;
  const VitLive* Lv; int nv;   // nv > VIT_LIST_MAX: overflow
  const VitLive* Lu; int nu;
  double vmax; int kstar;      // first global arg-max of V[t-1]
  double vvmax;                // max over the voiced V[t-1] a destination of this warp can see in band
  bool u_flat = false; double u_flat_val = 0.0;   // every unvoiced source in reach of this warp holds exactly this value
  unsigned lv_mask = 0xffffffffu;   // voiced live-list entries an unvoiced destination of this warp has to visit
  bool u_dead = false;         // every unvoiced source in reach of this warp's destinations is dead (below the
                               //   liveness threshold of the step that created it): its band need not be scanned
  // tables
  const cf64* lt2;             // [n_rows+1][51] (ls, lc), last row zeros
  double lt0, lt_max;
  int npb;
};

// new values and back-pointers of pitch bin j
HD void vit3_step_bin(const Vit3Step& c, int j, const uint8_t* rowp, double lp_v, double lp_u, double* out_v,
                      double* out_u, int* ptr_v, int* ptr_u) {
  VitBest2 bv, bu;                       // destination voiced / unvoiced
  bv.b = VIT_NEG; bv.a = 0x7fffffff; bu = bv;
  {
    const int ks = c.kstar >= c.npb ? c.kstar - c.npb : c.kstar;
    const int dist = ks > j ? ks - j : j - ks;
    if (dist > VIT_HW) { const double so = c.vmax + c.lt0; vit_offer(bv, so, c.kstar); vit_offer(bu, so, c.kstar); }
  }
  if (c.nu > VIT_LIST_MAX) {
    if (!c.u_dead) {
      double sb, wb; int sd, wd;
      vit_band_scan(c.Vu + j, c.lt2, rowp, &sb, &sd, &wb, &wd);
      vit_offer(bu, sb, c.npb + j - VIT_HW + sd);
      vit_offer(bv, wb, c.npb + j - VIT_HW + wd);
    }
  } else {
    for (int e = 0; e < c.nu; ++e) vit_list_offer(c.Lu[e], c.npb, c.lt2, j, bu, bv);
  }
  if (c.nv > VIT_LIST_MAX) {
    const double Wc = c.vvmax + c.lt_max;
    if (!(bv.b > Wc && bu.b > Wc)) {
      double sb, wb; int sd, wd;
      vit_band_scan(c.Vv + j, c.lt2, rowp, &sb, &sd, &wb, &wd);
      vit_offer(bv, sb, j - VIT_HW + sd);
      vit_offer(bu, wb, j - VIT_HW + wd);
    }
  } else {
    for (int e = 0; e < c.nv; ++e) vit_list_offer(c.Lv[e], 0, c.lt2, j, bv, bu);
  }
  *out_v = lp_v + bv.b; *out_u = lp_u + bu.b;
  *ptr_v = bv.a; *ptr_u = bu.a;
}

// light band scan: one table component (COMP 0 = same voicing, 1 = switch) -> one destination
template <int COMP>
HD void vit_band_scan1(const double* Vp, const cf64* lt2, const uint8_t* rowp, double* best, int* best_d) {
  double b = VIT_NEG;
  int bd = 0;
#pragma unroll
  for (int d = 0; d < VIT_TW; ++d) {
    const double v = Vp[d];
    const uint32_t row = rowp[d];
    const double* e = reinterpret_cast<const double*>(lt2 + row * VIT_TW + (2 * VIT_HW - d));
    const double s0 = v + e[COMP];
    if (s0 > b) { b = s0; bd = d; }
  }
  *best = b; *best_d = bd;
}

// the same scan for a warp whose 32 + 2 hw sources are all interior bins, while vmax <= uniform_vmax:
// one table row for everybody -> `row` is a kernel parameter, its entries immediate constant operands
HD void vit_band_scan1u(const double* Vp, const double* row, double* best, int* best_d) {
  double b = VIT_NEG;
  int bd = 0;
#pragma unroll
  for (int d = 0; d < VIT_TW; ++d) {
    const double s0 = Vp[d] + row[2 * VIT_HW - d];
    if (s0 > b) { b = s0; bd = d; }
  }
  *best = b; *best_d = bd;
}

HD void vit_list_offer1(const VitLive& e, int state_base, const cf64* lt2, int j, int comp, VitBest2& x) {
  const int dd = j - e.kb + VIT_HW;
  if ((unsigned)dd <= 2u * VIT_HW) {
    const double* l = reinterpret_cast<const double*>(lt2 + e.row * VIT_TW + dd);
    vit_offer(x, e.v + l[comp], state_base + e.kb);
  }
}

// sparse step, bin thread: the unvoiced destination of pitch bin j.  Part A visits the in-band sources
// (it needs nothing but the previous rows and lists, so it runs while one warp still reduces the global
// arg-max); part B adds the out-of-band offer of k*.  vit_offer breaks ties by state index, so the
// order in which offers arrive does not matter.
HD VitBest2 vit4_unvoiced_scan(const Vit3Step& c, int j, const uint8_t* rowp, const double* uniform_row) {
  VitBest2 bu;
  bu.b = VIT_NEG; bu.a = 0x7fffffff;
  if (c.nu > VIT_LIST_MAX) {
    if (!c.u_dead) {
      double sb; int sd;
      if (uniform_row && c.u_flat) { sb = c.u_flat_val + uniform_row[VIT_HW]; sd = VIT_HW; }   // rule 10: the centre tap wins
      else if (uniform_row) vit_band_scan1u(c.Vu + j, uniform_row, &sb, &sd);
      else vit_band_scan1<0>(c.Vu + j, c.lt2, rowp, &sb, &sd);
      vit_offer(bu, sb, c.npb + j - VIT_HW + sd);
    }
  } else {
    for (int e = 0; e < c.nu; ++e) vit_list_offer1(c.Lu[e], c.npb, c.lt2, j, 0, bu);
  }
  if (c.nv > VIT_LIST_MAX) {
    const double Wc = c.vvmax + c.lt_max;
    if (!(bu.b > Wc)) {
      double wb; int wd;
      vit_band_scan1<1>(c.Vv + j, c.lt2, rowp, &wb, &wd);
      vit_offer(bu, wb, j - VIT_HW + wd);
    }
  } else {
#if defined(__CUDA_ARCH__)
    for (unsigned mm = c.lv_mask; mm; mm &= mm - 1) vit_list_offer1(c.Lv[__ffs(mm) - 1], 0, c.lt2, j, 1, bu);
#else
    for (int e = 0; e < c.nv; ++e) if ((c.lv_mask >> e) & 1u) vit_list_offer1(c.Lv[e], 0, c.lt2, j, 1, bu);
#endif
  }
  return bu;
}
HD void vit4_unvoiced_finish(VitBest2 bu, int npb, double lt0, double vmax, int kstar, int j, double lp_u, double* out_u, int* ptr_u) {
  const int ks = kstar >= npb ? kstar - npb : kstar;
  const int dist = ks > j ? ks - j : j - ks;
  if (dist > VIT_HW) vit_offer(bu, vmax + lt0, kstar);
  *out_u = lp_u + bu.b; *ptr_u = bu.a;
}

// sparse step, candidate bin b (voiced destination): this lane's share of the offers.  `lane`/`nl`
// stride the sources (nl = 1 on the host); the caller reduces the partials to (max, smallest index).
HD VitBest2 vit4_cand_partial(const Vit3Step& c, const uint16_t* row_id, int b, int lane, int nl, bool u_dead_b = false) {
  VitBest2 x;
  x.b = VIT_NEG; x.a = 0x7fffffff;
  if (lane == 0) {
    const int ks = c.kstar >= c.npb ? c.kstar - c.npb : c.kstar;
    const int dist = ks > b ? ks - b : b - ks;
    if (dist > VIT_HW) vit_offer(x, c.vmax + c.lt0, c.kstar);
  }
  if (c.nu > VIT_LIST_MAX) {
    if (!u_dead_b)
      for (int d = lane; d < VIT_TW; d += nl) {
        const int k = b - VIT_HW + d;
        if (k >= 0 && k < c.npb) vit_offer(x, c.Vu[VIT_HW + k] + c.lt2[(int)row_id[k] * VIT_TW + (2 * VIT_HW - d)].y, c.npb + k);
      }
  } else {
    for (int e = lane; e < c.nu; e += nl) vit_list_offer1(c.Lu[e], c.npb, c.lt2, b, 1, x);
  }
  if (c.nv > VIT_LIST_MAX) {
    for (int d = lane; d < VIT_TW; d += nl) {
      const int k = b - VIT_HW + d;
      if (k >= 0 && k < c.npb) vit_offer(x, c.Vv[VIT_HW + k] + c.lt2[(int)row_id[k] * VIT_TW + (2 * VIT_HW - d)].x, k);
    }
  } else {
    for (int e = lane; e < c.nv; e += nl) vit_list_offer1(c.Lv[e], 0, c.lt2, b, 0, x);
  }
  return x;
}
// frames whose unvoiced observation is at least this far above log(tiny) take the sparse step
constexpr double VIT_SPARSE_MARGIN = 16.0;

// lower bound on vmax_t: the best one-step continuation of k*_{t-1} into the unvoiced state of its own
// bin or into one of frame t's candidate bins.  `lane`/`nl` stride the candidate list (nl = 1 on the host).
HD double vit3_lower_bound(const Vit3Step& c, const uint16_t* row_id, double lp_u, const uint16_t* cbin,
                           const double* clp, int nc, int lane, int nl) {
  const int ks = c.kstar >= c.npb ? c.kstar - c.npb : c.kstar;
  const bool kv = c.kstar < c.npb;
  const cf64* row = c.lt2 + (int)row_id[ks] * VIT_TW;
  double lb = VIT_NEG;
  if (lane == 0) { const cf64 e = row[VIT_HW]; lb = lp_u + (c.vmax + (kv ? e.y : e.x)); }
  // the first 32 candidates are enough for a lower bound (one per lane on the device: no loop)
  const int ncl = nc < 32 ? nc : 32;
  for (int q = lane; q < ncl; q += nl) {
    const int b = cbin[q];
    const int dd = b - ks + VIT_HW;
    double l = c.lt0;
    if ((unsigned)dd <= 2u * VIT_HW) { const cf64 e = row[dd]; l = kv ? e.x : e.y; }
    const double cand = clp[q] + (c.vmax + l);
    if (cand > lb) lb = cand;
    if (nl == 32) break;
  }
  return lb;
}

constexpr int VIT_STATS_N = 16;
#ifdef __CUDACC__
#ifdef ROAR_VIT_STATS
__device__ unsigned long long g_vit_stats[VIT_STATS_N];
#define VIT_STAT(i, v) do { if (tid == 0) atomicAdd(&g_vit_stats[i], (unsigned long long)(v)); } while (0)
#else
#define VIT_STAT(i, v) do { } while (0)
#endif
// Warp reductions over NEGATIVE finite doubles with redux.sync: for negative values, larger value <=>
// smaller (hi, lo) bit pattern as unsigned.  Every V of this HMM is < 0 (sums of logs of
// probabilities), and the fillers are -1e308.
__device__ __forceinline__ double vit_warp_max(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  return __hiloint2double((int)mhi, (int)mlo);
}
// minimum of negative finite doubles: smaller value <=> larger (hi, lo) bit pattern
__device__ __forceinline__ double vit_warp_min(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  return __hiloint2double((int)mhi, (int)mlo);
}
// (max value, smallest index attaining it)
__device__ __forceinline__ void vit_warp_argmax_neg(double& v, int& k) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  const bool is = hi == mhi && lo == mlo;
  k = (int)__reduce_min_sync(0xffffffffu, is ? (unsigned)k : 0x7fffffffu);
  v = __hiloint2double((int)mhi, (int)mlo);
}

// fixed shared-memory layout (compile-time offsets: nothing to re-derive per step)
constexpr int VIT_NPB_MAX = 608;
constexpr int VIT_VP_MAX = VIT_NPB_MAX + 2 * VIT_HW;
constexpr int VIT_KMAX_MAX = 352;
constexpr int VIT_ROWS_MAX = 64;
struct alignas(16) Vit3Shared {
  cf64 lt2[VIT_ROWS_MAX * VIT_TW];           // banded rows + one zero row
  double Vv[2][VIT_VP_MAX];                  // padded value rows, double-buffered on step parity
  double Vu[2][VIT_VP_MAX];
  double lpv[2][VIT_NPB_MAX];                // sparse voiced observations of the frame being entered
  double clp[2][VIT_KMAX_MAX];               // its candidate list (log-prob, bin)
  VitLive Lv[3][VIT_LIST_MAX];               // live lists: slot t % 3
  VitLive Lu[3][VIT_LIST_MAX];
  double wv[2][32], wvv[2][32];              // per-warp partials (= 32-bin segment stats)
  double wuu[2][32];                         // maximum of the segment's unvoiced values (dead-segment rule)
  double wul[2][32];                         // minimum of the segment's unvoiced values (flat-segment rule)
  int wk[2][32];
  int cnt[3][2];
  int pub_kstar, pub_doa;                    // pub_doa: dead-on-arrival dense step (see the kernel)
  double pub_vmax, pub_thr;                  // block-wide values of the running step (written by the lead warp)
  double nxt_lpu[2];                         // unvoiced observation / candidate count of frame f at [f & 1]
  int nxt_nc[2];                             //   (written one step ahead by the prefetch warp)
  uint16_t cbin[2][VIT_KMAX_MAX];
  uint16_t rowid[VIT_NPB_MAX];
  uint8_t rowpad[VIT_VP_MAX + 6];            // row id per padded source position (generic band scans)
};

// warp-aggregated append of the live states of one voicing
__device__ __forceinline__ void vit3_append(bool live, double v, int kb, int row, VitLive* L, int* cnt, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, live);
  if (m == 0) return;
  // already overflowed: the count only has to stay above VIT_LIST_MAX (the next step scans the band)
  if (*reinterpret_cast<volatile int*>(cnt) > VIT_LIST_MAX) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(cnt, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (live) {
    const int pos = base + __popc(m & ((1u << lane) - 1u));
    if (pos < VIT_LIST_MAX) { VitLive e; e.v = v; e.kb = kb; e.row = row; L[pos] = e; }
  }
}

// candidate bins of the frame being entered: one per warp, lanes stride the sources (sparse steps and
// dead-on-arrival dense steps)
#define VIT_CAND_LOOP \
      for (int q = warp; q < nc_cur; q += nwarp) { \
        const int b = s.cbin[wp][q]; \
        bool dead_b = false; \
        if (c.nu > VIT_LIST_MAX) { \
          const int sb = b >> 5; \
          double segu = s.wuu[rp][sb]; \
          if (sb > 0) { const double y = s.wuu[rp][sb - 1]; segu = y > segu ? y : segu; } \
          if (sb + 1 < nwarp) { const double y = s.wuu[rp][sb + 1]; segu = y > segu ? y : segu; } \
          dead_b = segu + p.lt_max < thr_prev; \
        } \
        VitBest2 x = vit4_cand_partial(c, s.rowid, b, lane, 32, dead_b); \
        vit_warp_argmax_neg(x.b, x.a); \
        if (lane == 0) { \
          const double cv = s.clp[wp][q] + x.b; \
          s.Vv[wp][VIT_HW + b] = cv; \
          pr[b] = (uint16_t)x.a; \
          if (cv > bestv || (cv == bestv && b < bestk)) { bestv = cv; bestk = b; } \
          if (cv + p.lt_max >= thr) { \
            const int pos = atomicAdd(&s.cnt[wl][0], 1); \
            if (pos < VIT_LIST_MAX) { VitLive e; e.v = cv; e.kb = b; e.row = s.rowid[b]; s.Lv[wl][pos] = e; } \
          } \
        } \
      }

// blockDim.x = npb rounded up to a warp multiple (<= 608); two CTAs (utterances) per SM
__global__ void __launch_bounds__(608, 2) k_pyin_viterbi51(const VitParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Vit3Shared& s = *reinterpret_cast<Vit3Shared*>(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int utt = p.order[blockIdx.x];
  const int64_t f0 = p.frame_off[utt];
  const int T = (int)(p.frame_off[utt + 1] - f0);
  if (T <= 0) return;
  const int npb = p.npb, kmax = p.kmax, zero_row = p.n_rows;
  {
    const int n = p.n_rows * VIT_TW;
    const cf64* src = reinterpret_cast<const cf64*>(p.lt_rows);
    for (int i = tid; i < n; i += nthr) s.lt2[i] = src[i];
    for (int i = tid; i < VIT_TW; i += nthr) { cf64 z; z.x = 0.0; z.y = 0.0; s.lt2[n + i] = z; }
    for (int i = tid; i < npb; i += nthr) { s.rowid[i] = p.row_id[i]; s.lpv[0][i] = p.lt0; s.lpv[1][i] = p.lt0; }
    for (int i = tid; i < VIT_VP_MAX; i += nthr) { s.Vv[0][i] = VIT_NEG; s.Vv[1][i] = VIT_NEG; s.Vu[0][i] = VIT_NEG; s.Vu[1][i] = VIT_NEG; }
    if (tid < 6) (&s.cnt[0][0])[tid] = tid < 2 ? VIT_LIST_MAX + 1 : 0;   // time 0: scan everything
  }
  const int j = tid;
  for (int i = tid; i < npb + 2 * VIT_HW; i += nthr) {
    const int b = i - VIT_HW;
    s.rowpad[i] = (uint8_t)((b >= 0 && b < npb) ? p.row_id[b] : zero_row);
  }
  const uint8_t* rid = s.rowpad + (j < npb ? j : 0);
  __syncthreads();
  // frame 0 observations
  int nc_cur = p.n_cand[f0];
  if (tid < nc_cur) s.lpv[0][p.cand_bin[(size_t)f0 * kmax + tid]] = p.cand_lp[(size_t)f0 * kmax + tid];
  __syncthreads();
  double bestv = VIT_NEG, vvb = VIT_NEG; int bestk = 0x7fffffff;
  if (j < npb) {
    const double vv = s.lpv[0][j] + p.li_voiced;
    const double vu = p.lp_unvoiced[f0] + p.li_unvoiced;
    s.lpv[0][j] = p.lt0;
    s.Vv[0][VIT_HW + j] = vv; s.Vu[0][VIT_HW + j] = vu;
    bestv = vv; bestk = j; vvb = vv;
    if (vu > bestv) { bestv = vu; bestk = npb + j; }
  }
  if (T > 1) {
    nc_cur = p.n_cand[f0 + 1];
    if (tid == 0) { s.nxt_nc[1] = nc_cur; s.nxt_lpu[1] = p.lp_unvoiced[f0 + 1]; }
    if (tid < nc_cur) {
      const uint16_t b = p.cand_bin[(size_t)(f0 + 1) * kmax + tid];
      const double l = p.cand_lp[(size_t)(f0 + 1) * kmax + tid];
      s.lpv[1][b] = l; s.cbin[1][tid] = b; s.clp[1][tid] = l;
    }
  }
  vit_warp_argmax_neg(bestv, bestk);
  vvb = vit_warp_max(vvb);
  if (lane == 0) { s.wv[0][warp] = bestv; s.wk[0][warp] = bestk; s.wvv[0][warp] = vvb; s.wuu[0][warp] = 0.0; s.wul[0][warp] = VIT_NEG; }   // 0.0: no bound yet (and not flat)
  __syncthreads();

  // One warp (`lead`, an interior one) reduces the block-wide quantities of a step -- first global
  // arg-max (k*, vmax), the lower bound and the liveness threshold -- and publishes them in shared
  // memory; it signals named barrier 1 (bar.arrive) and the other warps wait on it (bar.sync) only
  // after their in-band work, which does not depend on those values.
  const int lead = nwarp >> 1;
  const int pfw = lead > 0 ? lead - 1 : 0;            // prefetch warp
  double prev_vmax = 0.0;                              // vmax is strictly decreasing in t
  double thr_prev = VIT_NEG;                           // liveness threshold of the step that created the current sources
  uint16_t* pr = p.ptr + (size_t)(f0 + 1) * (2 * npb);
  for (int t = 1; t < T; ++t, pr += 2 * npb) {
    const int rp = (t - 1) & 1, wp = t & 1;           // V / partial parity: read, write
    const int rl = (t - 1) % 3, wl = t % 3, zl = (t + 1) % 3;   // live-list slots: read, write, reset
    Vit3Step c;
    c.Vv = s.Vv[rp]; c.Vu = s.Vu[rp];
    c.Lv = s.Lv[rl]; c.Lu = s.Lu[rl];
    c.nv = s.cnt[rl][0]; c.nu = s.cnt[rl][1];
    c.lt2 = s.lt2; c.lt0 = p.lt0; c.lt_max = p.lt_max; c.npb = npb;
    c.vmax = 0.0; c.kstar = 0; c.vvmax = 0.0;
    if (tid < 2) s.cnt[zl][tid] = 0;
    const double lp_u = s.nxt_lpu[wp];
    nc_cur = s.nxt_nc[wp];
    {   // dead-segment rule: the unvoiced values in reach of this warp's destinations against their creation threshold
      double segu = s.wuu[rp][warp];
      if (warp > 0) { const double x = s.wuu[rp][warp - 1]; segu = x > segu ? x : segu; }
      if (warp + 1 < nwarp) { const double x = s.wuu[rp][warp + 1]; segu = x > segu ? x : segu; }
      c.u_dead = segu + p.lt_max < thr_prev;
      // FLAT SEGMENTS (rule 10).  When the three segments in reach hold one and the same unvoiced value X (min ==
      // max, bitwise), every in-band offer to a destination of this warp is fl(X + row[d]) with the uniform row:
      // the largest is the centre tap's (strict row maximum by > 1e-3, far above an ulp of X >= -1e12), attained
      // there only -- value fl(X + row[hw]), arg-max the destination's own bin, exactly what the scan returns.
      // Far from the pitch track the unvoiced layer evolves identically in every bin, so this is the common case
      // (~45 % of the uniform band scans on the bench corpus).
      const double x = s.wuu[rp][warp];
      bool flat = p.flat_ok && warp > 0 && warp + 1 < nwarp && x == s.wul[rp][warp] && x >= -1e12;
      if (flat) flat = s.wuu[rp][warp - 1] == x && s.wul[rp][warp - 1] == x && s.wuu[rp][warp + 1] == x && s.wul[rp][warp + 1] == x;
      c.u_flat = flat; c.u_flat_val = x;
    }
    VIT_STAT(13, (c.nu > VIT_LIST_MAX) ? 1 : 0);
#ifdef ROAR_VIT_STATS
    if (lane == 0 && c.nu > VIT_LIST_MAX) atomicAdd(&g_vit_stats[c.u_dead ? 14 : 15], 1ull);
#endif
    double thr = 0.0;
    if (warp == lead) {
      double vmax = lane < nwarp ? s.wv[rp][lane] : VIT_NEG;
      int kstar = lane < nwarp ? s.wk[rp][lane] : 0x7fffffff;
      vit_warp_argmax_neg(vmax, kstar);
      c.vmax = vmax; c.kstar = kstar;
      // liveness threshold for the values this step produces
      double lb = vit3_lower_bound(c, s.rowid, lp_u, s.cbin[wp], s.clp[wp], nc_cur, lane, 32);
      lb = vit_warp_max(lb);
      thr = lb + p.lt0;
      // DEAD ON ARRIVAL.  No offer exceeds fl(vmax + lt_max), so a destination with observation lp ends at or
      // below U(lp) = fl(lp + fl(vmax + lt_max)); if fl(U + lt_max) < thr it is dead the moment it is created
      // (liveness test of the next step), i.e. neither its value nor its back-pointer can ever be read.  When
      // that holds for lp = lt0 (voiced states without a candidate) and for lp = lp_u (all unvoiced states) --
      // the frames whose voiced probability clipped to 1 -- a dense step only evaluates the candidate bins.
      const double top = vmax + p.lt_max;
      const int doa = (((p.lt0 + top) + p.lt_max) < thr && ((lp_u + top) + p.lt_max) < thr) ? 1 : 0;
      if (lane == 0) { s.pub_vmax = vmax; s.pub_thr = thr; s.pub_kstar = kstar; s.pub_doa = doa; }
      asm volatile("bar.arrive 1, %0;" ::"r"(nthr) : "memory");
    }
    // one warp prefetches the next frame's sparse observations for the block (loads issued here, stored
    // after the band work); lanes read list slots beyond n_cand too -- in bounds, discarded
    const bool pf = warp == pfw && t + 1 < T;
    int nc_next = 0; unsigned nb_bin = 0; double nb_lp = 0.0, nb_lpu = 0.0;
    if (pf) {
      const size_t fr = (size_t)(f0 + t + 1);
      nc_next = p.n_cand[fr];
      nb_lpu = p.lp_unvoiced[fr];
      if (lane < kmax) { nb_bin = p.cand_bin[fr * kmax + lane]; nb_lp = p.cand_lp[fr * kmax + lane]; }
    }
    bestv = VIT_NEG; bestk = 0x7fffffff; vvb = VIT_NEG;
    bool live_v = false, live_u = false;
    double nv = VIT_NEG, nu = VIT_NEG;
    const bool sparse = lp_u >= p.lt0 + VIT_SPARSE_MARGIN;     // block-uniform
    bool doa_step = false;
    VIT_STAT(sparse ? 0 : 1, 1);
    VIT_STAT(c.nu > VIT_LIST_MAX ? 2 : 4, 1);
    VIT_STAT(c.nv > VIT_LIST_MAX ? 3 : 5, 1);
    VIT_STAT(6, nc_cur);
    VIT_STAT(7, prev_vmax <= p.uniform_vmax ? 1 : 0);
    VIT_STAT(8, (!sparse && c.nu > VIT_LIST_MAX) ? 1 : 0);
    VIT_STAT(9, (!sparse && c.nv > VIT_LIST_MAX) ? 1 : 0);
    VIT_STAT(10, (sparse && c.nv > VIT_LIST_MAX) ? 1 : 0);
    if (sparse) {
      VitBest2 bu;
      bu.b = VIT_NEG; bu.a = 0x7fffffff;
      if (c.nv <= VIT_LIST_MAX) {
        // Which voiced live-list entries can matter for THIS warp's unvoiced destinations: (i) in reach of the
        // warp's 32 bins; (ii) not dominated by their unvoiced twin -- the unvoiced state of the same bin reaches
        // exactly the same destinations through the `same` component of the same table slot, which exceeds the
        // `switch` component by >= twin_gap, so a voiced source with V_v - V_u < twin_gap loses strictly at every
        // unvoiced destination.  (ii) needs the twin to be visited: only while the unvoiced band is scanned.
        bool need = false;
        if (lane < c.nv) {
          const VitLive e = c.Lv[lane];
          need = (unsigned)(e.kb - (32 * warp - VIT_HW)) <= (unsigned)(31 + 2 * VIT_HW);
          if (need && c.nu > VIT_LIST_MAX && !c.u_dead) need = !(e.v - c.Vu[VIT_HW + e.kb] < p.twin_gap);
        }
        c.lv_mask = __ballot_sync(0xffffffffu, need);
      }
      if (j < npb) {
        if (c.nv > VIT_LIST_MAX) c.vvmax = 0.0;   // (after a dense step the segment maxima would do; rare)
        // warp-uniform: every source of this warp's 32 destinations is an interior bin
        const bool uni = prev_vmax <= p.uniform_vmax && 32 * warp >= 2 * VIT_HW && 32 * warp + 31 + 2 * VIT_HW <= npb - 1;
        bu = vit4_unvoiced_scan(c, j, rid, uni ? p.ltu : nullptr);
      }
      if (warp != lead) {
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        c.vmax = s.pub_vmax; c.kstar = s.pub_kstar; thr = s.pub_thr;
      }
      VIT_CAND_LOOP
      if (j < npb) {
        int au;
        vit4_unvoiced_finish(bu, npb, p.lt0, c.vmax, c.kstar, j, lp_u, &nu, &au);
        if (s.lpv[wp][j] == p.lt0) s.Vv[wp][VIT_HW + j] = VIT_NEG; else s.lpv[wp][j] = p.lt0;
        s.Vu[wp][VIT_HW + j] = nu;
        pr[npb + j] = (uint16_t)au;
        if (nu > bestv || (nu == bestv && npb + j < bestk)) { bestv = nu; bestk = npb + j; }
        live_u = nu + p.lt_max >= thr;
      }
      vvb = 0.0;     // no per-segment voiced maxima in a sparse step: an overflowing voiced list is scanned
    } else {
      if (warp != lead) {
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        c.vmax = s.pub_vmax; c.kstar = s.pub_kstar; thr = s.pub_thr;
      }
      doa_step = s.pub_doa != 0;        // written before the lead's bar.arrive; the lead reads its own store
      if (doa_step) {
        VIT_STAT(11, 1);
        VIT_CAND_LOOP
        if (j < npb) {
          if (s.lpv[wp][j] == p.lt0) s.Vv[wp][VIT_HW + j] = VIT_NEG; else s.lpv[wp][j] = p.lt0;
          s.Vu[wp][VIT_HW + j] = VIT_NEG;
        }
        vvb = 0.0;
      } else if (j < npb) {
        // sources in band of this warp's destinations live in the 32-bin segments of warps w-1, w, w+1
        double seg = s.wvv[rp][warp];
        if (warp > 0) { const double x = s.wvv[rp][warp - 1]; if (x > seg) seg = x; }
        if (warp + 1 < nwarp) { const double x = s.wvv[rp][warp + 1]; if (x > seg) seg = x; }
        c.vvmax = seg;
        int av, au;
        vit3_step_bin(c, j, rid, s.lpv[wp][j], lp_u, &nv, &nu, &av, &au);
        s.lpv[wp][j] = p.lt0;
        s.Vv[wp][VIT_HW + j] = nv; s.Vu[wp][VIT_HW + j] = nu;
        pr[j] = (uint16_t)av; pr[npb + j] = (uint16_t)au;
        bestv = nv; bestk = j; vvb = nv;
        if (nu > bestv) { bestv = nu; bestk = npb + j; }
        live_v = nv + p.lt_max >= thr; live_u = nu + p.lt_max >= thr;
      }
    }
    prev_vmax = c.vmax;
    const int myrow = j < npb ? (int)s.rowid[j] : 0;
    if (!sparse && !doa_step) vit3_append(live_v, nv, j, myrow, s.Lv[wl], &s.cnt[wl][0], lane);
    if (!doa_step) vit3_append(live_u, nu, j, myrow, s.Lu[wl], &s.cnt[wl][1], lane);
    if (pf) {
      if (lane == 0) { s.nxt_nc[wp ^ 1] = nc_next; s.nxt_lpu[wp ^ 1] = nb_lpu; }
      if (lane < nc_next) { s.lpv[wp ^ 1][nb_bin] = nb_lp; s.cbin[wp ^ 1][lane] = (uint16_t)nb_bin; s.clp[wp ^ 1][lane] = nb_lp; }
      for (int q = 32 + lane; q < nc_next; q += 32) {       // long candidate lists (rare)
        const size_t fr = (size_t)(f0 + t + 1);
        const unsigned b = p.cand_bin[fr * kmax + q];
        const double l = p.cand_lp[fr * kmax + q];
        s.lpv[wp ^ 1][b] = l; s.cbin[wp ^ 1][q] = (uint16_t)b; s.clp[wp ^ 1][q] = l;
      }
    }
    // dead-on-arrival step: only lane 0 of a candidate warp holds a value, every unvoiced value is the sentinel
    if (!doa_step) vit_warp_argmax_neg(bestv, bestk);
    if (!sparse && !doa_step) vvb = vit_warp_max(vvb);     // sparse / dead-on-arrival: 0.0 everywhere
    const double uub = doa_step ? VIT_NEG : vit_warp_max(j < npb ? nu : VIT_NEG);
    // segment minimum (flat-segment rule); lanes beyond the last bin hold the maximum so that they never decide
    const double ulb = doa_step ? VIT_NEG : vit_warp_min(j < npb ? nu : uub);
    if (lane == 0) { s.wv[wp][warp] = bestv; s.wk[wp][warp] = bestk; s.wvv[wp][warp] = vvb; s.wuu[wp][warp] = uub; s.wul[wp][warp] = ulb; }
    thr_prev = thr;
    __syncthreads();
  }
  if (warp == 0) {
    const int rb = (T - 1) & 1;
    double vmax = lane < nwarp ? s.wv[rb][lane] : VIT_NEG;
    int kstar = lane < nwarp ? s.wk[rb][lane] : 0x7fffffff;
    vit_warp_argmax_neg(vmax, kstar);
    if (lane == 0) p.last_state[utt] = kstar;
  }
}

// K3b: back-track, one thread per utterance (independent latency chains run concurrently)
__global__ void k_pyin_backtrack(const VitParams p) {
  const int utt = blockIdx.x * blockDim.x + threadIdx.x;
  if (utt >= p.n_utts) return;
  const int64_t f0 = p.frame_off[utt];
  const int T = (int)(p.frame_off[utt + 1] - f0);
  if (T <= 0) return;
  int s = p.last_state[utt];
  for (int t = T - 1; t >= 0; --t) {
    const bool voiced = s < p.npb;
    p.f0[f0 + t] = voiced ? (float)p.freqs[s] : 0.f;
    p.voiced_flag[f0 + t] = voiced ? 1.f : 0.f;
    if (t > 0) s = p.ptr[(size_t)(f0 + t) * p.ptr_stride + (s < p.npb ? s : s - p.npb + p.ptr_uoff)];
  }
}

// counting sort of utterances by frame count, longest first -> order[]
__global__ void k_len_hist(const int64_t* frame_off, int32_t n_utts, int32_t max_T, int32_t* hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_utts) return;
  int T = (int)(frame_off[i + 1] - frame_off[i]);
  if (T > max_T) T = max_T;
  atomicAdd(&hist[max_T - T], 1);   // descending
}
// exclusive prefix sum in place, one CTA of 1024 threads, n <= 1024 * 1024
__global__ void k_len_scan(int32_t* hist, int32_t n) {
  __shared__ int32_t s_scan[1024];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b = tid * per, e = b + per < n ? b + per : n;
  int32_t tot = 0;
  for (int i = b; i < e; ++i) tot += hist[i];
  s_scan[tid] = tot;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    int32_t x = tid >= d ? s_scan[tid - d] : 0;
    __syncthreads();
    s_scan[tid] += x;
    __syncthreads();
  }
  int32_t acc = s_scan[tid] - tot;
  for (int i = b; i < e; ++i) { const int32_t v = hist[i]; hist[i] = acc; acc += v; }
}
__global__ void k_len_scatter(const int64_t* frame_off, int32_t n_utts, int32_t max_T, int32_t* cursor,
                              int32_t* order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_utts) return;
  int T = (int)(frame_off[i + 1] - frame_off[i]);
  if (T > max_T) T = max_T;
  order[atomicAdd(&cursor[max_T - T], 1)] = i;
}
#endif  // __CUDACC__

}  // namespace roar
