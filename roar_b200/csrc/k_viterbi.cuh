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
};

struct VitBest { double v; int k; };
HD bool vit_better(double v, int k, const VitBest& b) { return v > b.v || (v == b.v && k < b.k); }

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
    if (t > 0) s = p.ptr[(size_t)(f0 + t) * (2 * p.npb) + s];
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
