// Shared-memory Stockham autosort FFT building blocks (radix 2/4/8 butterflies in registers).
//
// One pass with current sub-transform size Ns and radix R, for butterfly index j in [0, M/R):
//     k  = j mod Ns
//     v[r] = in[j + r*M/R] * W_{Ns*R}^{r*k}            r = 0..R-1
//     y    = DFT_R(v)
//     out[(j-k)*R + k + q*Ns] = y[q]                    q = 0..R-1
// After passes whose radices multiply to M the output is in natural order.
//
// Bank conflicts: the Ns = 1 pass stores with a stride of R elements, which lands every lane of a
// warp on the same banks.  All buffers are therefore addressed through pidx(m) = m + (m >> 3)
// (one pad element per 8), which makes every access pattern of every pass conflict-free for 8-byte
// and 16-byte elements.  Twiddles come from per-pass tables laid out [r-1][k] (k contiguous across
// lanes), built on the host in float64 and read through L1 (read-only path).
#pragma once
#include "common.cuh"

namespace roar {

HD int pidx(int m) { return m + (m >> 3); }

template <class C> HD C ld_ro(const C* p) { return *p; }
#if defined(__CUDA_ARCH__)
template <> __device__ __forceinline__ cf32 ld_ro<cf32>(const cf32* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  cf32 r; r.x = v.x; r.y = v.y; return r;
}
#endif

template <bool INV, class C> HD void dft2(C* v) {
  C a = v[0];
  v[0] = cadd(a, v[1]);
  v[1] = csub(a, v[1]);
}

template <bool INV, class C> HD void dft4(C* v) {
  C t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
  C t2 = cadd(v[1], v[3]), t3 = cmul_mi<INV>(csub(v[1], v[3]));
  v[0] = cadd(t0, t2);
  v[2] = csub(t0, t2);
  v[1] = cadd(t1, t3);
  v[3] = csub(t1, t3);
}

template <bool INV, class C> HD void dft8(C* v) {
  typedef typename real_of<C>::type T;
  C e[4] = {v[0], v[2], v[4], v[6]};
  C o[4] = {v[1], v[3], v[5], v[7]};
  dft4<INV>(e);
  dft4<INV>(o);
  const T h = (T)0.70710678118654752440;
  // W8^1 = (1 -/+ i)/sqrt2, W8^2 = -/+ i, W8^3 = (-1 -/+ i)/sqrt2  (forward / inverse)
  C w1, w3;
  if (INV) {
    w1.x = (o[1].x - o[1].y) * h; w1.y = (o[1].x + o[1].y) * h;
    w3.x = (-o[3].x - o[3].y) * h; w3.y = (o[3].x - o[3].y) * h;
  } else {
    w1.x = (o[1].x + o[1].y) * h; w1.y = (o[1].y - o[1].x) * h;
    w3.x = (o[3].y - o[3].x) * h; w3.y = (-o[3].x - o[3].y) * h;
  }
  C w2 = cmul_mi<INV>(o[2]);
  v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
  v[1] = cadd(e[1], w1);   v[5] = csub(e[1], w1);
  v[2] = cadd(e[2], w2);   v[6] = csub(e[2], w2);
  v[3] = cadd(e[3], w3);   v[7] = csub(e[3], w3);
}

template <int R, bool INV, class C> HD void dftR(C* v) {
  if (R == 2) dft2<INV>(v);
  else if (R == 4) dft4<INV>(v);
  else dft8<INV>(v);
}

// radix plan for a complex FFT of size M = 2^lg: as many radix-8 passes as possible, then 4 or 2;
// tw_off[p] = offset of pass p's twiddle block ((R-1)*Ns entries, [r-1][k]) in the per-pass table
struct FftPlan {
  int n_pass;
  int radix[6];
  int ns[6];
  int tw_off[6];
  int tw_total;
};
HD FftPlan make_plan(int M) {
  FftPlan p;
  p.n_pass = 0;
  int lg = 0;
  while ((1 << lg) < M) ++lg;
  while (lg >= 3) { p.radix[p.n_pass++] = 8; lg -= 3; }
  if (lg == 2) p.radix[p.n_pass++] = 4;
  if (lg == 1) p.radix[p.n_pass++] = 2;
  int ns = 1, off = 0;
  for (int i = 0; i < p.n_pass; ++i) {
    p.ns[i] = ns;
    p.tw_off[i] = off;
    if (i > 0) off += (p.radix[i] - 1) * ns;
    ns *= p.radix[i];
  }
  p.tw_total = off;
  return p;
}

// tw_total of make_plan(M) without materialising the plan
HD int fft_tw_total(int M) {
  int lg = 0;
  while ((1 << lg) < M) ++lg;
  int ns = 1, off = 0, first = 1;
  while (lg > 0) {
    const int r = lg >= 3 ? 8 : (lg == 2 ? 4 : 2);
    if (!first) off += (r - 1) * ns;
    first = 0; ns *= r; lg -= lg >= 3 ? 3 : lg;
  }
  return off;
}

// twiddle (from the pass table `twp`, forward sign; conjugated for the inverse) + butterfly
template <int R, bool INV, class C, bool RO = true>
HD void stockham_twiddle_dft(C* v, int Ns, int j, const C* twp) {
  if (Ns > 1) {
    const int k = j & (Ns - 1);
#pragma unroll
    for (int r = 1; r < R; ++r) {
      C w = RO ? ld_ro(twp + (r - 1) * Ns + k) : twp[(r - 1) * Ns + k];   // RO: global table through L1
      if (INV) w.y = -w.y;
      v[r] = cmul(v[r], w);
    }
  }
  dftR<R, INV>(v);
}

template <int R, class C> HD void stockham_load(C* v, const C* in, int M, int j) {
  const int stride = M / R;
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = in[pidx(j + r * stride)];
}

template <int R, class C> HD void stockham_store(const C* v, C* out, int Ns, int j) {
  const int k = j & (Ns - 1);
  const int j0 = (j - k) * R + k;
#pragma unroll
  for (int q = 0; q < R; ++q) out[pidx(j0 + q * Ns)] = v[q];
}

}  // namespace roar
