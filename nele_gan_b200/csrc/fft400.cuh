// 400-point complex FFT for one warp, 400 = 16 x 25 (Cooley-Tukey), in place on one 400-element buffer:
//   n = 25 n1 + n2,  k = k1 + 16 k2
//   X[k1 + 16 k2] = sum_n2 W400^(n2 k1) [ sum_n1 x[25 n1 + n2] W16^(n1 k1) ] W25^(n2 k2)
// phase A: lanes 0..24 each run one 16-point FFT (fixed n2) in registers, apply the W400 twiddle and store
//          element (k1, n2) where they read element (n1 = k1, n2): position 25 k1 + n2;
// phase B: lanes 0..15 each run one 25-point DFT (5 x 5, fixed k1) in registers over the contiguous row
//          25 k1 .. 25 k1 + 24 and leave X[k1 + 16 k2] at position 25 k1 + k2 -- fft400_pos(k) below.
// Neither phase writes a position another lane reads, so one buffer per warp is enough (half the shared memory
// of a ping-pong pair: three CTAs per SM instead of two in siib_spec_kernel).  The caller puts a __syncwarp()
// between the phases.  Used by SIIB's 400/200 STFT (intel.py:52-54, scipy.fftpack.fft(n=400)).  The per-lane
// functions are __host__ __device__: tests/host_emul/fft400_emul.cpp checks them with g++ against a direct DFT.
#pragma once
#include "common.cuh"

namespace nele {

struct cpx {
  float x, y;
};
NELE_HD cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
NELE_HD cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
NELE_HD cpx cmulc(cpx a, cpx w) { return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }
NELE_HD cpx cmul_negi(cpx a) { return {a.y, -a.x}; }  // a * (-i)

// 5-point DFT, in place on v[0..4] (forward, e^{-2 pi i nk/5})
NELE_HD void dft5(cpx& v0, cpx& v1, cpx& v2, cpx& v3, cpx& v4) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;  // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;   // sin(2pi/5), sin(4pi/5)
  const cpx a14 = cadd(v1, v4), d14 = csub(v1, v4), a23 = cadd(v2, v3), d23 = csub(v2, v3);
  const cpx r0 = {v0.x + a14.x + a23.x, v0.y + a14.y + a23.y};
  const cpx p1 = {v0.x + c1 * a14.x + c2 * a23.x, v0.y + c1 * a14.y + c2 * a23.y};
  const cpx p2 = {v0.x + c2 * a14.x + c1 * a23.x, v0.y + c2 * a14.y + c1 * a23.y};
  // q1 = -i (s1 d14 + s2 d23), q2 = -i (s2 d14 - s1 d23)
  const cpx q1 = {s1 * d14.y + s2 * d23.y, -(s1 * d14.x + s2 * d23.x)};
  const cpx q2 = {s2 * d14.y - s1 * d23.y, -(s2 * d14.x - s1 * d23.x)};
  v0 = r0;
  v1 = cadd(p1, q1);
  v4 = csub(p1, q1);
  v2 = cadd(p2, q2);
  v3 = csub(p2, q2);
}

// position of X[k] in the buffer after phase B
NELE_HD int fft400_pos(int k) { return (k & 15) * 25 + (k >> 4); }

// a * W16^m for a compile-time m in 0..7: no table look-up, and no multiply at all for m = 0, 4
template <int M>
NELE_HD cpx cmul_w16(cpx a) {
  constexpr float h = 0.70710678118654752f, c = 0.92387953251128674f, s = 0.38268343236508977f;
  if (M == 0) return a;
  if (M == 4) return cmul_negi(a);
  if (M == 2) return {h * (a.x + a.y), h * (a.y - a.x)};   // (1 - i) / sqrt 2
  if (M == 6) return {h * (a.y - a.x), -h * (a.x + a.y)};  // (-1 - i) / sqrt 2
  if (M == 1) return cmulc(a, {c, -s});
  if (M == 3) return cmulc(a, {s, -c});
  if (M == 5) return cmulc(a, {-s, -c});
  return cmulc(a, {-c, -s});
}

template <int S, int J>
NELE_HD void fft16_butterfly(cpx (&a)[16]) {
  constexpr int half = 1 << S, pos = J & (half - 1), i0 = ((J >> S) << (S + 1)) + pos, i1 = i0 + half;
  const cpx u = cmul_w16<(pos << (3 - S))>(a[i1]);  // W16^(pos * 8 / half)
  const cpx v = a[i0];
  a[i0] = cadd(v, u);
  a[i1] = csub(v, u);
}
template <int S>
NELE_HD void fft16_stage(cpx (&a)[16]) {
  fft16_butterfly<S, 0>(a);
  fft16_butterfly<S, 1>(a);
  fft16_butterfly<S, 2>(a);
  fft16_butterfly<S, 3>(a);
  fft16_butterfly<S, 4>(a);
  fft16_butterfly<S, 5>(a);
  fft16_butterfly<S, 6>(a);
  fft16_butterfly<S, 7>(a);
}

// Twiddle tables (shared memory), both cut from w[k] = exp(-2 pi i k / 400):
//   twa[25 k1 + n2] = w[n2 k1]   phase A: lane n2 reads consecutive words for a fixed k1 (indexing w[n2 k1] directly
//                                 costs up to 13 shared-memory wavefronts per load: the stride 2 k1 words folds the 25
//                                 lanes onto few banks; siib_spec_kernel is bound by the shared-memory pipe)
//   tw25[m]         = w[16 m]     phase B: the 25th roots of unity, one broadcast read each
// phase A for lane n2 (< 25), in place on z[400]
NELE_HD void fft400_phase_a(int n2, cpx* z, const cpx* twa) {
  cpx a[16];
  // bit-reversed load for radix-2 decimation in time
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int br = ((n1 & 1) << 3) | ((n1 & 2) << 1) | ((n1 & 4) >> 1) | ((n1 & 8) >> 3);
    a[br] = z[25 * n1 + n2];
  }
  fft16_stage<0>(a);
  fft16_stage<1>(a);
  fft16_stage<2>(a);
  fft16_stage<3>(a);
  z[n2] = a[0];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) z[k1 * 25 + n2] = cmulc(a[k1], twa[k1 * 25 + n2]);
}

// phase B for lane k1 (< 16), in place on row k1 of z: X[k1 + 16 k2] ends at z[25 k1 + k2]
NELE_HD void fft400_phase_b(int k1, cpx* z, const cpx* tw25) {
  cpx v[25];
  cpx* row = z + k1 * 25;
#pragma unroll
  for (int n = 0; n < 25; ++n) v[n] = row[n];
  // 25 = 5 x 5: n = 5 na + nb, k = ka + 5 kb
#pragma unroll
  for (int nb = 0; nb < 5; ++nb) {
    dft5(v[nb], v[5 + nb], v[10 + nb], v[15 + nb], v[20 + nb]);  // index 5 ka + nb now holds ka
    if (nb > 0) {
#pragma unroll
      for (int ka = 1; ka < 5; ++ka) v[5 * ka + nb] = cmulc(v[5 * ka + nb], tw25[(nb * ka) % 25]);
    }
  }
#pragma unroll
  for (int ka = 0; ka < 5; ++ka) {
    dft5(v[5 * ka], v[5 * ka + 1], v[5 * ka + 2], v[5 * ka + 3], v[5 * ka + 4]);  // index 5 ka + kb
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) row[ka + 5 * kb] = v[5 * ka + kb];
  }
}

}  // namespace nele
