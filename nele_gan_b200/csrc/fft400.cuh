// 400-point complex FFT for one warp, 400 = 16 x 25 (Cooley-Tukey):
//   n = 25 n1 + n2,  k = k1 + 16 k2
//   X[k1 + 16 k2] = sum_n2 W400^(n2 k1) [ sum_n1 x[25 n1 + n2] W16^(n1 k1) ] W25^(n2 k2)
// phase A: lanes 0..24 each run one 16-point FFT (fixed n2) in registers, apply the
//          W400 twiddle and store t[k1][n2];
// phase B: lanes 0..15 each run one 25-point DFT (5 x 5, fixed k1) in registers.
// The caller puts a __syncwarp() between the phases.  Used by SIIB's 400/200 STFT
// (intel.py:52-54, scipy.fftpack.fft(n=400)).  The per-lane functions are
// __host__ __device__ so tests/host_emul can check them with g++.
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

// tw400[k] = exp(-2 pi i k / 400), k = 0..399 (shared or global memory)
// phase A for lane n2 (< 25): z = input [400], t = scratch [400] laid out t[k1 * 25 + n2]
NELE_HD void fft400_phase_a(int n2, const cpx* z, cpx* t, const cpx* tw400) {
  cpx a[16];
  // bit-reversed load for radix-2 decimation in time
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int br = ((n1 & 1) << 3) | ((n1 & 2) << 1) | ((n1 & 4) >> 1) | ((n1 & 8) >> 3);
    a[br] = z[25 * n1 + n2];
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int half = 1 << s;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
      // W16^(pos * 8 / half) = tw400[25 * pos * (8 / half)]
      const cpx w = tw400[25 * (pos << (3 - s))];
      const cpx u = cmulc(a[i1], w);
      const cpx v = a[i0];
      a[i0] = cadd(v, u);
      a[i1] = csub(v, u);
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) t[k1 * 25 + n2] = cmulc(a[k1], tw400[(n2 * k1) % 400]);
}

// phase B for lane k1 (< 16): t = phase-A output, out[k1 + 16 k2], k2 = 0..24
NELE_HD void fft400_phase_b(int k1, const cpx* t, cpx* out, const cpx* tw400) {
  cpx v[25];
#pragma unroll
  for (int n = 0; n < 25; ++n) v[n] = t[k1 * 25 + n];
  // 25 = 5 x 5: n = 5 na + nb, k = ka + 5 kb
#pragma unroll
  for (int nb = 0; nb < 5; ++nb) {
    dft5(v[nb], v[5 + nb], v[10 + nb], v[15 + nb], v[20 + nb]);  // index 5 ka + nb now holds ka
#pragma unroll
    for (int ka = 1; ka < 5; ++ka) v[5 * ka + nb] = cmulc(v[5 * ka + nb], tw400[16 * ((nb * ka) % 25)]);
  }
#pragma unroll
  for (int ka = 0; ka < 5; ++ka) {
    dft5(v[5 * ka], v[5 * ka + 1], v[5 * ka + 2], v[5 * ka + 3], v[5 * ka + 4]);  // index 5 ka + kb
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) out[k1 + 16 * (ka + 5 * kb)] = v[5 * ka + kb];
  }
}

}  // namespace nele
