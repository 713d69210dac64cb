// Host-side (plain C++) construction of every constant table the kernels use.
// Each builder cites the reference lines (or the published algorithm of the
// un-vendored dependency) it restates.  No CUDA here, so the tables can also
// be built by the g++ emulation harness in tests/host_emul.
#pragma once
#include <math.h>
#include <stdint.h>

#include <numeric>
#include <vector>

#include "ear_core.cuh"

namespace nele {
namespace host {

constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------- HASPI bands
inline void center_freqs(double* cf, int nchan = kBands) {  // pyhaspi2.py:753-777 (shift branch dead)
  const double lo = 80.0, hi = 8000.0, earq = 9.26449, minbw = 24.7;
  std::vector<double> v(nchan);
  v[0] = hi;
  for (int k = 1; k < nchan; ++k)
    v[k] = -(earq * minbw) +
           exp(k * (-log(hi + earq * minbw) + log(lo + earq * minbw)) / (nchan - 1)) * (hi + earq * minbw);
  for (int k = 0; k < nchan; ++k) cf[k] = v[nchan - 1 - k];
}

inline double interp1(double x, const double* xp, const double* fp, int n) {  // np.interp
  if (x <= xp[0]) return fp[0];
  if (x >= xp[n - 1]) return fp[n - 1];
  int j = 0;
  while (j < n - 2 && x >= xp[j + 1]) ++j;
  const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  return slope * (x - xp[j]) + fp[j];
}

struct LossParams {
  double attn_ohc[kBands], bw[kBands], lowknee[kBands], cr[kBands], attn_ihc[kBands];
};

inline void loss_parameters(const double* HL, const double* cf, LossParams& p) {  // pyhaspi2.py:779-807
  const double fv[8] = {cf[0], 250.0, 500.0, 1000.0, 2000.0, 4000.0, 6000.0, cf[kBands - 1]};
  const double lv[8] = {HL[0], HL[0], HL[1], HL[2], HL[3], HL[4], HL[5], HL[5]};
  for (int i = 0; i < kBands; ++i) {
    double loss = interp1(cf[i], fv, lv, 8);
    if (loss < 0) loss = 0.0;
    const double cr = 1.25 + 2.25 * i / (kBands - 1);
    const double max_ohc = 70.0 * (1.0 - 1.0 / cr), thr_ohc = 1.25 * max_ohc;
    if (loss < thr_ohc) {
      p.attn_ohc[i] = 0.8 * loss;
      p.attn_ihc[i] = 0.2 * loss;
    } else {
      p.attn_ohc[i] = 0.8 * thr_ohc;
      p.attn_ihc[i] = 0.2 * thr_ohc + (loss - thr_ohc);
    }
    const double r = p.attn_ohc[i] / 50.0;
    p.bw[i] = 1.0 + r + 2.0 * pow(r, 6.0);
    p.lowknee[i] = p.attn_ohc[i] + 30.0;
    const double upamp = 30.0 + 70.0 / cr;
    p.cr[i] = (100.0 - p.lowknee[i]) / (upamp + p.attn_ohc[i] - p.lowknee[i]);
  }
}

// eb_EarModel's pre-loop tables, itype = 0 (pyhaspi2.py:1157-1171)
inline void make_band_consts(const double* hl /*6 or null*/, BandConst* out) {
  double cf[kBands];
  center_freqs(cf);
  const double zero[6] = {0, 0, 0, 0, 0, 0}, full[6] = {100, 100, 100, 100, 100, 100};
  LossParams px, py, pm;
  loss_parameters(zero, cf, px);
  loss_parameters(hl ? hl : zero, cf, py);
  loss_parameters(full, cf, pm);
  for (int i = 0; i < kBands; ++i) {
    BandConst& b = out[i];
    b.cf = cf[i];
    b.erb = 24.7 + cf[i] / 9.26449;
    b.bw1 = pm.bw[i];
    const LossParams* pp[2] = {&px, &py};
    for (int q = 0; q < 2; ++q) {
      b.attn_ohc[q] = pp[q]->attn_ohc[i];
      b.bwmin[q] = pp[q]->bw[i];
      b.lowknee[q] = pp[q]->lowknee[i];
      b.cr[q] = pp[q]->cr[i];
      b.attn_ihc[q] = pp[q]->attn_ihc[i];
    }
  }
}

// ------------------------------------------------------------- window helpers
inline double bessel_i0(double x) {  // power series, converges fast for the betas used here (< 20)
  const double q = x * x / 4.0;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 500; ++k) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}
inline void kaiser(int m, double beta, std::vector<double>& w) {  // np.kaiser(m, beta)
  w.resize(m);
  const double alpha = (m - 1) / 2.0, d = bessel_i0(beta);
  for (int n = 0; n < m; ++n) {
    const double r = (n - alpha) / alpha;
    w[n] = bessel_i0(beta * sqrt(fmax(0.0, 1.0 - r * r))) / d;
  }
}
inline double sinc(double x) { return x == 0.0 ? 1.0 : sin(kPi * x) / (kPi * x); }  // np.sinc

// -------------------------------------------- resampy kaiser_best (HASPI 24 kHz)
// librosa.resample -> resampy.resample(filter='kaiser_best') (pyhaspi2.py:815;
// see oracle/resampy_kaiser.py for the published algorithm).  For up-sampling
// by the reduced ratio up/down the interpolation weights depend only on the
// phase r = (t * down) mod up, so they are tabulated once: taps[r][0..63] are
// the left-wing weights of x[n - i], taps[r][64..127] the right-wing weights
// of x[n + 1 + k].
struct ResampyTaps {
  int up = 1, down = 1;
  std::vector<double> taps;  // [up][128]
};

inline void make_resampy_taps(int fs_in, int fs_out, ResampyTaps& rt) {
  const int g = std::gcd(fs_in, fs_out);
  rt.up = fs_out / g;
  rt.down = fs_in / g;
  const int num_zeros = 64, per_zero = 512, n = per_zero * num_zeros;
  const double rolloff = 0.9475937167399596, beta = 14.769656459379492;
  std::vector<double> kw, win(n + 1), dwin(n + 1, 0.0);
  kaiser(2 * n + 1, beta, kw);
  for (int i = 0; i <= n; ++i) {
    const double t = (double)num_zeros * i / n;  // np.linspace(0, num_zeros, n + 1)
    win[i] = kw[n + i] * rolloff * sinc(rolloff * t);
  }
  for (int i = 0; i < n; ++i) dwin[i] = win[i + 1] - win[i];
  rt.taps.assign((size_t)rt.up * 128, 0.0);
  const int nwin = n + 1;
  for (int r = 0; r < rt.up; ++r) {
    double frac = (double)r / rt.up;  // scale = 1 when up-sampling
    double f = frac * per_zero;
    int off = (int)f;
    double eta = f - off;
    int cnt = (nwin - off) / per_zero;
    for (int i = 0; i < cnt && i < 64; ++i)
      rt.taps[(size_t)r * 128 + i] = win[off + i * per_zero] + eta * dwin[off + i * per_zero];
    frac = 1.0 - frac;
    f = frac * per_zero;
    off = (int)f;
    eta = f - off;
    cnt = (nwin - off) / per_zero;
    for (int k = 0; k < cnt && k < 64; ++k)
      rt.taps[(size_t)r * 128 + 64 + k] = win[off + k * per_zero] + eta * dwin[off + k * per_zero];
  }
}

// ------------------------------------------------ HASPI modulation filterbank
// ebm_ModFilt (pyhaspi2.py:275-339): demodulate by sqrt(2) exp(-j w n), Hann
// low-pass, re-modulate and take the real part.  Algebraically that is one
// real, zero-phase FIR band-pass per band,
//   g_m[k] = 2 h_m[k] cos(w_m (k - nh_m)),  w_m = pi cf_m / 1280   (m >= 1)
//   g_0[k] = h_0[k]
// applied as out[i] = sum_k g_m[k] x[i + nh_m - k] with zeros outside [0, n).
struct ModFilters {
  int nhalf[kNumMod];
  int offset[kNumMod + 1];  // start of band m's taps in `taps`
  std::vector<float> taps;
};

inline void make_mod_filters(ModFilters& mf, double fsub = 2560.0) {
  const double cf[kNumMod] = {2, 6, 10, 16, 25, 40, 64, 100, 160, 256};
  const double fnyq = 0.5 * fsub;
  mf.taps.clear();
  mf.offset[0] = 0;
  for (int m = 0; m < kNumMod; ++m) {
    const double t = (m < 2) ? 0.24 : 0.24 * cf[2] / cf[m];
    const int nfir = 2 * (int)floor(t * fsub / 2.0);
    const int nh = nfir / 2;
    mf.nhalf[m] = nh;
    std::vector<double> h(nfir + 1);
    double s = 0.0;
    for (int k = 0; k <= nfir; ++k) {
      h[k] = 0.5 - 0.5 * cos(2.0 * kPi * k / nfir);  // np.hanning(nfir + 1)
      s += h[k];
    }
    const double w = kPi * cf[m] / fnyq;
    for (int k = 0; k <= nfir; ++k) {
      const double g = (m == 0) ? h[k] / s : 2.0 * (h[k] / s) * cos(w * (k - nh));
      mf.taps.push_back((float)g);
    }
    mf.offset[m + 1] = (int)mf.taps.size();
  }
}

// Recursive (sliding-DFT) form of the same ten filters.  The Hann-windowed cosine
//   g_m(j) = (2 / S) (1/2 + 1/2 cos(pi j / nh)) cos(w j),  |j| <= nh      (m >= 1; band 0: no 2, w = 0)
// is a sum of three complex exponentials, so
//   out[i] = Re sum_k wgt_k C_k(i),   C_k(i) = sum_{|j| <= nh} e^{i th_k j} x[i - j],
//   th = {w, w + pi/nh, w - pi/nh},  wgt = {1, 1/2, 1/2} / S  ({1/2, 1/4, 1/4} / S for band 0),
// and each C_k slides in O(1):  C(i+1) = e^{i th} C(i) + e^{-i th nh} x[i+1+nh] - e^{i th (nh+1)} x[i-nh].
// ~27 FP64 operations per output and band instead of up to 615 FP32 taps.
struct ModRecBand {
  double rot[3][2], cin[3][2], cout[3][2], wgt[3];
  int nh, pad;
};
inline void make_mod_recursions(ModRecBand* mr /*[10]*/, double fsub = 2560.0) {
  const double cf[kNumMod] = {2, 6, 10, 16, 25, 40, 64, 100, 160, 256};
  const double fnyq = 0.5 * fsub;
  for (int m = 0; m < kNumMod; ++m) {
    const double t = (m < 2) ? 0.24 : 0.24 * cf[2] / cf[m];
    const int nfir = 2 * (int)floor(t * fsub / 2.0), nh = nfir / 2;
    double s = 0.0;
    for (int k = 0; k <= nfir; ++k) s += 0.5 - 0.5 * cos(2.0 * kPi * k / nfir);
    const double w = (m == 0) ? 0.0 : kPi * cf[m] / fnyq, delta = kPi / nh;
    const double th[3] = {w, w + delta, w - delta};
    const double wg[3] = {(m == 0) ? 0.5 : 1.0, (m == 0) ? 0.25 : 0.5, (m == 0) ? 0.25 : 0.5};
    mr[m].nh = nh;
    mr[m].pad = 0;
    for (int k = 0; k < 3; ++k) {
      mr[m].rot[k][0] = cos(th[k]);
      mr[m].rot[k][1] = sin(th[k]);
      mr[m].cin[k][0] = cos(th[k] * nh);
      mr[m].cin[k][1] = -sin(th[k] * nh);
      mr[m].cout[k][0] = cos(th[k] * (nh + 1));
      mr[m].cout[k][1] = sin(th[k] * (nh + 1));
      mr[m].wgt[k] = wg[k] / s;
    }
  }
}

// cosine basis of ebm_CepCoef (pyhaspi2.py:343-349), coefficients 1..5 only
inline void make_cep_basis(float* cepm /*[32][5]*/) {
  for (int nb = 1; nb <= kNumCep; ++nb) {
    double b[kBands], nrm = 0.0;
    for (int k = 0; k < kBands; ++k) {
      b[k] = cos(nb * kPi * k / (kBands - 1));
      nrm += b[k] * b[k];
    }
    nrm = sqrt(nrm);
    for (int k = 0; k < kBands; ++k) cepm[k * kNumCep + (nb - 1)] = (float)(b[k] / nrm);
  }
}

// ------------------------------------------------------------ ESTOI resampler
// pystoi.utils.resample_oct: Octave-compatible Kaiser-windowed sinc, applied by
// scipy.signal.resample_poly(x, up, down, window=h / sum(h)).
struct PolyFilter {
  int up = 1, down = 1, half = 0;
  std::vector<double> h;  // up * h / sum(h), length 2 half + 1
};

inline void make_estoi_resampler(int fs_in, int fs_out, PolyFilter& pf) {
  const int g = std::gcd(fs_in, fs_out);
  const int p = fs_out / g, q = fs_in / g;
  pf.up = p;
  pf.down = q;
  const double fc = 1.0 / (2.0 * (p > q ? p : q)), roll = fc / 10.0, rej = 60.0;
  const int L = (int)ceil((rej - 8.0) / (28.714 * roll));
  const double beta = 0.1102 * (rej - 8.7);
  std::vector<double> kw;
  kaiser(2 * L + 1, beta, kw);
  pf.half = L;
  pf.h.resize(2 * L + 1);
  double s = 0.0;
  for (int i = 0; i <= 2 * L; ++i) {
    pf.h[i] = kw[i] * 2.0 * p * fc * sinc(2.0 * fc * (i - L));
    s += pf.h[i];
  }
  for (auto& v : pf.h) v = v / s * p;
}

// Polyphase form of resample_poly's upfirdn: y[m] = sum_{k=-K..K} taps[r][k+K] x[n-k],
// n = (m * down) / up, r = (m * down) % up, taps[r][k+K] = h[half + r + up k] (0 outside).
struct PolyTaps {
  int up = 1, down = 1, K = 0;
  std::vector<double> taps;  // [up][2K+1]
};
inline void make_estoi_polytaps(int fs_in, int fs_out, PolyTaps& pt) {
  PolyFilter pf;
  make_estoi_resampler(fs_in, fs_out, pf);
  pt.up = pf.up;
  pt.down = pf.down;
  pt.K = pf.half / pf.up + 1;
  const int nt = 2 * pt.K + 1;
  pt.taps.assign((size_t)pt.up * nt, 0.0);
  for (int r = 0; r < pt.up; ++r)
    for (int k = -pt.K; k <= pt.K; ++k) {
      const int i = pf.half + r + pt.up * k;
      if (i >= 0 && i <= 2 * pf.half) pt.taps[(size_t)r * nt + (k + pt.K)] = pf.h[i];
    }
}

// one-third-octave band edges as rfft bin ranges [lo, hi) (pystoi.utils.thirdoct)
inline void make_thirdoct_bins(int* lo, int* hi, int fs = 10000, int nfft = 512, int nb = 15, double fmin = 150.0) {
  const int nf = nfft / 2 + 1;
  std::vector<double> f(nf);
  for (int i = 0; i < nf; ++i) f[i] = (double)fs * i / nfft;
  for (int k = 0; k < nb; ++k) {
    const double fl = fmin * pow(2.0, (2.0 * k - 1) / 6.0), fh = fmin * pow(2.0, (2.0 * k + 1) / 6.0);
    int a = 0, b = 0;
    double da = 1e300, db = 1e300;
    for (int i = 0; i < nf; ++i) {
      const double ea = (f[i] - fl) * (f[i] - fl), eb = (f[i] - fh) * (f[i] - fh);
      if (ea < da) { da = ea; a = i; }
      if (eb < db) { db = eb; b = i; }
    }
    lo[k] = a;
    hi[k] = b;
  }
}

// ----------------------------------------------------------- SIIB gammatone
// squared gammatone magnitude responses [28][201] (pysiib gammatone(), 4th order,
// Holdsworth bandwidth normalisation, peak-normalised)
constexpr int kSiibBands = 28;
constexpr int kSiibBins = 201;
inline void make_siib_g2(float* g2 /*[28][201]*/) {
  const double mn = 100.0, mx = 6500.0;
  const double e0 = 21.4 * log10(4.37 * (mn / 1000.0) + 1.0), e1 = 21.4 * log10(4.37 * (mx / 1000.0) + 1.0);
  const double a = 36.0 / (kPi * 720.0 * pow(2.0, -6.0));  // (3!)^2 / (pi 6! 2^-6)
  for (int j = 0; j < kSiibBands; ++j) {
    const double erb = e0 + (e1 - e0) * j / (kSiibBands - 1);
    const double cf = (pow(10.0, erb / 21.4) - 1.0) / 4.37 * 1000.0;
    const double b = a * 24.7 * (4.37 * cf / 1000.0 + 1.0);
    double row[kSiibBins], mxv = 0.0;
    for (int i = 0; i < kSiibBins; ++i) {
      const double f = 16000.0 * i / 400.0;
      row[i] = 1.0 / pow(b * b + (f - cf) * (f - cf), 2.0);
      if (row[i] > mxv) mxv = row[i];
    }
    for (int i = 0; i < kSiibBins; ++i) {
      const double v = row[i] / mxv;
      g2[j * kSiibBins + i] = (float)(v * v);
    }
  }
}

}  // namespace host
}  // namespace nele
