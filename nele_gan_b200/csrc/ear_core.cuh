// Per-band arithmetic of the HASPI ear model (one lane = one auditory band of
// one signal).  Shared by the CUDA kernels in haspi.cu and by the host-side
// emulation harness tests/host_emul (g++), so the band math can be checked on
// a machine without a GPU.  Reference: pyHASPI/pyhaspi2.py (line numbers in
// the comments).
//
// T is the type of the linear recurrences (gammatone poles, carrier rotation,
// compression low-pass, IHC adaptation); the pointwise log/pow/sqrt section is
// always single precision (|error| ~1e-5 dB, three orders below the 0.1 dB
// dither the model itself adds).
#pragma once
#include <math.h>

#include "common.cuh"

namespace nele {

// Per-band constants of one scoring call (depend on the audiogram HL only).
// Index q: 0 = reference signal x (always normal hearing, pyhaspi2.py:1162-1167),
//          1 = processed signal y.
struct BandConst {
  double cf;        // centre frequency                      (:753-777)
  double erb;       // 24.7 + cf / 9.26449                   (:866)
  double bw1;       // control-path bandwidth, HL = 100 dB   (:1168-1171)
  double attn_ohc[2], bwmin[2], lowknee[2], cr[2], attn_ihc[2];  // (:779-807)
};

// Linearised IHC adaptation update (pyhaspi2.py:1040-1071):
//   V1' = m11 V1 + m12 V2 + g1 V0,  V2' = m21 V1 + m22 V2 + g2 V0,
//   out = max((V0 - V1') * r1inv, 0).
struct IhcConst {
  double m11, m12, m21, m22, g1, g2, r1inv;
};

inline IhcConst make_ihc_const(double delta = 2.0, double fs = 24000.0) {
  if (delta < 1.0001) delta = 1.0001;
  const double tau1 = 0.002, tau2 = 0.060, T = 1.0 / fs;
  const double R1 = 1.0 / delta, R2 = 0.5 * (1.0 - R1), R3 = R2;
  const double C1 = tau1 * (R1 + R2) / (R1 * R2), C2 = tau2 / ((R1 + R2) * R3);
  const double a11 = R1 + R2 + R1 * R2 * (C1 / T), a12 = -R1, a21 = -R3,
               a22 = R2 + R3 + R2 * R3 * (C2 / T);
  const double denom = 1.0 / (a11 * a22 - a21 * a12);
  const double R12C1 = R1 * R2 * (C1 / T), R23C2 = R2 * R3 * (C2 / T);
  IhcConst k;
  k.m11 = denom * a22 * R12C1;
  k.m12 = -denom * a12 * R23C2;
  k.g1 = denom * a22 * R2;
  k.m21 = -denom * a21 * R12C1;
  k.m22 = denom * a11 * R23C2;
  k.g2 = -denom * a21 * R2;
  k.r1inv = 1.0 / R1;
  return k;
}

// 4th-order baseband gammatone (pyhaspi2.py:870-878):
//   H(z) = gain (1 + 2a z^-1)^2 / (1 - a z^-1)^4,  a = exp(-2 pi/fs 1.019 BW ERB)
// run as four cascaded one-pole sections followed by the numerator.
template <typename T>
struct GtCoef {
  T a, c1, c2;  // pole, 4a, 4a^2
  float gain;
};

template <typename T>
NELE_HD GtCoef<T> make_gt(double bw, double erb) {
  const double tpt = 2.0 * 3.14159265358979323846 / 24000.0;
  const double a = exp(-(bw * tpt * erb * 1.019));
  const double a1 = 4.0 * a, a2 = -6.0 * a * a, a3 = 4.0 * a * a * a, a4 = -a * a * a * a, a5 = 4.0 * a * a;
  GtCoef<T> g;
  g.a = (T)a;
  g.c1 = (T)a1;
  g.c2 = (T)a5;
  g.gain = (float)(2.0 * (1.0 - a1 - a2 - a3 - a4) / (1.0 + a1 + a5));
  return g;
}

// DC group delay of the filter above, rounded as np.round does
// (scipy.signal.group_delay(..., w=1) evaluates at omega = 0; pyhaspi2.py:1117-1118).
NELE_HD double gt_group_delay(double bw, double erb) {
  const double tpt = 2.0 * 3.14159265358979323846 / 24000.0;
  const double a = exp(-(bw * tpt * erb * 1.019));
  return rint(4.0 * a / (1.0 - a) + 4.0 * a / (1.0 + 2.0 * a));
}

template <typename T>
struct Gt4 {
  T r1, r2, r3, r4, rp, i1, i2, i3, i4, ip;
  NELE_HD void reset() { r1 = r2 = r3 = r4 = rp = i1 = i2 = i3 = i4 = ip = (T)0; }
  // -> squared magnitude of the (unnormalised) complex output
  NELE_HD T step(const GtCoef<T>& k, T xr, T xi) {
    r1 = k.a * r1 + xr;
    i1 = k.a * i1 + xi;
    r2 = k.a * r2 + r1;
    i2 = k.a * i2 + i1;
    r3 = k.a * r3 + r2;
    i3 = k.a * i3 + i2;
    const T nr = k.a * r4 + r3;
    const T ni = k.a * i4 + i3;
    const T ur = nr + k.c1 * r4 + k.c2 * rp;
    const T ui = ni + k.c1 * i4 + k.c2 * ip;
    rp = r4;
    ip = i4;
    r4 = nr;
    i4 = ni;
    return ur * ur + ui * ui;
  }
  // same, also handing back the complex output (the basilar-membrane path needs its phase)
  NELE_HD T step_ri(const GtCoef<T>& k, T xr, T xi, T& ur, T& ui) {
    r1 = k.a * r1 + xr;
    i1 = k.a * i1 + xi;
    r2 = k.a * r2 + r1;
    i2 = k.a * i2 + i1;
    r3 = k.a * r3 + r2;
    i3 = k.a * i3 + i2;
    const T nr = k.a * r4 + r3;
    const T ni = k.a * i4 + i3;
    ur = nr + k.c1 * r4 + k.c2 * rp;
    ui = ni + k.c1 * i4 + k.c2 * ip;
    rp = r4;
    ip = i4;
    r4 = nr;
    i4 = ni;
    return ur * ur + ui * ui;
  }
};

// carrier cos(w t), -sin(w t) (pyhaspi2.py:843-861 rotates by -w); advanced by
// the same rotation recurrence and re-seeded from the exact phase by the
// caller every few hundred samples so single precision cannot drift.
template <typename T>
struct Carrier {
  T c, s, cn, sn;
  double w;
  NELE_HD void init(double cf) {
    w = 2.0 * 3.14159265358979323846 / 24000.0 * cf;
    cn = (T)cos(w);
    sn = (T)sin(w);
  }
  // state such that the next advance() yields sample t
  NELE_HD void seed_before(long long t) {
    const double ph = w * (double)(t - 1);
    c = (T)cos(ph);
    s = (T)(-sin(ph));
  }
  NELE_HD void advance() {
    const T nc = c * cn + s * sn;
    s = s * cn - c * sn;
    c = nc;
  }
};

#define NELE_LOG2_10 3.3219280948873623f
#define NELE_20_OVER_LOG2_10 6.0205999132796239f /* 20 log10(2) */
#define NELE_10_OVER_LOG2_10 3.0102999566398120f /* 10 log10(2) */

// single-instruction log2 / exp2 (MUFU.LG2 / MUFU.EX2, denormals flushed): __log2f / exp2f wrap
// the same instructions in a denormal rescue (compare, scale, fix-up: 4 and 3 extra instructions
// per call) that the ear model does not need -- inputs below 1.2e-38 are exact silence and clamp
// to the same floor as log2(0) = -inf.
NELE_HD float fast_lg2(float v) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
#else
  return log2f(v);
#endif
}
NELE_HD float fast_ex2(float v) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
#else
  return exp2f(v);
#endif
}
NELE_HD float db20(float v) { return NELE_20_OVER_LOG2_10 * fast_lg2(v); }          // 20 log10(v)
NELE_HD float undb20(float d) { return fast_ex2(d * (NELE_LOG2_10 / 20.0f)); }      // 10^(d/20)

// Control-path lane: envelope power accumulation for eb_BWadjust
// (pyhaspi2.py:1202-1205, 917-980).
template <typename T>
struct ControlLane {
  Carrier<T> car;
  GtCoef<T> k;
  Gt4<T> f;
  NELE_HD void init(const BandConst& b) {
    car.init(b.cf);
    k = make_gt<T>(b.bw1, b.erb);
    f.reset();
  }
  NELE_HD T step(T x) {
    car.advance();
    return f.step(k, x * car.c, x * car.s);
  }
};

// cdB -> bandwidth (pyhaspi2.py:971-980); sumsq = sum of unnormalised |u|^2.
NELE_HD double bw_from_control(double sumsq, double gain, int n, double bwmin, double bwmax) {
  const double crms = gain * sqrt(sumsq / (double)n);
  const double cdb = 20.0 * log10(crms) + 65.0;
  if (cdb < 50.0) return bwmin;
  if (cdb > 100.0) return bwmax;
  return bwmin + ((cdb - 50.0) / 50.0) * (bwmax - bwmin);
}

// Main-pass lane: control + signal gammatone, OHC compression, dB SL,
// IHC adaptation, then the 52-tap Hann FIR of ebm_EnvFilt evaluated only at
// the kept (every 9th) positions.  The lane runs on its own time axis delayed
// by `shift` samples so that group-delay compensation (pyhaspi2.py:1124-1129)
// costs nothing: at loop index i it consumes input sample i - shift.  The
// carrier is owned by the caller (the clean and the processed signal of a pair
// share it: same band, same shift).
template <typename T>
struct EarLane {
  GtCoef<T> kc, ks;
  Gt4<T> fc, fs;
  T zlp, v1, v2;
  T m11, m12, m21, m22, g1, g2;
  float r1inv, thr_low;
  float crfac_l2;  // (1 - 1/CR) log2(10)/20: compression slope in log2 units per dB
  float ohc_l2;    // -attnOHC log2(10)/20
  float ctl_db;    // 65 + 20 log10(control gain): level of the control envelope = ctl_db + 10 log10(|u|^2)
  float sig_db;    // 65 - attnIHC + 20 log10(signal gain)
  float acc[6];

  NELE_HD void init(const BandConst& b, int q, double bw_sig, const IhcConst& ih) {
    kc = make_gt<T>(b.bw1, b.erb);
    ks = make_gt<T>(bw_sig, b.erb);
    fc.reset();
    fs.reset();
    zlp = v1 = v2 = (T)0;
    m11 = (T)ih.m11; m12 = (T)ih.m12; m21 = (T)ih.m21; m22 = (T)ih.m22;
    g1 = (T)ih.g1; g2 = (T)ih.g2;
    r1inv = (float)ih.r1inv;
    thr_low = (float)b.lowknee[q];
    crfac_l2 = (float)((1.0 - 1.0 / b.cr[q]) * 3.3219280948873623 / 20.0);
    ohc_l2 = (float)(-b.attn_ohc[q] * 3.3219280948873623 / 20.0);
    ctl_db = (float)(65.0 + 20.0 * log10((double)kc.gain));
    sig_db = (float)(65.0 - b.attn_ihc[q] + 20.0 * log10((double)ks.gain));
    for (int d = 0; d < 6; ++d) acc[d] = 0.f;
  }

  // one input sample, already demodulated by the carrier (xr, xi) -> IHC-adapted
  // envelope in dB SL (pyhaspi2.py:1207-1229).  Levels are taken from the squared
  // magnitudes (10 log10 |u|^2 = 20 log10 |u|), which saves the two square roots;
  // the reference's 1e-30 floors only matter for exact silence, where both forms
  // clamp to the same value (thrLow for the control level, 0 dB SL for the envelope).
  NELE_HD float sample(T xr, T xi) {
    const float pc = (float)fc.step(kc, xr, xi);
    const float ps = (float)fs.step(ks, xr, xi);
    // eb_EnvCompressBM (:982-999)
    float le = fmaf(NELE_10_OVER_LOG2_10, fast_lg2(pc), ctl_db);
    le = fminf(fmaxf(le, thr_low), 100.0f);
    const float g = fast_ex2(fmaf(thr_low - le, crfac_l2, ohc_l2));  // 10^((-attnOHC - (le - thrLow)(1 - 1/CR)) / 20)
    const T b0 = (T)0.095107983402496;
    const T glp = b0 * (T)g + zlp;
    zlp = b0 * (T)g + (T)0.809784033195007 * glp;
    // eb_EnvSL2 (:1080-1083) on env = glp * gain * sqrt(ps)
    const float gl = (float)glp;
    const float v0 = fmaxf(fmaf(NELE_10_OVER_LOG2_10, fast_lg2(gl * gl * ps), sig_db), 0.0f);
    // eb_IHCadapt (:1065-1073)
    const T V0 = (T)v0;
    const T n1 = m11 * v1 + m12 * v2 + g1 * V0;
    const T n2 = m21 * v1 + m22 * v2 + g2 * V0;
    v1 = n1;
    v2 = n2;
    return fmaxf((float)(V0 - n1) * r1inv, 0.0f);
  }

  // HASPI version 1 needs the basilar-membrane motion as well (pyhaspi2.py:899, 996-998,
  // 1086-1087, 1075-1077): bm = gain (ur cos + ui sin) scaled by the three gains the envelope
  // went through.  The dB-SL and adaptation gains telescope, (y + e)/(env + e) * (out + e)/(y + e)
  // = (out + e)/(env + e) with e = 1e-30, so bm = glp gain (ur c + ui s) (out + e) / (env + e),
  // env = glp gain |u| the compressed envelope.  (cs, sn) is the carrier the caller demodulated
  // with, i.e. the reference's (coscf, sincf).
  NELE_HD float sample_bm(T xr, T xi, T cs, T sn, float& bm, float& ps_out) {
    const float pc = (float)fc.step(kc, xr, xi);
    T ur, ui;
    const float ps = (float)fs.step_ri(ks, xr, xi, ur, ui);
    ps_out = ps;  // squared magnitude of the (unnormalised) signal-path output, for HASQI's average levels
    float le = fmaf(NELE_10_OVER_LOG2_10, fast_lg2(pc), ctl_db);
    le = fminf(fmaxf(le, thr_low), 100.0f);
    const float g = fast_ex2(fmaf(thr_low - le, crfac_l2, ohc_l2));
    const T b0 = (T)0.095107983402496;
    const T glp = b0 * (T)g + zlp;
    zlp = b0 * (T)g + (T)0.809784033195007 * glp;
    const float gl = (float)glp;
    const float v0 = fmaxf(fmaf(NELE_10_OVER_LOG2_10, fast_lg2(gl * gl * ps), sig_db), 0.0f);
    const T V0 = (T)v0;
    const T n1 = m11 * v1 + m12 * v2 + g1 * V0;
    const T n2 = m21 * v1 + m22 * v2 + g2 * V0;
    v1 = n1;
    v2 = n2;
    const float out = fmaxf((float)(V0 - n1) * r1inv, 0.0f);
    const float sg = gl * ks.gain;
    const float env = sg * sqrtf(ps);
    const float fine = (float)(ur * cs + ui * sn);
    bm = (sg * fine) * ((out + 1.0e-30f) / (env + 1.0e-30f));
    return out;
  }

  // FIR bookkeeping.  Sample at loop index i = 9 b + P contributes
  // fir[d][P] * v to output j = b - 2 + d, d = 0..5  (fir[d][P] = h[9(d-2)+26-P]).
  template <int P>
  NELE_HD void accumulate(float v, const float* fir) {
#pragma unroll
    for (int d = 0; d < 6; ++d) acc[d] = fmaf(fir[d * 9 + P], v, acc[d]);
  }
  // after phase 8 of block b: returns output j = b - 2 and rotates
  NELE_HD float emit() {
    const float o = acc[0];
#pragma unroll
    for (int d = 0; d < 5; ++d) acc[d] = acc[d + 1];
    acc[5] = 0.f;
    return o;
  }
};

// Two FP32 values in one 64-bit register for Blackwell's two-wide FP32 instructions
// (fma / mul / add .rn.f32x2); component-wise float arithmetic on the host (tests/host_emul).
struct F2 {
#if defined(__CUDA_ARCH__)
  unsigned long long v;
#else
  float lo, hi;
#endif
};
NELE_HD F2 f2_pack(float a, float b) {
  F2 r;
#if defined(__CUDA_ARCH__)
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
#else
  r.lo = a;
  r.hi = b;
#endif
  return r;
}
NELE_HD void f2_unpack(F2 p, float& a, float& b) {
#if defined(__CUDA_ARCH__)
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
#else
  a = p.lo;
  b = p.hi;
#endif
}
NELE_HD F2 f2_fma(F2 a, F2 b, F2 c) {
  F2 d;
#if defined(__CUDA_ARCH__)
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
#else
  d.lo = fmaf(a.lo, b.lo, c.lo);
  d.hi = fmaf(a.hi, b.hi, c.hi);
#endif
  return d;
}
NELE_HD F2 f2_mul(F2 a, F2 b) {
  F2 d;
#if defined(__CUDA_ARCH__)
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
#else
  d.lo = a.lo * b.lo;
  d.hi = a.hi * b.hi;
#endif
  return d;
}
NELE_HD F2 f2_add(F2 a, F2 b) {
  F2 d;
#if defined(__CUDA_ARCH__)
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
#else
  d.lo = a.lo + b.lo;
  d.hi = a.hi + b.hi;
#endif
  return d;
}

// The main-pass lane of both signals of a pair, two-wide (experimental: haspi_ear_x2_kernel, NELE_F32X2=1).
// Same arithmetic per component as EarLane<float>::sample; tests/host_emul/ear_x2_emul.cpp checks the two against
// each other on the host.
struct EarLane2 {
  float kca, kcc1, kcc2, ctl_db;   // control filter: the same for both signals
  F2 ksa, ksc1, ksc2, sig_db;      // signal filter: (x, y)
  F2 cr1, cr2, cr3, cr4, crp, ci1, ci2, ci3, ci4, cip;
  F2 sr1, sr2, sr3, sr4, srp, si1, si2, si3, si4, sip;
  F2 zlp, v1, v2;
  float m11, m12, m21, m22, g1, g2, r1inv;
  float thr_x, thr_y;
  F2 thr, crfac, ohc;
  F2 acc[6];

  NELE_HD void init(const BandConst& b, double bwx, double bwy, const IhcConst& ih) {
    const GtCoef<float> kc = make_gt<float>(b.bw1, b.erb);
    const GtCoef<float> kx = make_gt<float>(bwx, b.erb), ky = make_gt<float>(bwy, b.erb);
    kca = kc.a;
    kcc1 = kc.c1;
    kcc2 = kc.c2;
    ksa = f2_pack(kx.a, ky.a);
    ksc1 = f2_pack(kx.c1, ky.c1);
    ksc2 = f2_pack(kx.c2, ky.c2);
    const F2 z = f2_pack(0.f, 0.f);
    cr1 = cr2 = cr3 = cr4 = crp = ci1 = ci2 = ci3 = ci4 = cip = z;
    sr1 = sr2 = sr3 = sr4 = srp = si1 = si2 = si3 = si4 = sip = z;
    zlp = v1 = v2 = z;
    m11 = (float)ih.m11; m12 = (float)ih.m12; m21 = (float)ih.m21; m22 = (float)ih.m22;
    g1 = (float)ih.g1; g2 = (float)ih.g2;
    r1inv = (float)ih.r1inv;
    thr_x = (float)b.lowknee[0];
    thr_y = (float)b.lowknee[1];
    thr = f2_pack(thr_x, thr_y);
    crfac = f2_pack((float)((1.0 - 1.0 / b.cr[0]) * 3.3219280948873623 / 20.0),
                    (float)((1.0 - 1.0 / b.cr[1]) * 3.3219280948873623 / 20.0));
    ohc = f2_pack((float)(-b.attn_ohc[0] * 3.3219280948873623 / 20.0), (float)(-b.attn_ohc[1] * 3.3219280948873623 / 20.0));
    ctl_db = (float)(65.0 + 20.0 * log10((double)kc.gain));
    sig_db = f2_pack((float)(65.0 - b.attn_ihc[0] + 20.0 * log10((double)kx.gain)),
                     (float)(65.0 - b.attn_ihc[1] + 20.0 * log10((double)ky.gain)));
#pragma unroll
    for (int d = 0; d < 6; ++d) acc[d] = z;
  }

  // one demodulated sample pair -> IHC-adapted envelopes (x, y) in dB SL; EarLane<float>::sample two-wide
  NELE_HD F2 sample(F2 xr, F2 xi) {
    const F2 KA = f2_pack(kca, kca), KC1 = f2_pack(kcc1, kcc1), KC2 = f2_pack(kcc2, kcc2);
    cr1 = f2_fma(KA, cr1, xr);
    ci1 = f2_fma(KA, ci1, xi);
    cr2 = f2_fma(KA, cr2, cr1);
    ci2 = f2_fma(KA, ci2, ci1);
    cr3 = f2_fma(KA, cr3, cr2);
    ci3 = f2_fma(KA, ci3, ci2);
    const F2 cnr = f2_fma(KA, cr4, cr3), cni = f2_fma(KA, ci4, ci3);
    const F2 cur = f2_fma(KC2, crp, f2_fma(KC1, cr4, cnr)), cui = f2_fma(KC2, cip, f2_fma(KC1, ci4, cni));
    crp = cr4;
    cip = ci4;
    cr4 = cnr;
    ci4 = cni;
    const F2 pc = f2_fma(cur, cur, f2_mul(cui, cui));
    sr1 = f2_fma(ksa, sr1, xr);
    si1 = f2_fma(ksa, si1, xi);
    sr2 = f2_fma(ksa, sr2, sr1);
    si2 = f2_fma(ksa, si2, si1);
    sr3 = f2_fma(ksa, sr3, sr2);
    si3 = f2_fma(ksa, si3, si2);
    const F2 snr = f2_fma(ksa, sr4, sr3), sni = f2_fma(ksa, si4, si3);
    const F2 sur = f2_fma(ksc2, srp, f2_fma(ksc1, sr4, snr)), sui = f2_fma(ksc2, sip, f2_fma(ksc1, si4, sni));
    srp = sr4;
    sip = si4;
    sr4 = snr;
    si4 = sni;
    const F2 ps = f2_fma(sur, sur, f2_mul(sui, sui));
    // eb_EnvCompressBM
    const F2 C10 = f2_pack(NELE_10_OVER_LOG2_10, NELE_10_OVER_LOG2_10), NEG1 = f2_pack(-1.f, -1.f);
    float pcx, pcy;
    f2_unpack(pc, pcx, pcy);
    float lx, ly;
    f2_unpack(f2_fma(C10, f2_pack(fast_lg2(pcx), fast_lg2(pcy)), f2_pack(ctl_db, ctl_db)), lx, ly);
    lx = fminf(fmaxf(lx, thr_x), 100.0f);
    ly = fminf(fmaxf(ly, thr_y), 100.0f);
    float ax, ay;
    f2_unpack(f2_fma(f2_fma(f2_pack(lx, ly), NEG1, thr), crfac, ohc), ax, ay);   // (thrLow - le) crfac + ohc
    const F2 G = f2_pack(fast_ex2(ax), fast_ex2(ay));
    const F2 B0 = f2_pack(0.095107983402496f, 0.095107983402496f), K8 = f2_pack(0.809784033195007f, 0.809784033195007f);
    const F2 glp = f2_fma(B0, G, zlp);
    zlp = f2_fma(K8, glp, f2_mul(B0, G));
    // eb_EnvSL2
    float ex, ey;
    f2_unpack(f2_mul(f2_mul(glp, glp), ps), ex, ey);
    float v0x, v0y;
    f2_unpack(f2_fma(C10, f2_pack(fast_lg2(ex), fast_lg2(ey)), sig_db), v0x, v0y);
    const F2 V0 = f2_pack(fmaxf(v0x, 0.0f), fmaxf(v0y, 0.0f));
    // eb_IHCadapt
    const F2 n1 = f2_fma(f2_pack(g1, g1), V0, f2_fma(f2_pack(m12, m12), v2, f2_mul(f2_pack(m11, m11), v1)));
    const F2 n2 = f2_fma(f2_pack(g2, g2), V0, f2_fma(f2_pack(m22, m22), v2, f2_mul(f2_pack(m21, m21), v1)));
    v1 = n1;
    v2 = n2;
    float ox, oy;
    f2_unpack(f2_mul(f2_fma(n1, NEG1, V0), f2_pack(r1inv, r1inv)), ox, oy);
    return f2_pack(fmaxf(ox, 0.0f), fmaxf(oy, 0.0f));
  }

  template <int P>
  NELE_HD void accumulate(F2 v, const float* fir) {
#pragma unroll
    for (int d = 0; d < 6; ++d) acc[d] = f2_fma(f2_pack(fir[d * 9 + P], fir[d * 9 + P]), v, acc[d]);
  }
  NELE_HD F2 emit() {
    const F2 o = acc[0];
#pragma unroll
    for (int d = 0; d < 5; ++d) acc[d] = acc[d + 1];
    acc[5] = f2_pack(0.f, 0.f);
    return o;
  }
};


// h[k] = np.hanning(52)[k] / sum  -> fir[d*9 + p] = h[9(d-2) + 26 - p] (0 outside 0..51)
inline void make_env_fir(float* fir /*[54]*/) {
  double h[kEnvTaps], s = 0.0;
  for (int k = 0; k < kEnvTaps; ++k) {
    h[k] = 0.5 - 0.5 * cos(2.0 * 3.14159265358979323846 * k / (kEnvTaps - 1));
    s += h[k];
  }
  for (int d = 0; d < 6; ++d)
    for (int p = 0; p < 9; ++p) {
      const int k = 9 * (d - 2) + kEnvHalf - p;
      fir[d * 9 + p] = (k >= 0 && k < kEnvTaps) ? (float)(h[k] / s) : 0.f;
    }
}

}  // namespace nele
