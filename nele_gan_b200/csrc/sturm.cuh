// Sturm count of a symmetric tridiagonal matrix for the bisection kernels of siib_eig.cu (siib_trieig_kernel,
// siib_smallvec_kernel).  The routine is __host__ __device__ so that tests/host_emul/sturm_emul.cpp can run the very
// same arithmetic with g++ against a ratio-form count in long double (graded, clustered and Toeplitz matrices, the
// overflow / underflow guard included); on the device the bit operations map to single instructions.
#pragma once
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace nele {

#if defined(__CUDACC__)
typedef double2 sturm_pair;
#else
struct alignas(16) sturm_pair {
  double x, y;
};
#endif

NELE_HD int sturm_hi(double v) {  // high word: sign, exponent, top of the mantissa
#if defined(__CUDA_ARCH__)
  return __double2hiint(v);
#else
  uint64_t u;
  memcpy(&u, &v, 8);
  return (int)(uint32_t)(u >> 32);
#endif
}
NELE_HD double sturm_from_hi(int hi) {  // the double whose high word is hi and whose low word is 0
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, 0);
#else
  const uint64_t u = (uint64_t)(uint32_t)hi << 32;
  double v;
  memcpy(&v, &u, 8);
  return v;
#endif
}
NELE_HD unsigned sturm_push_sign(unsigned sg, double v) {  // (sg << 1) | sign(v)
#if defined(__CUDA_ARCH__)
  return __funnelshift_l((unsigned)__double2hiint(v), sg, 1);
#else
  return (sg << 1) | ((unsigned)sturm_hi(v) >> 31);
#endif
}
NELE_HD int sturm_popc(unsigned v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

// number of eigenvalues of T (scaled so that |entries| <= 1) below x: sign changes of the
// three-term recurrence p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}.
// de[i] = {d_i, e_{i-1}^2} (de[0].y unused), padded to LEN = sturm_len(n) entries with {1000, 0} (p keeps its sign
// there, |x| <= 3): one 128-bit shared-memory load per step, broadcast to the whole CTA, and blocks of 16 steps with a
// compile-time trip count.  The kernels are issue bound (ncu: 71 % issue slots, FP64 pipe 39 %; 15 instructions per
// step before this form, of which 3 are the FP64 arithmetic), so instructions are time:
//   * the sign of every p_i is shifted into a bit mask (one funnel shift per step) and the changes of a block counted
//     with one xor + popc, instead of an xor + shift + add per step;
//   * the range check that guards the recurrence against overflow looks at the exponent fields of the high words and
//     rescales by an exact power of two: |p| grows by at most 7 per step (2^45 per block), the window is 2^+-128.
constexpr int kSturmBlk = 16;
constexpr int sturm_len(int n) { return 1 + ((n - 1 + kSturmBlk - 1) / kSturmBlk) * kSturmBlk; }

template <int LEN>
NELE_HD int sturm_count(const sturm_pair* __restrict__ de, double x) {
  double pm = 1.0, p = de[0].x - x;
  unsigned sg = (unsigned)sturm_hi(p) >> 31;  // bit 0: sign of the newest p; p_{-1} = 1 is positive
  int cnt = (int)sg;
#pragma unroll 1
  for (int i0 = 1; i0 < LEN; i0 += kSturmBlk) {
#pragma unroll
    for (int i = 0; i < kSturmBlk; ++i) {
      const sturm_pair c = de[i0 + i];
      const double pn = fma(c.x - x, p, -c.y * pm);
      sg = sturm_push_sign(sg, pn);
      pm = p;
      p = pn;
    }
    cnt += sturm_popc((sg ^ (sg >> 1)) & 0xffffu);  // a sign change = one more eigenvalue below x
    const int ep = sturm_hi(p) & 0x7ff00000, em = sturm_hi(pm) & 0x7ff00000;
    const int e = ep > em ? ep : em;  // exponent field of max(|p|, |pm|)
    if ((unsigned)(e - ((1023 - 128) << 20)) > (unsigned)(256 << 20)) {
      const double sc = sturm_from_hi(0x7fe00000 - e);  // 2^-(exponent): exact, the signs stay
      p *= sc;
      pm *= sc;
    }
  }
  return cnt;
}

}  // namespace nele
