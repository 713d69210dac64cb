// Counter-based normal generator shared by the HASPI kernels: the cepstral dither of
// ebm_CepCoef (pyhaspi2.py:362-365) and the basilar-membrane noise of eb_BMaddnoise
// (pyhaspi2.py:1091-1095).  The reference draws both from numpy's global stream; here every
// draw is a pure function of (seed, stream, row, band), so a batch is reproducible whatever
// the launch geometry.
#pragma once
#include <stdint.h>

namespace nele {

__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t* hi) {
  const uint64_t p = (uint64_t)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
// Philox4x32-10 counter-based generator (Salmon et al. 2011)
__device__ __forceinline__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, h1;
    const uint32_t l0 = mulhilo(0xD2511F53u, c[0], &h0), l1 = mulhilo(0xCD9E8D57u, c[2], &h1);
    const uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// unit normal for (seed, stream = pair*2+q, row, band)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t stream, uint32_t row, uint32_t band) {
  uint32_t c[4] = {row, band >> 1, (uint32_t)stream, (uint32_t)(stream >> 32)};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  return (band & 1) ? r * sn : r * cs;
}


// two independent unit normals for (seed, stream, row, band) from one Philox block
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint64_t stream, uint32_t row, uint32_t band,
                                               float& n0, float& n1) {
  uint32_t c[4] = {row, band, (uint32_t)stream, (uint32_t)(stream >> 32) ^ 0x5bd1e995u};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = ((float)(c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u4 = ((float)(c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  n0 = sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
  n1 = sqrtf(-2.0f * __logf(u3)) * __cosf(6.283185307179586f * u4);
}

// four independent unit normals (two Box-Muller pairs, sine and cosine branch of each)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t stream, uint32_t row, uint32_t band, float (&z)[4]) {
  uint32_t c[4] = {row, band, (uint32_t)stream, (uint32_t)(stream >> 32) ^ 0x2545f491u};
  philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = ((float)(c[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u4 = ((float)(c[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r0 = sqrtf(-2.0f * __logf(u1)), r1 = sqrtf(-2.0f * __logf(u3));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u2, &s0, &c0);
  __sincosf(6.283185307179586f * u4, &s1, &c1);
  z[0] = r0 * c0;
  z[1] = r0 * s0;
  z[2] = r1 * c1;
  z[3] = r1 * s1;
}

}  // namespace nele
