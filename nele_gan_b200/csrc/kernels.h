// Launch interface between the engine (engine.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ear_core.cuh"

namespace nele {

// Per-pair geometry of one sub-batch ("chunk"), device arrays of length n.
struct PairGeom {
  const int64_t* off16;  // start of the pair in the 16 kHz (input-rate) buffers
  const int32_t* len16;  // samples at the input rate
  const int64_t* off24;  // start in the 24 kHz buffers (per signal plane)
  const int32_t* n24;    // ceil(len * 24000 / fs)
  const int64_t* offsub; // start (in rows of 32) in the decimated-envelope buffers
  const int32_t* nsub;   // ceil(n24 / 9)
};

struct HaspiBuffers {
  const float* ref;      // device, input rate
  const float* deg;
  float* x24;            // [2][tot24]
  double* mid;           // [2][tot24]
  int64_t tot24;
  double* bw;            // [n][2][32]
  int32_t* shift;        // [n][32]
  float* envlp;          // [2][totsub][32]
  int64_t totsub;
  int32_t* rowsel;       // [totsub] compacted index of each envelope row, -1 = dropped
  int32_t* nsel;         // [n]
  float* cep;            // [2][5][totsub]  (pair p, signal q, coef j at ((q*5+j)*totsub + offsub[p]))
  double* cepmean;       // [n][2][5]
  double* modsum;        // [n][5][10][5]  {Sx, Sy, Sxx, Syy, Sxy}
  const BandConst* bands;  // [32] device
  const double* rs_taps;   // [up][128] device
  int rs_up, rs_down;      // 24 kHz resampler ratio (1/1 = identity)
  const float* dither;     // [2][dither_rows][32] device or null
  int64_t dither_rows;
  uint64_t seed;
  int no_dither;
  int64_t pair_base;       // global index of the chunk's first pair (dither keying)
};

void haspi_upload_constants(cudaStream_t s);
// returns number of kernel launches issued
int haspi_run(const PairGeom& g, const HaspiBuffers& b, int n, int max_nsub, bool f64, cudaStream_t s);
// scores: raw Intel [n] and aveCM [n][10]; status byte per pair
int haspi_finish(const HaspiBuffers& b, int n, double* intel, double* raw10, int32_t* status, cudaStream_t s);

}  // namespace nele
