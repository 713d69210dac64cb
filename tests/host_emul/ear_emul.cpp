// Host emulation of the per-band device code in nele_gan_b200/csrc/ear_core.cuh.
// TEST TOOL ONLY: compiled by tests/test_host_emul.py with g++ to check the
// band arithmetic against the oracle on machines without a GPU.  It is not
// linked into libnele_score.so and is not a CPU path of the product.
//
// usage: ear_emul <in.bin> <out.bin> <f32|f64>
//   in : int32 N, f64 xmid[N], f64 ymid[N]
//   out: f64 bw[2][32], i32 shift[32], i32 nsub, f32 envlp[2][nsub][32]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../nele_gan_b200/csrc/host_tables.hpp"

using namespace nele;

template <typename T>
static void run(const std::vector<double> mid[2], int N, FILE* fo) {
  BandConst bc[kBands];
  host::make_band_consts(nullptr, bc);
  IhcConst ih = make_ihc_const();
  float fir[54];
  make_env_fir(fir);
  double bw[2][kBands];
  for (int q = 0; q < 2; ++q)
    for (int b = 0; b < kBands; ++b) {
      ControlLane<T> cl;
      cl.init(bc[b]);
      cl.car.seed_before(0);
      double acc = 0.0;
      for (int t = 0; t < N; ++t) {
        if (t % 256 == 0) cl.car.seed_before(t);
        acc += (double)cl.step((T)mid[q][t]);
      }
      bw[q][b] = bw_from_control(acc, cl.k.gain, N, bc[b].bwmin[q], bc[b].bw1);
    }
  int shift[kBands];
  double gd[kBands], gmax = -1e300;
  for (int b = 0; b < kBands; ++b) {
    gd[b] = gt_group_delay(bw[0][b], bc[b].erb);
    if (gd[b] > gmax) gmax = gd[b];
  }
  for (int b = 0; b < kBands; ++b) shift[b] = (int)(gmax - gd[b]);
  const int nsub = (N + kDecim - 1) / kDecim;
  std::vector<float> out((size_t)2 * nsub * kBands, 0.f);
  for (int q = 0; q < 2; ++q)
    for (int b = 0; b < kBands; ++b) {
      EarLane<T> L;
      Carrier<T> car;
      car.init(bc[b].cf);
      L.init(bc[b], q, bw[q][b], ih);
      for (int blk = 0; blk < nsub + 2; ++blk) {
        float v[9];
        for (int p = 0; p < 9; ++p) {
          const int i = blk * 9 + p, t = i - shift[b];
          if (t >= 0 && (t % 576 == 0)) car.seed_before(t);
          if (t >= 0 && i < N) {
            car.advance();
            const T xs = (T)mid[q][t];
            v[p] = L.sample(xs * car.c, xs * car.s);
          } else {
            v[p] = 0.f;
          }
        }
        L.template accumulate<0>(v[0], fir); L.template accumulate<1>(v[1], fir);
        L.template accumulate<2>(v[2], fir); L.template accumulate<3>(v[3], fir);
        L.template accumulate<4>(v[4], fir); L.template accumulate<5>(v[5], fir);
        L.template accumulate<6>(v[6], fir); L.template accumulate<7>(v[7], fir);
        L.template accumulate<8>(v[8], fir);
        const float o = L.emit();
        const int j = blk - 2;
        if (j >= 0 && j < nsub) out[((size_t)q * nsub + j) * kBands + b] = o;
      }
    }
  fwrite(bw, sizeof(double), 2 * kBands, fo);
  fwrite(shift, sizeof(int), kBands, fo);
  fwrite(&nsub, sizeof(int), 1, fo);
  fwrite(out.data(), sizeof(float), out.size(), fo);
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) return 3;
  int N = 0;
  if (fread(&N, sizeof(int), 1, fi) != 1) return 4;
  std::vector<double> mid[2];
  for (int q = 0; q < 2; ++q) {
    mid[q].resize(N);
    if (fread(mid[q].data(), sizeof(double), N, fi) != (size_t)N) return 5;
  }
  fclose(fi);
  FILE* fo = fopen(argv[2], "wb");
  if (!fo) return 6;
  if (!strcmp(argv[3], "f32")) run<float>(mid, N, fo); else run<double>(mid, N, fo);
  fclose(fo);
  return 0;
}
