// Host check of the two-wide main-pass lane (EarLane2, ear_core.cuh) against two scalar lanes
// (EarLane<float>), on a synthetic amplitude-modulated tone per band.  TEST TOOL ONLY: compiled by
// tests/test_host_emul.py with g++; not linked into libnele_score.so and not a CPU path of the product.
// Prints, per band, the largest absolute difference of the decimated envelopes (dB SL) over the run.
//
// usage: ear_x2_emul <n_samples>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../nele_gan_b200/csrc/host_tables.hpp"

using namespace nele;

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 9000;
  BandConst bc[kBands];
  host::make_band_consts(nullptr, bc);
  const IhcConst ih = make_ihc_const();
  float fir[54];
  make_env_fir(fir);
  double worst = 0.0;
  for (int b = 0; b < kBands; b += 3) {
    // different bandwidths for the two signals, as eb_BWadjust produces them
    const double bwx = bc[b].bwmin[0] * 1.00, bwy = bc[b].bwmin[1] * 1.37;
    EarLane<float> Lx, Ly;
    EarLane2 L2;
    Carrier<float> car;
    car.init(bc[b].cf);
    Lx.init(bc[b], 0, bwx, ih);
    Ly.init(bc[b], 1, bwy, ih);
    L2.init(bc[b], bwx, bwy, ih);
    double dmax = 0.0, omax = 0.0;
    for (int blk = 0; blk * 9 + 9 <= N; ++blk) {
      float vx[9], vy[9];
      F2 v2[9];
      for (int p = 0; p < 9; ++p) {
        const int t = blk * 9 + p;
        if (t % 576 == 0) car.seed_before(t);
        car.advance();
        const double env = 0.5 + 0.5 * sin(2.0 * 3.14159265358979 * 4.0 * t / 24000.0);
        const float xs = (float)(0.8 * env * sin(2.0 * 3.14159265358979 * bc[b].cf * t / 24000.0) + 0.01 * sin(0.37 * t));
        const float ys = (float)(0.5 * xs + 0.05 * sin(0.11 * t + 1.0));
        vx[p] = Lx.sample(xs * car.c, xs * car.s);
        vy[p] = Ly.sample(ys * car.c, ys * car.s);
        const F2 XY = f2_pack(xs, ys);
        v2[p] = L2.sample(f2_mul(XY, f2_pack(car.c, car.c)), f2_mul(XY, f2_pack(car.s, car.s)));
      }
      Lx.accumulate<0>(vx[0], fir); Ly.accumulate<0>(vy[0], fir); L2.accumulate<0>(v2[0], fir);
      Lx.accumulate<1>(vx[1], fir); Ly.accumulate<1>(vy[1], fir); L2.accumulate<1>(v2[1], fir);
      Lx.accumulate<2>(vx[2], fir); Ly.accumulate<2>(vy[2], fir); L2.accumulate<2>(v2[2], fir);
      Lx.accumulate<3>(vx[3], fir); Ly.accumulate<3>(vy[3], fir); L2.accumulate<3>(v2[3], fir);
      Lx.accumulate<4>(vx[4], fir); Ly.accumulate<4>(vy[4], fir); L2.accumulate<4>(v2[4], fir);
      Lx.accumulate<5>(vx[5], fir); Ly.accumulate<5>(vy[5], fir); L2.accumulate<5>(v2[5], fir);
      Lx.accumulate<6>(vx[6], fir); Ly.accumulate<6>(vy[6], fir); L2.accumulate<6>(v2[6], fir);
      Lx.accumulate<7>(vx[7], fir); Ly.accumulate<7>(vy[7], fir); L2.accumulate<7>(v2[7], fir);
      Lx.accumulate<8>(vx[8], fir); Ly.accumulate<8>(vy[8], fir); L2.accumulate<8>(v2[8], fir);
      const float ox = Lx.emit(), oy = Ly.emit();
      float px, py;
      f2_unpack(L2.emit(), px, py);
      dmax = fmax(dmax, fmax(fabs((double)ox - px), fabs((double)oy - py)));
      omax = fmax(omax, fmax(fabs((double)ox), fabs((double)oy)));
    }
    printf("band %2d  max |diff| %.3e dB  (max level %.2f dB SL)\n", b, dmax, omax);
    worst = fmax(worst, dmax);
  }
  printf("WORST %.6e\n", worst);
  return 0;
}
