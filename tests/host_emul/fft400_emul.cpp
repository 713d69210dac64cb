// Host check of the in-place 400-point warp FFT (fft400.cuh) against a direct float64 DFT.  TEST TOOL ONLY:
// compiled by tests/test_host_emul.py with g++; not linked into libnele_score.so and not a CPU path of the product.
// The "lanes" run one after the other here, which is legal because no lane writes a position another lane of the
// same phase reads -- the property the single-buffer layout rests on; the run also checks it explicitly by
// executing the lanes of each phase in reverse order and comparing.
//
// prints: MAXERR <largest |X - DFT| / max|DFT|>  ORDER <largest difference between the two lane orders>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../nele_gan_b200/csrc/fft400.cuh"

using namespace nele;

static void run(std::vector<cpx>& z, const cpx* tw, bool reverse) {
  // the two tables siib_spec_kernel cuts from w[k] = exp(-2 pi i k / 400)
  std::vector<cpx> twa(400), tw25(25);
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n2 = 0; n2 < 25; ++n2) twa[k1 * 25 + n2] = tw[n2 * k1];
  for (int m = 0; m < 25; ++m) tw25[m] = tw[16 * m];
  for (int i = 0; i < 25; ++i) fft400_phase_a(reverse ? 24 - i : i, z.data(), twa.data());
  for (int i = 0; i < 16; ++i) fft400_phase_b(reverse ? 15 - i : i, z.data(), tw25.data());
}

int main() {
  std::vector<cpx> tw(400);
  for (int k = 0; k < 400; ++k) tw[k] = {(float)cos(-2.0 * M_PI * k / 400.0), (float)sin(-2.0 * M_PI * k / 400.0)};
  double worst = 0.0, order = 0.0;
  srand(7);
  for (int trial = 0; trial < 8; ++trial) {
    std::vector<cpx> x(400);
    for (auto& v : x) v = {(float)(rand() / (double)RAND_MAX - 0.5), (float)(rand() / (double)RAND_MAX - 0.5)};
    if (trial == 0)
      for (int n = 0; n < 400; ++n) x[n] = {n == 3 ? 1.f : 0.f, 0.f};  // a shifted impulse: every bin has magnitude 1
    std::vector<cpx> z = x, zr = x;
    run(z, tw.data(), false);
    run(zr, tw.data(), true);
    double peak = 0.0;
    std::vector<double> re(400), im(400);
    for (int k = 0; k < 400; ++k) {
      double sr = 0.0, si = 0.0;
      for (int n = 0; n < 400; ++n) {
        const double a = -2.0 * M_PI * (double)((n * k) % 400) / 400.0;
        sr += x[n].x * cos(a) - x[n].y * sin(a);
        si += x[n].x * sin(a) + x[n].y * cos(a);
      }
      re[k] = sr;
      im[k] = si;
      peak = fmax(peak, hypot(sr, si));
    }
    for (int k = 0; k < 400; ++k) {
      const cpx v = z[fft400_pos(k)], w = zr[fft400_pos(k)];
      worst = fmax(worst, hypot(v.x - re[k], v.y - im[k]) / peak);
      order = fmax(order, hypot((double)v.x - w.x, (double)v.y - w.y));
    }
  }
  printf("MAXERR %.3e ORDER %.3e\n", worst, order);
  return 0;
}
