// Host check of the Sturm count of sturm.cuh (sign masks + popc, 16-step blocks over a padded table, exponent-field
// range guard) against the classic ratio-form count in long double, and of the bisection plan of the two kernels
// (one count per thread on a uniform grid, binary search of the counts, then k bisection steps inside the cell)
// against plain bisection from the whole interval.  TEST TOOL ONLY: compiled by tests/test_host_emul.py with g++; not
// linked into libnele_score.so and not a CPU path of the product.
//
// prints, per matrix: <name> n=<n> points=<tested> skipped=<points within rounding of an eigenvalue> MISMATCH=<k> EIGERR=<max>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../nele_gan_b200/csrc/sturm.cuh"

using namespace nele;

// eigenvalues of T below x, LAPACK dlaebz style: q_i = (d_i - x) - e_{i-1}^2 / q_{i-1}, negatives counted
static int ref_count(const std::vector<double>& d, const std::vector<double>& e, long double x) {
  const long double pivmin = 1e-300L;
  long double q = (long double)d[0] - x;
  if (fabsl(q) < pivmin) q = -pivmin;
  int cnt = q < 0 ? 1 : 0;
  for (size_t i = 1; i < d.size(); ++i) {
    q = ((long double)d[i] - x) - (long double)e[i - 1] * (long double)e[i - 1] / q;
    if (fabsl(q) < pivmin) q = -pivmin;
    cnt += q < 0 ? 1 : 0;
  }
  return cnt;
}

template <int N>
static int run(const char* name, std::vector<double> d, std::vector<double> e) {
  constexpr int LEN = sturm_len(N);
  // scale as the kernels do: max |entry| = 1
  double mx = 0.0;
  for (int i = 0; i < N; ++i) mx = fmax(mx, fmax(fabs(d[i]), i < N - 1 ? fabs(e[i]) : 0.0));
  for (int i = 0; i < N; ++i) {
    d[i] /= mx;
    if (i < N - 1) e[i] /= mx;
  }
  std::vector<sturm_pair> de(LEN);
  for (int i = 0; i < LEN; ++i) de[i] = {1000.0, 0.0};
  for (int i = 0; i < N; ++i) {
    de[i].x = d[i];
    de[i].y = i ? e[i - 1] * e[i - 1] : 0.0;
  }
  // ---- counts on a grid, on points crowding zero (clusters of tiny eigenvalues), and at the interval ends
  std::vector<double> xs;
  for (int k = 0; k <= 600; ++k) xs.push_back(-3.0 + 6.0 * k / 600.0 + 1.234e-4);
  for (int k = 0; k < 60; ++k) {
    xs.push_back(pow(10.0, -0.25 * k));
    xs.push_back(-pow(10.0, -0.25 * k));
  }
  int mismatch = 0, skipped = 0, tested = 0;
  for (double x : xs) {
    if (x < -3.0 || x > 3.0) continue;
    const int lo = ref_count(d, e, (long double)x - fabsl(x) * 1e-9L - 1e-15L);
    const int hi = ref_count(d, e, (long double)x + fabsl(x) * 1e-9L + 1e-15L);
    if (lo != hi) {  // an eigenvalue within rounding of x: either answer is right
      ++skipped;
      continue;
    }
    ++tested;
    if (sturm_count<LEN>(de.data(), x) != lo) ++mismatch;
  }
  // ---- the bisection plan of the kernels against plain bisection with the reference count
  const double cell = 6.0 / N;
  std::vector<int> grid(N + 1);
  for (int t = 0; t < N; ++t) grid[t] = t ? sturm_count<LEN>(de.data(), -3.0 + cell * t) : 0;
  grid[N] = N;
  const int steps = N == 420 ? 38 : 52;   // siib_trieig_kernel / siib_smallvec_kernel
  double eigerr = 0.0;
  for (int want = 0; want < N; ++want) {
    int klo = 0, khi = N;
    while (khi - klo > 1) {
      const int km = (klo + khi) >> 1;
      if (grid[km] > want) khi = km;
      else klo = km;
    }
    double lo = -3.0 + cell * klo, hi = -3.0 + cell * khi;
    for (int it = 0; it < steps; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (sturm_count<LEN>(de.data(), mid) > want) hi = mid;
      else lo = mid;
    }
    long double rl = -3.0L, rh = 3.0L;
    for (int it = 0; it < 70; ++it) {
      const long double mid = 0.5L * (rl + rh);
      if (ref_count(d, e, mid) > want) rh = mid;
      else rl = mid;
    }
    eigerr = fmax(eigerr, fabs(0.5 * (lo + hi) - (double)(0.5L * (rl + rh))));
  }
  printf("%s n=%d points=%d skipped=%d MISMATCH=%d EIGERR=%.3e\n", name, N, tested, skipped, mismatch, eigerr);
  return mismatch;
}

static double urand() { return rand() / (double)RAND_MAX; }

template <int N>
static int all_cases() {
  int bad = 0;
  std::vector<double> d(N), e(N - 1);
  // dense spectrum
  for (int i = 0; i < N; ++i) d[i] = 2.0 * urand() - 1.0;
  for (int i = 0; i < N - 1; ++i) e[i] = urand();
  bad += run<N>("random", d, e);
  // graded over 14 decades, the shape of a covariance tridiagonalised from its largest entries down
  for (int i = 0; i < N; ++i) d[i] = pow(10.0, -14.0 * i / N);
  for (int i = 0; i < N - 1; ++i) e[i] = 0.3 * sqrt(d[i] * d[i + 1]);
  bad += run<N>("graded", d, e);
  // a numerically rank-deficient tail: the last third is rounding noise with off-diagonals near 1e-15
  for (int i = 0; i < N; ++i) d[i] = i < 2 * N / 3 ? 1.0 / (1.0 + i) : 1e-16 * (urand() - 0.5);
  for (int i = 0; i < N - 1; ++i) e[i] = i < 2 * N / 3 ? 0.1 / (1.0 + i) : 1e-15 * urand();
  bad += run<N>("null-tail", d, e);
  // Toeplitz (0, 1/2): eigenvalues cos(k pi / (n + 1)), the recurrence neither grows nor decays
  for (int i = 0; i < N; ++i) d[i] = 0.0;
  for (int i = 0; i < N - 1; ++i) e[i] = 0.5;
  bad += run<N>("toeplitz", d, e);
  // tiny couplings everywhere: p shrinks by ~1e-10 per step around x = 0 (the underflow side of the range guard)
  for (int i = 0; i < N; ++i) d[i] = (i % 7 == 0) ? 1.0 : 1e-10 * (urand() - 0.5);
  for (int i = 0; i < N - 1; ++i) e[i] = 1e-10 * urand();
  bad += run<N>("decoupled", d, e);
  // wide spectrum with couplings near 1: p grows by up to 7 per step at the interval ends (the overflow side)
  for (int i = 0; i < N; ++i) d[i] = (i & 1) ? 1.0 : -1.0;
  for (int i = 0; i < N - 1; ++i) e[i] = 1.0;
  bad += run<N>("growth", d, e);
  return bad;
}

int main() {
  srand(11);
  const int bad = all_cases<420>() + all_cases<112>();
  printf("TOTAL MISMATCH %d\n", bad);
  return 0;
}
