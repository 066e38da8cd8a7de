// Host-side check of the digit scheme of csrc/ozaki.cuh (compiled with nvcc, runs on the CPU: no GPU needed).
//   * quantise / digit_bytes / digit: the seven balanced base-256 digits reconstruct the rounded value exactly,
//   * scale_exponent keeps every operand inside the representable range and wastes at most two bits,
//   * tile_off is a bijection onto the [rows x 32 B] tile for both images,
//   * an Ozaki dot product (28 exact integer level sums, FP64 recombination) matches long double to one FP64 rounding.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>
#include "ozaki.cuh"
using namespace nls;

static double urand() { return rand() / (double)RAND_MAX; }

int main() {
  int bad = 0;
  srand(7);
  for (int it = 0; it < 200000; ++it) {
    const int e = rand() % 80 - 40;
    const double amax = ldexp(0.5 + 0.5 * urand(), e);
    const int ex = oz::scale_exponent(amax);
    const double top = amax * ldexp(1.0, -ex);
    if (!(top <= 0.495 && top > 0.12)) { if (bad++ < 5) printf("exponent: amax %.17g -> e %d, scaled %.6f\n", amax, ex, top); }
    const double x = (2.0 * urand() - 1.0) * amax;
    const long long v = oz::quantise(x, ldexp(1.0, oz::FRAC_BITS - ex));
    const unsigned long long bytes = oz::digit_bytes(v);
    long long back = 0;
    double xr = 0.0;
    for (int p = 0; p < oz::S; ++p) {
      const int d = oz::digit(bytes, p);
      if (d < -128 || d > 127) ++bad;
      back = back * 256 + d;
      xr += d * ldexp(1.0, -oz::RADIX_BITS * (p + 1));
    }
    if (back != v) { if (bad++ < 5) printf("digits: v %lld back %lld\n", v, back); }
    if (fabs(ldexp(xr, ex) - x) > ldexp(1.0, ex - oz::FRAC_BITS - 1) * 1.0000001) { if (bad++ < 5) printf("reconstruction: x %.17g xr %.17g\n", x, ldexp(xr, ex)); }
  }
  if (oz::scale_exponent(0.0) != 0) ++bad;
  for (int t = 0; t < oz::S; ++t)
    if (oz::level_weight(t) != ldexp(1.0, -oz::RADIX_BITS * (t + 2))) ++bad;
  {  // tile images
    std::set<unsigned> a, b;
    for (int r = 0; r < oz::TM; ++r)
      for (int c = 0; c < 2; ++c) {
        a.insert(oz::tile_off<6>(r, c));
        b.insert(oz::tile_off<0>(r, c));
        if (oz::tile_off<6>(r, c) % 16 || oz::tile_off<6>(r, c) >= (unsigned)oz::A_TILE || oz::tile_off<0>(r, c) >= (unsigned)oz::A_TILE) ++bad;
      }
    if (a.size() != 2u * oz::TM || b.size() != 2u * oz::TM) ++bad;
  }
  if (oz::feature_ksteps(1024) != 64 || oz::feature_ksteps(100) != 8 || oz::feature_ksteps(4096) != 256) ++bad;
  {  // Ozaki dot products
    const int K = 2048;
    long double worst = 0;
    for (int rep = 0; rep < 200; ++rep) {
      static double a[K], b[K];
      double am = 0, bm = 0;
      for (int k = 0; k < K; ++k) {
        a[k] = (2 * urand() - 1) * 0.03125;
        b[k] = (2 * urand() - 1) * pow(10.0, -4.0 * urand());
        am = fmax(am, fabs(a[k]));
        bm = fmax(bm, fabs(b[k]));
      }
      const int ea = oz::scale_exponent(am), eb = oz::scale_exponent(bm);
      long long lev[oz::S] = {0};
      long double ref = 0, bound = 0;
      for (int k = 0; k < K; ++k) {
        const unsigned long long qa = oz::digit_bytes(oz::quantise(a[k], ldexp(1.0, oz::FRAC_BITS - ea)));
        const unsigned long long qb = oz::digit_bytes(oz::quantise(b[k], ldexp(1.0, oz::FRAC_BITS - eb)));
        for (int p = 0; p < oz::S; ++p)
          for (int q = 0; p + q < oz::S; ++q) lev[p + q] += (long long)oz::digit(qa, p) * oz::digit(qb, q);
        ref += (long double)a[k] * b[k];
        bound += fabsl((long double)a[k] * b[k]);
      }
      double sum = 0;
      for (int t = oz::S - 1; t >= 0; --t) {
        if (llabs(lev[t]) >= (1LL << 31)) ++bad;  // must fit the INT32 accumulators
        sum = fma((double)lev[t], oz::level_weight(t), sum);
      }
      worst = fmaxl(worst, fabsl((long double)ldexp(sum, ea + eb) - ref) / bound);
    }
    printf("ozaki dot products: max |err| / sum|a b| = %.3Le\n", worst);
    if (worst > 1.2e-16L) ++bad;  // the one FP64 rounding of the recombined sum; an FP64 FMA loop over K = 2048 is at ~1e-15
  }
  printf(bad ? "FAILED %d\n" : "ok\n", bad);
  return bad != 0;
}
