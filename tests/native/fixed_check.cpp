// Host-side check of bedops_b200/csrc/fixed_exact.cuh (the source the device emitter compiles) against glibc
// printf("%.*f") on random doubles and exact-tie cases, precisions 0..18.  Exit 0 = byte-identical everywhere.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../../bedops_b200/csrc/fixed_exact.cuh"

static int render(double x, int prec, char* out) {
  bk::Fixed f;
  if (!bk::to_fixed(x, prec, f)) return -1;
  int n = 0;
  if (f.neg) out[n++] = '-';
  if (f.special) return n + sprintf(out + n, "%s", f.special == 1 ? "nan" : "inf");
  n += sprintf(out + n, "%llu", (unsigned long long)f.ip);
  if (prec > 0) n += sprintf(out + n, ".%0*llu", prec, (unsigned long long)f.frac);
  return n;
}

int main(int argc, char** argv) {
  long            n = argc > 1 ? atol(argv[1]) : 2000000;
  std::mt19937_64 rng(99);
  long            bad = 0, refused = 0, checked = 0;
  char            a[512], b[512];
  auto check = [&](double x, int prec) {
    checked++;
    if (render(x, prec, a) < 0) { refused++; return; }
    snprintf(b, sizeof b, "%.*f", prec, x);
    const int fl = bk::fixed_len_fast(x, prec);  // the length pass's shortcut must agree whenever it answers
    if (fl >= 0 && fl != (int)strlen(b)) {
      if (bad < 20) fprintf(stderr, "LENGTH MISMATCH x=%.17g prec=%d: got %d ref %s\n", x, prec, fl, b);
      bad++;
    }
    if (strcmp(a, b) != 0) {
      if (bad < 20) fprintf(stderr, "MISMATCH x=%.17g prec=%d: got %s ref %s\n", x, prec, a, b);
      bad++;
    }
  };
  const double fixed[] = {0.0, -0.0, 0.5, 1.5, 2.5, 349.5, 350.5, 0.125, 0.375, 1e-7, -1e-9, 5e-324, 0.9999995, 0.99999949999999,
                          9.9999995, 999999.9999995, 43.442622950819674, 1.0 / 3, 2.0 / 3, 1e15 + 0.5, 4503599627370497.5,
                          9007199254740991.0, 9.2233720368547748e18, 123456.7890125, 0.0000005, 0.0000015, 0.0000025};
  for (double x : fixed)
    for (int p = 0; p <= 18; p++) check(x, p), check(-x, p);
  for (long i = 0; i < n; i++) {
    int    kind = (int)(rng() % 5);
    double x;
    if (kind == 0) {  // sums / means of integer scores
      x = (double)(rng() % 100000) / (double)(1 + rng() % 97);
    } else if (kind == 1) {  // exact binary ties at some decimal place
      x = (double)(rng() % 2000000) + (double)(rng() % 1024) / 1024.0;
    } else if (kind == 2) {  // random magnitudes below 2^63
      x = std::ldexp((double)(rng() >> 11), (int)(rng() % 120) - 110);
    } else if (kind == 3) {  // values printed and re-read at 6 decimals (typical score text)
      x = std::round(std::ldexp((double)(rng() >> 11), -40) * 1e6) / 1e6;
    } else {  // tiny and subnormal
      uint64_t bits = rng() >> (1 + rng() % 12);
      memcpy(&x, &bits, 8);
      if (!(x == x) || std::isinf(x) || std::fabs(x) >= 9.2e18) continue;
    }
    if (rng() & 1) x = -x;
    check(x, (int)(rng() % 19));
  }
  printf("checked %ld values: %ld mismatches, %ld refused\n", checked, bad, refused);
  return bad == 0 ? 0 : 1;
}
