// Host-side check of to_sci (bedops_b200/csrc/fixed_exact.cuh) against glibc printf("%.*e").
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../../bedops_b200/csrc/fixed_exact.cuh"

static int render(double x, int prec, char* out) {
  bk::Sci s;
  if (!bk::to_sci(x, prec, s)) return -1;
  int n = 0;
  if (s.neg) out[n++] = '-';
  if (s.special) return n + sprintf(out + n, "%s", s.special == 1 ? "nan" : "inf");
  char d[32];
  sprintf(d, "%0*llu", prec + 1, (unsigned long long)s.digits);
  out[n++] = d[0];
  if (prec > 0) { out[n++] = '.'; memcpy(out + n, d + 1, prec); n += prec; }
  n += sprintf(out + n, "e%c%02d", s.exp10 < 0 ? '-' : '+', abs(s.exp10));
  return n;
}

int main(int argc, char** argv) {
  long            n = argc > 1 ? atol(argv[1]) : 2000000;
  std::mt19937_64 rng(7);
  long            bad = 0, refused = 0, checked = 0;
  char            a[128], b[128];
  auto check = [&](double x, int prec) {
    checked++;
    if (render(x, prec, a) < 0) { refused++; return; }
    snprintf(b, sizeof b, "%.*e", prec, x);
    if (strcmp(a, b) != 0) {
      if (bad < 20) fprintf(stderr, "MISMATCH x=%.17g prec=%d: got %s ref %s\n", x, prec, a, b);
      bad++;
    }
  };
  const double fixed[] = {0.0, -0.0, 1.0, 9.5, 9.999999, 99999.95, 0.5, 0.25, 0.125, 1e-7, 1e10, 123456789.0, 43.442622950819674,
                          1.0 / 3, 2.0 / 3, 1e15, 9.007199254740991e15, 4.5e18, 1e-20, 2.5e-7, 999999.5, 1e22 / 1e5, 0.000123456, 5e-5};
  for (double x : fixed)
    for (int p = 0; p <= 17; p++) check(x, p), check(-x, p);
  for (long i = 0; i < n; i++) {
    int    kind = (int)(rng() % 4);
    double x;
    if (kind == 0) x = (double)(rng() % 100000) / (double)(1 + rng() % 97);
    else if (kind == 1) x = std::ldexp((double)(rng() >> 11), (int)(rng() % 130) - 120);
    else if (kind == 2) x = std::pow(10.0, (int)(rng() % 38) - 20) * (1 + (double)(rng() % 1000) / 64.0);
    else x = (double)(rng() % 2000000) + (double)(rng() % 1024) / 1024.0;
    if (rng() & 1) x = -x;
    check(x, (int)(rng() % 18));
  }
  printf("checked %ld values: %ld mismatches, %ld refused (%.2f%%)\n", checked, bad, refused, 100.0 * refused / checked);
  return bad == 0 ? 0 : 1;
}
