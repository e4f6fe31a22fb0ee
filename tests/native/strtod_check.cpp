// Host-side check of bedops_b200/csrc/strtod_exact.cuh (the same source the device parser compiles) against
// glibc strtod on random and adversarial decimal literals.  Exit 0 = every accepted literal is bit-identical
// to strtod and the refusal rate is within bounds.  Built and run by tests/test_strtod_host.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include "../../bedops_b200/csrc/strtod_exact.cuh"

struct Cur {
  const char* p;
  unsigned char at(int64_t q) const { return (unsigned char)p[q]; }
};

int main(int argc, char** argv) {
  long          n = argc > 1 ? atol(argv[1]) : 2000000;
  std::mt19937_64 rng(12345);
  long          bad = 0, refused = 0, checked = 0;
  auto check = [&](const std::string& s) {
    std::string t = s + "\n";
    Cur         c{t.c_str()};
    int64_t     q = 0;
    double      v = 0;
    int         rc = bk::parse_decimal(c, q, v);
    char*       endp = nullptr;
    double      ref = strtod(t.c_str(), &endp);
    checked++;
    if (rc == 6) { refused++; return; }
    if (rc != 0) { if (endp != t.c_str()) { bad++; fprintf(stderr, "refused-as-parse-error: %s\n", s.c_str()); } return; }
    if (memcmp(&v, &ref, 8) != 0 || (t.c_str() + q) != endp) {
      if (bad < 20) fprintf(stderr, "MISMATCH %s: got %.17g ref %.17g (consumed %ld vs %ld)\n", s.c_str(), v, ref, (long)q, (long)(endp - t.c_str()));
      bad++;
    }
  };
  const char* fixed[] = {"0", "-0", "0.0", "1", "1e3", "-0.000001", "123456789012345.678", "+.5", "7.", "1e", "1e+", "1.5e-3x",
                         "9007199254740993", "9007199254740992.5", "4.9e-324", "2.2250738585072014e-308", "2.2250738585072011e-308",
                         "1.7976931348623157e308", "1.7976931348623159e308", "1e309", "1e-400", "0.000000000000000000000000001",
                         "123456789012345678901234567890", "0.30000000000000004", "8.98846567431158e307", "1e23", "8.5e22",
                         "9.5e-5", "5e-324", "2.4703282292062327e-324", "2.4703282292062328e-324", "6.0221409e+23",
                         "1.00000000000000011102230246251565404236316680908203125", "1.00000000000000011102230246251565404236316680908203124",
                         "1.00000000000000011102230246251565404236316680908203126", "4.20586e-06", "2.21622e-06"};
  for (const char* f : fixed) check(f);
  char buf[128];
  for (long i = 0; i < n; i++) {
    int kind = (int)(rng() % 6);
    if (kind == 0) {  // %.6f style scores
      double v = std::ldexp((double)(rng() >> 11), -53) * std::pow(10.0, (int)(rng() % 12) - 3);
      snprintf(buf, sizeof buf, "%s%.6f", (rng() & 1) ? "-" : "", v);
    } else if (kind == 1) {  // random bit patterns printed with 17 significant digits
      uint64_t b = rng();
      double   v;
      memcpy(&v, &b, 8);
      if (!(v == v) || std::isinf(v)) continue;
      snprintf(buf, sizeof buf, "%.17g", v);
    } else if (kind == 2) {  // shortest-ish (%g)
      double v = std::ldexp((double)(rng() >> 11), (int)(rng() % 200) - 100);
      snprintf(buf, sizeof buf, "%.*g", 1 + (int)(rng() % 17), v);
    } else if (kind == 3) {  // long digit strings
      int  nd = 1 + (int)(rng() % 30);
      int  dot = (int)(rng() % (nd + 1));
      int  k = 0;
      for (int d = 0; d < nd; d++) {
        if (d == dot) buf[k++] = '.';
        buf[k++] = (char)('0' + rng() % 10);
      }
      if (rng() % 3 == 0) k += snprintf(buf + k, sizeof buf - k, "e%d", (int)(rng() % 600) - 300);
      buf[k] = 0;
    } else if (kind == 4) {  // integers
      snprintf(buf, sizeof buf, "%llu", (unsigned long long)(rng() >> (rng() % 64)));
    } else {  // halfway cases near 2^53..2^63
      uint64_t m = (1ull << 53) + (rng() % (1ull << 20)) * 2 + 1;
      int      sh = (int)(rng() % 10);
      snprintf(buf, sizeof buf, "%llu", (unsigned long long)(m << sh));
    }
    check(buf);
  }
  printf("checked %ld literals: %ld mismatches, %ld refused (%.4f%%)\n", checked, bad, refused, 100.0 * refused / checked);
  return (bad == 0 && refused * 200 < checked) ? 0 : 1;
}
