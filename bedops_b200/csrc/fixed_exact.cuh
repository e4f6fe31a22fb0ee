// fixed_exact.cuh -- exact "%.<prec>f" decomposition of a double for the device BED writer.
//
// Replaces printf("%.<prec>lf") (utility/Formats.hpp:42-49, visitors/helpers/ProcessBedVisitorRow.hpp:152-171):
// the binary value is expanded exactly and rounded half-to-even on the exact remainder, which is what glibc does
// in the default rounding mode (SURVEY hard part 4).  Supported: 0 <= prec <= 18 and |x| < 2^63; to_fixed returns
// false otherwise and the caller raises BK_ERR_UNSUPPORTED (never a guessed digit).
// The same source compiles as plain C++ (tests/native/fixed_check.cpp drives it against glibc on the host).
#pragma once
#include <stdint.h>
#include "strtod_exact.cuh"  // BK_FN, mul64x64

namespace bk {

BK_FN uint64_t double_bits(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t b;
  __builtin_memcpy(&b, &x, 8);
  return b;
#endif
}

#ifdef __CUDACC__
__device__
#endif
static const uint64_t kPow10u[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
                                     100000000ull, 1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull,
                                     10000000000000ull, 100000000000000ull, 1000000000000000ull, 10000000000000000ull,
                                     100000000000000000ull, 1000000000000000000ull, 10000000000000000000ull};
BK_FN uint64_t pow10_u64(int k) { return kPow10u[k]; }  // 10^k, 0 <= k <= 19

struct Fixed {
  bool     neg;
  uint64_t ip;       // integer part
  uint64_t frac;     // fraction scaled by 10^prec, < 10^prec
  int      special;  // 0 finite, 1 nan, 2 inf
};

BK_FN bool to_fixed(double x, int prec, Fixed& f) {
  const uint64_t bits = double_bits(x);
  f.neg = bits >> 63;
  f.special = 0;
  const int e = (int)((bits >> 52) & 0x7FF);
  uint64_t  m = bits & ((1ull << 52) - 1);
  if (e == 0x7FF) {
    f.special = m ? 1 : 2;
    f.ip = f.frac = 0;
    return true;
  }
  if (prec > 18 || prec < 0) return false;
  int sh;  // |x| = m / 2^sh
  if (e == 0) {
    sh = 1074;
  } else {
    m |= 1ull << 52;
    sh = 1075 - e;
  }
  if (sh <= 0) {
    if (-sh > 10) return false;
    f.ip = m << (-sh);
    f.frac = 0;
    return true;
  }
  uint64_t fm;
  if (sh < 64) {
    f.ip = m >> sh;
    fm = m & ((1ull << sh) - 1);
  } else {
    f.ip = 0;
    fm = m;
  }
  const uint64_t pw = pow10_u64(prec);
  uint64_t       lo, hi;
  mul64x64(fm, pw, hi, lo);
  uint64_t q;
  int      cmp;  // remainder vs half: -1 below, 0 tie, +1 above
  if (sh > 127) {
    q = 0;
    cmp = -1;
  } else if (sh < 64) {
    q = (lo >> sh) | (hi << (64 - sh));
    uint64_t r = lo & ((1ull << sh) - 1), half = 1ull << (sh - 1);
    cmp = r > half ? 1 : (r == half ? 0 : -1);
  } else if (sh == 64) {
    q = hi;
    uint64_t half = 1ull << 63;
    cmp = lo > half ? 1 : (lo == half ? 0 : -1);
  } else {
    q = hi >> (sh - 64);
    uint64_t rh = hi & ((1ull << (sh - 64)) - 1), halfh = 1ull << (sh - 65);
    cmp = rh > halfh ? 1 : (rh < halfh ? -1 : (lo ? 1 : 0));
  }
  // ties go to the even last printed digit: that digit lives in q, or in the integer part when prec == 0
  if (cmp > 0 || (cmp == 0 && ((prec > 0 ? q : f.ip) & 1))) q++;
  if (q >= pw) {
    q -= pw;
    f.ip++;
  }
  f.frac = q;
  return true;
}

// Length of the "%.<prec>f" text without producing a digit, for the emitter's length pass.  -1 = not decidable
// cheaply (a carry into the integer part is possible, or the value is large/special): the caller runs to_fixed.
// A carry needs fraction >= 1 - 0.5*10^-prec >= 0.95 when prec >= 1.
BK_FN int fixed_len_fast(double x, int prec) {
  const uint64_t bits = double_bits(x);
  const double   ax = bits_to_double(bits & 0x7FFFFFFFFFFFFFFFull);
  if (prec < 1 || prec > 18 || !(ax < 4.0e9)) return -1;
  const uint32_t ip = (uint32_t)ax;          // truncation
  if (ax - (double)ip >= 0.9) return -1;     // exact subtraction
  int nd = 1;
  for (uint32_t t = ip; t >= 10u; t /= 10u) nd++;
  return (int)(bits >> 63) + nd + 1 + prec;
}

// ---- exact "%.<prec>e" -----------------------------------------------------------------------------------------------
// d.ddd...e+XX of the exact binary value, rounded half-to-even like glibc.  Supported: 0 <= prec <= 17 and
// 2^-75 <= |x| < 2^63 (or x == 0); returns false otherwise (caller raises BK_ERR_UNSUPPORTED).
struct Sci {
  bool     neg;
  uint64_t digits;   // prec+1 significant digits
  int      exp10;
  int      special;  // 0 finite, 1 nan, 2 inf
};

// scaled = |x| * 10^k rounded half-to-even, x = m / 2^sh (sh >= 0) or m * 2^(-sh) (sh < 0).  false if out of range.
BK_FN bool scaled_round(uint64_t m, int sh, int k, uint64_t& out) {
  if (sh <= 0) {  // integer value V = m << -sh  (< 2^63 by the caller's range check)
    const uint64_t V = m << (-sh);
    if (k >= 0) {
      if (k > 19) return false;
      const uint64_t pw = pow10_u64(k);
      uint64_t       hi, lo;
      mul64x64(V, pw, hi, lo);
      if (hi) return false;
      out = lo;
      return true;
    }
    if (-k > 19) { out = 0; return true; }
    const uint64_t D = pow10_u64(-k), q = V / D, r = V % D;
    out = q + ((r > D - r || (r == D - r && (q & 1))) ? 1 : 0);
    return true;
  }
  if (k >= 0) {  // P = m * 10^k (up to 192 bits), out = round(P / 2^sh)
    if (k > 38 || sh > 127) return false;
    // 10^k as (ph, pl): 10^k = 10^a * 10^b with a, b <= 19
    const int      a = k > 19 ? 19 : k, b = k - a;
    uint64_t       ph, pl;
    mul64x64(pow10_u64(a), pow10_u64(b), ph, pl);
    // P = m * (ph:pl)  -> limbs p2:p1:p0
    uint64_t c1, p0, t1, t0;
    mul64x64(m, pl, c1, p0);
    mul64x64(m, ph, t1, t0);
    uint64_t p1 = t0 + c1;
    uint64_t p2 = t1 + (p1 < t0 ? 1 : 0);
    // shift right by sh (1..127) with exact remainder compare
    uint64_t q0, q1, q2;  // quotient limbs
    uint64_t r_hi, r_lo, h_hi, h_lo;  // remainder and half (2^(sh-1)), 128-bit
    if (sh < 64) {
      q0 = (p0 >> sh) | (p1 << (64 - sh));
      q1 = (p1 >> sh) | (p2 << (64 - sh));
      q2 = p2 >> sh;
      r_hi = 0; r_lo = p0 & ((1ull << sh) - 1);
      h_hi = 0; h_lo = 1ull << (sh - 1);
    } else if (sh == 64) {
      q0 = p1; q1 = p2; q2 = 0;
      r_hi = 0; r_lo = p0;
      h_hi = 0; h_lo = 1ull << 63;
    } else {
      const int t = sh - 64;  // 1..63
      q0 = (p1 >> t) | (p2 << (64 - t));
      q1 = p2 >> t;
      q2 = 0;
      r_hi = p1 & ((1ull << t) - 1); r_lo = p0;
      h_hi = 1ull << (t - 1); h_lo = 0;
    }
    if (q1 || q2) return false;
    const int cmp = r_hi > h_hi ? 1 : (r_hi < h_hi ? -1 : (r_lo > h_lo ? 1 : (r_lo < h_lo ? -1 : 0)));
    out = q0 + ((cmp > 0 || (cmp == 0 && (q0 & 1))) ? 1 : 0);
    return true;
  }
  // k < 0: out = round(m / (10^-k * 2^sh))
  if (-k > 19 || sh > 63) { out = 0; return sh <= 63 ? true : false; }
  uint64_t dh, dl;
  {
    const uint64_t D = pow10_u64(-k);
    dl = D << sh;
    dh = sh ? (D >> (64 - sh)) : 0;
  }
  if (dh) { out = 0; return true; }  // divisor > m: rounds to 0 or 1 -- the caller's digit-count check re-scales
  const uint64_t q = m / dl, r = m % dl;
  out = q + ((r > dl - r || (r == dl - r && (q & 1))) ? 1 : 0);
  return true;
}

BK_FN int ndigits10_u64(uint64_t v) {
  int n = 1;
  while (v >= 10) { v /= 10; n++; }
  return n;
}

BK_FN bool to_sci(double x, int prec, Sci& s) {
  const uint64_t bits = double_bits(x);
  s.neg = bits >> 63;
  s.special = 0;
  s.digits = 0;
  s.exp10 = 0;
  const int e = (int)((bits >> 52) & 0x7FF);
  uint64_t  m = bits & ((1ull << 52) - 1);
  if (e == 0x7FF) {
    s.special = m ? 1 : 2;
    return true;
  }
  if (prec < 0 || prec > 17) return false;
  if (e == 0 && m == 0) return true;  // +-0 -> 0.000e+00
  if (e == 0) return false;           // subnormal: outside the supported range
  m |= 1ull << 52;
  const int sh = 1075 - e;            // |x| = m / 2^sh
  if (sh < -10 || sh > 127) return false;
  // first guess of floor(log10|x|) from the binary exponent: |x| in [2^(52-sh), 2^(53-sh))
  int E = (int)(((long long)(52 - sh) * 1233) >> 12);  // floor((52-sh) * log10(2)) within one
  if (52 - sh < 0) E = -(int)((((long long)(sh - 52)) * 1233 + 4095) >> 12);
  const uint64_t lo_lim = pow10_u64(prec), hi_lim = lo_lim * 10;  // digits must lie in [10^prec, 10^(prec+1))
  for (int it = 0; it < 4; it++) {
    uint64_t N;
    if (!scaled_round(m, sh, prec - E, N)) return false;
    if (N >= hi_lim) {
      if (N == hi_lim) {  // rounded up to the next power of ten
        s.digits = lo_lim;
        s.exp10 = E + 1;
        return true;
      }
      E++;
    } else if (N < lo_lim) {
      E--;
    } else {
      s.digits = N;
      s.exp10 = E;
      return true;
    }
  }
  return false;
}

}  // namespace bk
