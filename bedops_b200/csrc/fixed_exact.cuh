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

BK_FN uint64_t pow10_u64(int k) {  // 10^k, 0 <= k <= 19
  uint64_t p = 1;
  for (int i = 0; i < k; i++) p *= 10;
  return p;
}

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

}  // namespace bk
