// strtod_exact.cuh -- correctly rounded decimal -> double for the device BED parser (column 5, "%lf").
//
// Replaces glibc strtod behind fscanf("%lf") (Bed.hpp:829-860).  Three stages, each exact or refused:
//   1. Clinger fast path: mantissa < 2^53 and |exp10| <= 22 -> one correctly rounded multiply/divide;
//   2. Eisel-Lemire: 64-bit mantissa x 128-bit power of five, with the published error analysis deciding
//      when the 128-bit product is enough;
//   3. literals with more than 19 significant digits: accepted when the truncated mantissa w and w+1 round
//      to the same double, refused (BK_ERR_UNSUPPORTED, never guessed) otherwise.
// The same source compiles as plain C++ (tests/strtod_check.cpp drives it against glibc on the host).
#pragma once
#include <stdint.h>
#include "pow5_table.cuh"

#ifdef __CUDACC__
#define BK_FN __device__ __forceinline__
#else
#define BK_FN static inline
#endif

namespace bk {

BK_FN void mul64x64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 p = (unsigned __int128)a * b;
  lo = (uint64_t)p;
  hi = (uint64_t)(p >> 64);
#endif
}

BK_FN int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}

BK_FN double bits_to_double(uint64_t b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  __builtin_memcpy(&d, &b, 8);
  return d;
#endif
}

#ifdef __CUDACC__
__device__
#endif
static const double kPow10d[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Eisel-Lemire: w * 10^q -> IEEE bits.  Returns false when the algorithm cannot decide (caller refuses).
BK_FN bool eisel_lemire(uint64_t w, int q, uint64_t& bits) {
  if (w == 0 || q < -342) {
    bits = 0;
    return true;
  }
  if (q > 308) {
    bits = 0x7FF0000000000000ull;
    return true;
  }
  int lz = clz64(w);
  w <<= lz;
  const uint64_t th = kPow5[2 * (q + 342)], tl = kPow5[2 * (q + 342) + 1];
  uint64_t       hi, lo;
  mul64x64(w, th, hi, lo);
  const uint64_t precision_mask = 0xFFFFFFFFFFFFFFFFull >> 55;  // 52 + 3 bits wanted
  if ((hi & precision_mask) == precision_mask) {
    uint64_t shi, slo;
    mul64x64(w, tl, shi, slo);
    lo += shi;
    if (shi > lo) hi++;
  }
  if (lo == 0xFFFFFFFFFFFFFFFFull) {
    if (!(q >= -27 && q <= 55)) return false;
  }
  const int upperbit = (int)(hi >> 63);
  uint64_t  mantissa = hi >> (upperbit + 64 - 52 - 3);
  int power2 = (((152170 + 65536) * q) >> 16) + 63 + upperbit - lz + 1023;
  if (power2 <= 0) {  // subnormal
    if (-power2 + 1 >= 64) {
      bits = 0;
      return true;
    }
    mantissa >>= -power2 + 1;
    mantissa += (mantissa & 1);
    mantissa >>= 1;
    power2 = (mantissa < (1ull << 52)) ? 0 : 1;
    bits = (mantissa & ~(1ull << 52)) | ((uint64_t)power2 << 52);
    return true;
  }
  if (lo <= 1 && q >= -4 && q <= 23 && (mantissa & 3) == 1) {
    if ((mantissa << (upperbit + 64 - 52 - 3)) == hi) mantissa &= ~1ull;  // exactly halfway: round to even
  }
  mantissa += (mantissa & 1);
  mantissa >>= 1;
  if (mantissa >= (2ull << 52)) {
    mantissa = 1ull << 52;
    power2++;
  }
  mantissa &= ~(1ull << 52);
  if (power2 >= 0x7FF) {
    bits = 0x7FF0000000000000ull;
    return true;
  }
  bits = mantissa | ((uint64_t)power2 << 52);
  return true;
}

// value = mant * 10^exp10 where mant holds the first (up to 19) significant digits; `truncated` = more non-zero
// digits followed.  Returns 0 and the double, or non-zero when the literal cannot be converted exactly here.
BK_FN int decimal_to_double(uint64_t mant, int exp10, bool truncated, double& out) {
  if (mant == 0) {
    out = 0.0;
    return 0;
  }
  if (!truncated && mant < (1ull << 53)) {  // Clinger
    const double* p10 = kPow10d;
    if (exp10 >= 0 && exp10 <= 22) {
      out = (double)mant * p10[exp10];
      return 0;
    }
    if (exp10 < 0 && exp10 >= -22) {
      out = (double)mant / p10[-exp10];
      return 0;
    }
  }
  uint64_t b0;
  if (!eisel_lemire(mant, exp10, b0)) return 1;
  if (truncated) {
    uint64_t b1;
    if (!eisel_lemire(mant + 1, exp10, b1) || b1 != b0) return 1;
  }
  out = bits_to_double(b0);
  return 0;
}

// strtod over a byte cursor (C::at(int64_t) -> unsigned char): [sign] digits [. digits] [e [sign] digits].
// q is advanced past the literal.  Returns 0, 4 (no number: BK_ERR_PARSE) or 6 (nan/inf/hex or a literal the exact
// stages above refuse: BK_ERR_UNSUPPORTED).
template <class C>
BK_FN int parse_decimal(const C& c, int64_t& q, double& out) {
  unsigned char ch = c.at(q);
  bool          neg = false;
  if (ch == '+' || ch == '-') {
    neg = ch == '-';
    ch = c.at(++q);
  }
  uint64_t mant = 0;
  int      nd = 0, exp10 = 0;
  bool     any = false, truncated = false;
  if (ch == '0' && (c.at(q + 1) == 'x' || c.at(q + 1) == 'X')) return 6;  // hex float
  while (ch >= '0' && ch <= '9') {
    any = true;
    if (nd < 19) {
      mant = mant * 10 + (ch - '0');
      if (mant) nd++;
    } else {
      exp10++;
      truncated |= ch != '0';
    }
    ch = c.at(++q);
  }
  if (ch == '.') {
    ch = c.at(++q);
    while (ch >= '0' && ch <= '9') {
      any = true;
      if (nd < 19) {
        mant = mant * 10 + (ch - '0');
        if (mant) nd++;
        exp10--;
      } else {
        truncated |= ch != '0';
      }
      ch = c.at(++q);
    }
  }
  if (!any) return (ch == 'n' || ch == 'N' || ch == 'i' || ch == 'I') ? 6 : 4;
  if (ch == 'e' || ch == 'E') {
    int64_t       q2 = q + 1;
    unsigned char c2 = c.at(q2);
    bool          eneg = false;
    if (c2 == '+' || c2 == '-') {
      eneg = c2 == '-';
      c2 = c.at(++q2);
    }
    if (c2 >= '0' && c2 <= '9') {
      int e = 0;
      while (c2 >= '0' && c2 <= '9') {
        if (e < 100000) e = e * 10 + (c2 - '0');
        c2 = c.at(++q2);
      }
      exp10 += eneg ? -e : e;
      q = q2;
    }
  }
  double v;
  if (decimal_to_double(mant, exp10, truncated, v)) return 6;
  out = neg ? -v : v;
  return 0;
}

}  // namespace bk
