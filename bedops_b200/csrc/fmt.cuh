// fmt.cuh -- device-side ASCII emission primitives (every field the reference prints with printf).
//
// Sinks: a row is formatted twice by the same code, first into a CountSink (byte length for the look-back scan),
// then into a MemSink (shared-memory staging or, for oversized tiles, global memory).
#pragma once
#include "common.cuh"

namespace bk {
#ifdef __CUDACC__

__device__ __forceinline__ int ndigits_u64(uint64_t v) {
  int n = 1;
  if (v >= 10000000000000000ull) { v /= 10000000000000000ull; n += 16; }
  if (v >= 100000000ull) { v /= 100000000ull; n += 8; }
  if (v >= 10000ull) { v /= 10000ull; n += 4; }
  if (v >= 100ull) { v /= 100ull; n += 2; }
  if (v >= 10ull) n += 1;
  return n;
}
__device__ __forceinline__ int ndigits_u32(uint32_t v) {
  int n = 1;
  if (v >= 100000000u) { v /= 100000000u; n += 8; }
  if (v >= 10000u) { v /= 10000u; n += 4; }
  if (v >= 100u) { v /= 100u; n += 2; }
  if (v >= 10u) n += 1;
  return n;
}

struct CountSink {
  uint64_t n = 0;
  __device__ __forceinline__ void put(char) { n++; }
  __device__ __forceinline__ void put_u64(uint64_t v) { n += ndigits_u64(v); }
  __device__ __forceinline__ void put_u32(uint32_t v) { n += ndigits_u32(v); }
  __device__ __forceinline__ void put_padded(uint64_t, int width) { n += width; }
  __device__ __forceinline__ void copy(const char*, uint64_t len) { n += len; }
  __device__ __forceinline__ void puts_(const char* s, int len) { n += len; }
  static constexpr bool counting = true;
};

struct MemSink {  // generic-address writer (shared or global)
  char* p;
  __device__ __forceinline__ void put(char c) { *p++ = c; }
  __device__ __forceinline__ void put_u64(uint64_t v) {
    int n = ndigits_u64(v);
    char* e = p + n;
    do {
      *--e = (char)('0' + (v % 10));
      v /= 10;
    } while (v);
    p += n;
  }
  __device__ __forceinline__ void put_u32(uint32_t v) {
    int n = ndigits_u32(v);
    char* e = p + n;
    do {
      *--e = (char)('0' + (v % 10));
      v /= 10;
    } while (v);
    p += n;
  }
  __device__ __forceinline__ void put_padded(uint64_t v, int width) {  // zero-padded to width digits
    char* e = p + width;
    for (int i = 0; i < width; i++) {
      *--e = (char)('0' + (v % 10));
      v /= 10;
    }
    p += width;
  }
  __device__ __forceinline__ void copy(const char* src, uint64_t len) {
    for (uint64_t i = 0; i < len; i++) p[i] = src[i];
    p += len;
  }
  __device__ __forceinline__ void puts_(const char* s, int len) {
    for (int i = 0; i < len; i++) p[i] = s[i];
    p += len;
  }
  static constexpr bool counting = false;
};

__constant__ uint64_t kPow10u[20] = {1ull,
                                     10ull,
                                     100ull,
                                     1000ull,
                                     10000ull,
                                     100000ull,
                                     1000000ull,
                                     10000000ull,
                                     100000000ull,
                                     1000000000ull,
                                     10000000000ull,
                                     100000000000ull,
                                     1000000000000ull,
                                     10000000000000ull,
                                     100000000000000ull,
                                     1000000000000000ull,
                                     10000000000000000ull,
                                     100000000000000000ull,
                                     1000000000000000000ull,
                                     10000000000000000000ull};

// Exact "%.<prec>f" of a double: the binary value is expanded exactly and rounded half-to-even on the exact
// remainder, which is what glibc printf does in the default rounding mode (SURVEY hard part 4).
// Supported: prec <= 18 and |x| < 2^63; returns false otherwise (caller raises BK_ERR_UNSUPPORTED).
struct Fixed {
  bool     neg;
  uint64_t ip;    // integer part
  uint64_t frac;  // fraction scaled by 10^prec, < 10^prec
  int      special;  // 0 finite, 1 nan, 2 inf
};
__device__ __forceinline__ bool to_fixed(double x, int prec, Fixed& f) {
  const uint64_t bits = (uint64_t)__double_as_longlong(x);
  f.neg = bits >> 63;
  f.special = 0;
  const int      e = (int)((bits >> 52) & 0x7FF);
  uint64_t       m = bits & ((1ull << 52) - 1);
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
  const uint64_t pw = kPow10u[prec];
  const uint64_t lo = fm * pw, hi = __umul64hi(fm, pw);
  uint64_t       q;
  int            cmp;  // remainder vs half: -1 below, 0 tie, +1 above
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
  if (cmp > 0 || (cmp == 0 && (q & 1))) q++;
  if (q >= pw) {
    q -= pw;
    f.ip++;
  }
  f.frac = q;
  return true;
}

template <class Sink>
__device__ __forceinline__ void put_fixed(Sink& s, const Fixed& f, int prec) {
  if (f.special) {
    if (f.neg) s.put('-');
    if (f.special == 1) s.puts_("nan", 3); else s.puts_("inf", 3);
    return;
  }
  if (f.neg) s.put('-');
  s.put_u64(f.ip);
  if (prec > 0) {
    s.put('.');
    s.put_padded(f.frac, prec);
  }
}

#endif
}  // namespace bk
