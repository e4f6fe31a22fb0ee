// fmt.cuh -- device-side ASCII emission primitives (every field the reference prints with printf).
//
// Sinks: a row is formatted twice by the same code, first into a CountSink (byte length for the look-back scan),
// then into a MemSink (shared-memory staging or, for oversized tiles, global memory).
#pragma once
#include "common.cuh"
#include "fixed_exact.cuh"

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
