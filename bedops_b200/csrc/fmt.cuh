// fmt.cuh -- device-side ASCII emission primitives (every field the reference prints with printf).
//
// Sinks: a row is formatted twice by the same code, first into a CountSink (byte length, k_emit_len),
// then into a MemSink (shared-memory staging or, for oversized tiles, global memory).
#pragma once
#include "common.cuh"
#include "fixed_exact.cuh"

namespace bk {
#ifdef __CUDACC__

__device__ __forceinline__ int ndigits_u32(uint32_t v) {  // compare tree: no divisions
  return v < 100000u ? (v < 100u ? (v < 10u ? 1 : 2) : (v < 1000u ? 3 : (v < 10000u ? 4 : 5)))
                     : (v < 10000000u ? (v < 1000000u ? 6 : 7) : (v < 100000000u ? 8 : (v < 1000000000u ? 9 : 10)));
}
__device__ __forceinline__ int ndigits_u64(uint64_t v) {
  if ((v >> 32) == 0) return ndigits_u32((uint32_t)v);
  int n = 1;
  if (v >= 10000000000000000ull) { v /= 10000000000000000ull; n += 16; }
  if (v >= 100000000ull) { v /= 100000000ull; n += 8; }
  if (v >= 10000ull) { v /= 10000ull; n += 4; }
  if (v >= 100ull) { v /= 100ull; n += 2; }
  if (v >= 10ull) n += 1;
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
  __device__ __forceinline__ uint64_t gpos() const { return n; }
  __device__ __forceinline__ void     skip(uint64_t len) { n += len; }
  static constexpr bool counting = true;
};

struct MemSink {  // generic-address writer (shared or global)
  char*    p;
  char*    p0 = nullptr;  // where the row began, and the byte offset of that place in the result text: a column can
  uint64_t g0 = 0;        // ask where it lands (gpos) and leave its bytes to a later cooperative pass (skip)
  __device__ __forceinline__ uint64_t gpos() const { return g0 + (uint64_t)(p - p0); }
  __device__ __forceinline__ void     skip(uint64_t n) { p += n; }
  __device__ __forceinline__ void put(char c) { *p++ = c; }
  __device__ __forceinline__ void put_u32(uint32_t v) {
    const int n = ndigits_u32(v);
    char*     e = p + n;
    do {
      const uint32_t q = v / 10u;
      *--e = (char)('0' + (v - q * 10u));
      v = q;
    } while (v);
    p += n;
  }
  __device__ __forceinline__ void put_u64(uint64_t v) {
    if ((v >> 32) == 0) {  // 32-bit arithmetic whenever the value allows (64-bit division by 10 is ~5x the work)
      put_u32((uint32_t)v);
      return;
    }
    int n = ndigits_u64(v);
    char* e = p + n;
    do {
      *--e = (char)('0' + (v % 10));
      v /= 10;
    } while (v);
    p += n;
  }
  __device__ __forceinline__ void put_padded(uint64_t v, int width) {  // zero-padded to width digits
    char* e = p + width;
    if ((v >> 32) == 0) {
      uint32_t w = (uint32_t)v;
      for (int i = 0; i < width; i++) {
        const uint32_t q = w / 10u;
        *--e = (char)('0' + (w - q * 10u));
        w = q;
      }
    } else {
      for (int i = 0; i < width; i++) {
        *--e = (char)('0' + (v % 10));
        v /= 10;
      }
    }
    p += width;
  }
  // Byte copy global -> here.  Word-wise once the destination is 4-byte aligned: the source words are read aligned
  // (two neighbours funnel-shifted), four independent loads in flight.  The aligned reads may touch up to 3 bytes in
  // front of src -- inside the text buffer, whose base is 16-byte aligned (the reader requires it) -- and never read
  // past src+len rounded up to the word that holds the last byte.
  __device__ __forceinline__ void copy(const char* __restrict__ src, uint64_t len) {
    uint64_t i = 0;
    while (i < len && (reinterpret_cast<uintptr_t>(p + i) & 3)) {
      p[i] = src[i];
      i++;
    }
    if (len - i >= 8) {
      const unsigned  k = (unsigned)(reinterpret_cast<uintptr_t>(src + i) & 3);
      const uint32_t* a = reinterpret_cast<const uint32_t*>(src + i - k);
      uint32_t*       d = reinterpret_cast<uint32_t*>(p + i);
      const uint64_t  nw = (len - i - 4) >> 2;  // words whose right neighbour still starts inside [src, src+len)
      uint32_t        cur = __ldg(a);
      uint64_t        j = 0;
      for (; j + 4 <= nw; j += 4) {
        const uint32_t w1 = __ldg(a + j + 1), w2 = __ldg(a + j + 2), w3 = __ldg(a + j + 3), w4 = __ldg(a + j + 4);
        d[j] = __funnelshift_r(cur, w1, 8 * k);
        d[j + 1] = __funnelshift_r(w1, w2, 8 * k);
        d[j + 2] = __funnelshift_r(w2, w3, 8 * k);
        d[j + 3] = __funnelshift_r(w3, w4, 8 * k);
        cur = w4;
      }
      for (; j < nw; j++) {
        const uint32_t w1 = __ldg(a + j + 1);
        d[j] = __funnelshift_r(cur, w1, 8 * k);
        cur = w1;
      }
      i += nw * 4;
    }
    for (; i < len; i++) p[i] = src[i];
    p += len;
  }
  __device__ __forceinline__ void puts_(const char* s, int len) {
    if (len == 1) {
      *p++ = s[0];
      return;
    }
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
