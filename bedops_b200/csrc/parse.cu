// parse.cu -- the BED reader as two sm_100a passes over the text.
//
// Replaces Bed::allocate_iterator_starch_bed<T*>::operator++ -> T::readline(FILE*) with fscanf formats
// "%s\t%lu\t%lu%[^\n]s\n" (B3Rest, Bed.hpp:380-382), "...\t%s%[^\n]s\n" (B4Rest, :644-646) and
// "...\t%s\t%lf%[^\n]s\n" (B5Rest, :901-903).
//
// Design (B200): pass 1 (k_count_rows) streams the text once and counts the rows that start in every 8 KiB tile;
// a one-CTA scan turns the counts into the first row of every tile, so the tiles of pass 2 are independent (no
// look-back chain) and the columns are allocated exactly.  Pass 2 (k_parse) stages a tile plus a small halo in
// shared memory with 16-byte cp.async copies issued one tile ahead (double buffer), finds field boundaries with
// SWAR byte masks, and one thread per line tokenises its line out of shared memory and writes the SoA columns
// with coalesced 4/8-byte stores.  Nothing but the SoA columns is written.
#include <algorithm>
#include "common.cuh"
#include "parse.cuh"
#include "strtod_exact.cuh"

namespace bk {

// ---- byte cursor: shared-memory window with a global-memory fallback for lines that leave the window --------
struct Cursor {
  const unsigned char* sm;    // window base
  int64_t              g0;    // global byte offset of sm[0] (may be negative for tile 0)
  const unsigned char* text;  // global text
  uint64_t             nbytes;
  __device__ __forceinline__ unsigned char at(int64_t q) const {  // q = index into the window
    if (q >= 0 && q < P_BUF) return sm[q];
    int64_t g = g0 + q;
    if (g < 0 || (uint64_t)g >= nbytes) return '\n';
    return text[g];
  }
};
struct SmCursor {  // fast path: the whole line is known to lie inside the window
  const unsigned char* sm;
  __device__ __forceinline__ unsigned char at(int64_t q) const { return sm[q]; }
};

struct RowOut {
  uint64_t start, end;
  double   score;
  int64_t  tok0;  // window index of the chromosome token
  int      toklen;
  int64_t  id0;
  int      idlen;
  int      err;  // 0 ok, else BK_ERR_*
};

// strtoul-like: optional '+', decimal digits.  Returns false when there is no digit.
template <class C>
__device__ __forceinline__ bool parse_u64(const C& c, int64_t& q, uint64_t& v) {
  unsigned char ch = c.at(q);
  if (ch == '+') ch = c.at(++q);
  if (ch < '0' || ch > '9') return false;
  uint64_t acc = 0;
  bool     sat = false;
  do {
    if (acc > 1844674407370955160ull) sat = true;
    acc = acc * 10 + (ch - '0');
    ch = c.at(++q);
  } while (ch >= '0' && ch <= '9');
  v = sat ? ~0ull : acc;
  return true;
}

// tokenise one line starting at window index q0
template <class C>
__device__ __forceinline__ void parse_line(const C& c, int64_t q0, int min_fields, unsigned cols, RowOut& r) {
  r.err = 0;
  r.score = 0.0;
  r.id0 = 0;
  r.idlen = 0;
  int64_t q = q0;
  while (is_ws(c.at(q))) q++;
  r.tok0 = q;
  while (is_tok(c.at(q))) q++;
  r.toklen = (int)(q - r.tok0);
  if (r.toklen > 127) { r.err = BK_ERR_PARSE; return; }  // MAXCHROMSIZE, BEDOPS.Constants.hpp:32
  while (is_ws(c.at(q))) q++;
  if (!parse_u64(c, q, r.start)) { r.err = BK_ERR_PARSE; return; }
  while (is_ws(c.at(q))) q++;
  if (!parse_u64(c, q, r.end)) { r.err = BK_ERR_PARSE; return; }
  if (r.start >= 0xFFFFFFFFull || r.end >= 0xFFFFFFFFull) { r.err = BK_ERR_COORD_RANGE; return; }
  if (min_fields >= 4) {
    while (is_ws(c.at(q))) q++;
    r.id0 = q;
    while (is_tok(c.at(q))) q++;
    r.idlen = (int)(q - r.id0);
    if (r.idlen == 0 || r.idlen > 16383 || (r.id0 - r.tok0) > 65535) { r.err = BK_ERR_PARSE; return; }
    if (min_fields >= 5) {
      while (is_ws(c.at(q))) q++;
      if (c.at(q) == '\n') { r.err = BK_ERR_PARSE; return; }
      if (cols & BK_COL_SCORE) {
        int e = parse_decimal(c, q, r.score);
        if (e) { r.err = e; return; }
      }
    }
  }
}

__device__ __noinline__ void parse_line_slow(const Cursor& c, int64_t q0, int min_fields, unsigned cols, RowOut& r) {
  parse_line(c, q0, min_fields, cols, r);
}
__device__ __noinline__ int parse_score_slow(const unsigned char* sm, int q, double& out) {
  SmCursor sc{sm};
  int64_t  qq = q;
  return parse_decimal(sc, qq, out);
}

// compare the chromosome token at window index a (length la) with the token starting at window index b
template <class C>
__device__ __forceinline__ bool same_token(const C& c, int64_t a, int la, int64_t b) {
  for (int i = 0; i < la; i++)
    if (c.at(a + i) != c.at(b + i)) return false;
  return !is_tok(c.at(b + la));
}

// window index of the first token of the last non-blank line that ends before window index p0 (p0 = a line start);
// INT64_MIN if there is none.  Rare path (first row of a tile only): walks backwards through the window/global text.
__device__ __noinline__ int64_t prev_line_token(const Cursor& c, int64_t p0) {
  int64_t q = p0 - 1;  // the '\n' that terminates the previous line
  while (true) {
    if (c.g0 + q < 0) return INT64_MIN;
    int64_t s = q;
    while (c.g0 + s > 0 && c.at(s - 1) != '\n') s--;
    int64_t t = s;
    while (t < q && is_ws(c.at(t))) t++;
    if (t < q) return t;
    if (c.g0 + s <= 0) return INT64_MIN;
    q = s - 1;
  }
}

// ---- fast path helpers: control-byte bitmask + SWAR digit conversion -------------------------------------------
// ctlp: one bit per text byte, set iff the byte is < 0x21 (TAB, NL, CR, space, ...).  Canonical BED has exactly one
// such byte after every field, so the set bits of the 64-bit window that starts at a line start ARE the field
// boundaries; reading the byte found tells TAB from NL from anything that sends the line to the general tokeniser.
constexpr int P_NW = (P_TILE + P_POST) / 32;  // mask words; word w covers window bytes [P_PRE+32w, P_PRE+32w+32)

__device__ __forceinline__ uint32_t ctl_mask4(uint32_t w) {  // bit 8j+7 set iff byte j < 0x21
  return ~(((w & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) | w) & 0x80808080u;
}
__device__ __forceinline__ uint32_t nl_mask4(uint32_t w, uint32_t ctl) {  // bytes == '\n' (given the ctl mask)
  return ~(((w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) & ctl;
}
// Byte flags (bit 7 of each byte) of eight consecutive words -> one bit per byte, bit 4i+j = byte j of word i.
// One multiply gathers the four flags of a word in its top nibble (7->28, 15->29, 23->30, 31->31: the partial
// products land on distinct bits, no carries); a funnel shift appends that nibble to the accumulator.  Words are
// fed last to first so that word 0 ends in the lowest nibble.  Two instructions per word (IMAD + SHF).
__device__ __forceinline__ uint32_t push_flags4(uint32_t acc, uint32_t m) {
  return __funnelshift_l(m * 0x00204081u, acc, 4);  // (acc << 4) | (top nibble of the product)
}

// unaligned 32-bit read of window bytes [x, x+4)
__device__ __forceinline__ uint32_t ld32u(const unsigned char* sm, int x) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(sm) + (x >> 2);
  return __funnelshift_r(w[0], w[1], (x & 3) * 8);
}

// four ASCII digits (first digit in the lowest byte) -> 0..9999; bad accumulates a non-zero value on a non-digit.
// Three multiplies and two masks: the kernel is bound by the ALU pipe (logic/shift/compare), multiplies run on the
// FMA pipe.  t = digit values d0..d3 in bytes 0..3;  t*0x0A01 puts 10*d0+d1 in byte 1 and 10*d2+d3 in byte 3;
// the high half of (a<<8 | b<<24) * (100<<24 | 1<<8) is 100*a + b (+ a multiple of 2^16).
__device__ __forceinline__ uint32_t digits4(uint32_t w, uint32_t& bad) {
  const uint32_t t = w - 0x30303030u;
  bad |= ((w + 0x46464646u) | t) & 0x80808080u;
  const uint32_t y = (t * 0x0A01u) & 0xFF00FF00u;
  return __umulhi(y, 0x64000100u) & 0xFFFFu;
}

// decimal field of len digits (1..9) that ENDS at window byte index e (exclusive) -> value
__device__ __forceinline__ uint32_t parse_digits_swar(const unsigned char* sm, int e, int len, uint32_t& bad) {
  uint32_t g0 = ld32u(sm, e - 4), g1 = ld32u(sm, e - 8);
  // bytes in front of the field (the lower bytes) are replaced by '0':  keep the top min(len,4) bytes of g0 and
  // the top clamp(len-4,0,4) bytes of g1
  const uint64_t mm = len >= 8 ? 0ull : (~0ull >> (8 * len));  // one 64-bit shift instead of two compare/select chains
  const uint32_t m1 = (uint32_t)mm, m0 = (uint32_t)(mm >> 32);
  g0 = (g0 & ~m0) | (0x30303030u & m0);
  g1 = (g1 & ~m1) | (0x30303030u & m1);
  uint32_t v = digits4(g1, bad) * 10000u + digits4(g0, bad);
  if (len == 9) {
    const uint32_t d = (uint32_t)sm[e - 9] - '0';
    bad |= d > 9 ? 1u : 0u;
    v += d * 100000000u;
  }
  return v;
}

// first 8 bytes of a token of length len (>= 1) at window index x, bytes beyond the token zeroed
__device__ __forceinline__ uint2 token8(const unsigned char* sm, int x, int len) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(sm) + (x >> 2);
  const int       sh = (x & 3) * 8;
  uint32_t        lo = __funnelshift_r(w[0], w[1], sh), hi = __funnelshift_r(w[1], w[2], sh);
  if (len < 4) {
    lo &= (1u << (8 * len)) - 1u;
    hi = 0;
  } else if (len < 8) {
    hi &= (1u << (8 * (len - 4))) - 1u;  // len == 4 -> 0
  }
  return make_uint2(lo, hi);
}

// --ec/--header: is the line that starts at q a header (UCSC "browser"/"track" keyword, or first byte '@' / '#')?
// (BedCheckIterator.hpp:315-360).  C::at(q) returns the byte at q.
template <class C>
__device__ __forceinline__ bool is_header_line(const C& c, int64_t q) {
  unsigned char ch = c.at(q);
  if (ch == '@' || ch == '#') return true;
  if ((ch | 32) != 'b' && (ch | 32) != 't') return false;
  const char* w = (ch | 32) == 'b' ? "browser" : "track";
  const int   wl = (ch | 32) == 'b' ? 7 : 5;
  for (int i = 0; i < wl; i++) {
    unsigned char x = c.at(q + i);
    if (x >= 'A' && x <= 'Z') x += 32;
    if (x != (unsigned char)w[i]) return false;
  }
  const unsigned char e = c.at(q + wl);
  return e == ' ' || e == '\t' || e == '\n';
}
struct TextCursor {  // global text with the parser's end-of-text convention
  const unsigned char* t;
  uint64_t             eff;
  __device__ __forceinline__ unsigned char at(int64_t q) const { return (q < 0 || (uint64_t)q >= eff) ? '\n' : t[q]; }
};

// packed control-byte mask and NL mask (bit i = byte i) of 32 text bytes held in 8 words
__device__ __forceinline__ void pack_masks32(const uint32_t (&w8)[8], uint32_t& cm, uint32_t& nlp) {
  cm = 0;
  nlp = 0;
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    const uint32_t ctl = ctl_mask4(w8[i]);
    cm = push_flags4(cm, ctl);
    nlp = push_flags4(nlp, nl_mask4(w8[i], ctl));
  }
}

// ---- pass 1: rows per tile ---------------------------------------------------------------------------------------
// A row starts at p iff (p == 0 or byte p-1 is NL), p < eff, and the line is not blank.  Same definition, same helper
// functions as the parser below, so the two passes always agree on the row numbering.  Streams the text once
// (pure bandwidth); the exclusive scan of the counts gives every tile its first row, which removes any ordering
// between tiles from the parser (no look-back chain to wait for) and makes the row count -- hence the column
// allocation -- exact.
// One WARP owns a contiguous range of `tiles_per_warp` tiles and walks it 1 KiB at a time (32 lanes x 32 bytes,
// two 16-byte streaming loads per lane), so there is no block barrier and no atomics: local_prefix[tile] = rows of the
// warp's earlier tiles, warp_total[warp] = rows of its whole range.  The NL that precedes a lane's span comes from the
// neighbouring lane by shuffle (and from the previous step's lane 31).
constexpr int CR_THREADS = 256;
__global__ void __launch_bounds__(CR_THREADS) k_count_rows(const unsigned char* __restrict__ text, uint64_t nbytes_raw,
                                                           const uint64_t* __restrict__ scratch,
                                                           uint32_t* __restrict__ local_prefix, uint64_t* __restrict__ warp_total,
                                                           uint32_t ntiles, uint32_t tiles_per_warp, uint32_t nwarps,
                                                           int skip_headers, int strict_blank, uint32_t* __restrict__ cm_arr,
                                                           uint32_t* __restrict__ ls_arr) {
  const int      lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * CR_THREADS + threadIdx.x) >> 5;
  if (wid >= nwarps) return;
  const uint64_t eff = scratch[SC_EFFLEN];
  const uint32_t t0 = wid * tiles_per_warp, t1 = (t0 + tiles_per_warp < ntiles) ? t0 + tiles_per_warp : ntiles;
  uint64_t       run = 0;
  uint32_t       carry = 0;  // was the byte just before this step's first byte a NL?
  if (t0 < t1) {
    const uint64_t g = (uint64_t)t0 * P_TILE;
    carry = (g == 0 || (g - 1 < nbytes_raw && text[g - 1] == '\n')) ? 1u : 0u;
  }
  for (uint32_t tile = t0; tile < t1; tile++) {
    if (lane == 0) local_prefix[tile] = (uint32_t)run;  // a warp range holds far fewer than 2^32 rows
    uint32_t cnt = 0;
#pragma unroll 2
    for (int step = 0; step < P_TILE / 1024; step++) {
      const uint64_t p0 = (uint64_t)tile * P_TILE + (uint64_t)step * 1024 + (uint64_t)lane * 32;
      uint32_t       w8[8];
      if (p0 + 32 <= nbytes_raw) {
        const uint4 a = ldg_stream16(text + p0), c = ldg_stream16(text + p0 + 16);
        w8[0] = a.x; w8[1] = a.y; w8[2] = a.z; w8[3] = a.w; w8[4] = c.x; w8[5] = c.y; w8[6] = c.z; w8[7] = c.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          uint32_t w = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint64_t g = p0 + 4 * i + j;
            w |= (uint32_t)(g < nbytes_raw ? text[g] : 0) << (8 * j);
          }
          w8[i] = w;
        }
      }
      uint32_t cm, nlp;
      pack_masks32(w8, cm, nlp);
      uint32_t prev = __shfl_up_sync(0xffffffffu, nlp >> 31, 1);
      if (lane == 0) prev = carry;
      carry = __shfl_sync(0xffffffffu, nlp >> 31, 31);
      uint32_t smask = (nlp << 1) | prev;
      if (p0 >= eff) smask = 0;
      else if (p0 + 32 > eff) smask &= (1u << (int)(eff - p0)) - 1u;
      for (uint32_t m = smask & cm; m; m &= m - 1) {  // a line that begins with a control byte may be blank
        const int j = __ffs(m) - 1;
        uint64_t  q = p0 + j;
        if (!strict_blank)  // sort-bed skips only empty lines (SortDetails.cpp:625-629); fscanf skips any whitespace
          while (q < eff && is_ws(text[q])) q++;
        if (q >= eff || text[q] == '\n') smask &= ~(1u << j);
      }
      if (skip_headers) {
        const TextCursor tc{text, eff};
        for (uint32_t m = smask; m; m &= m - 1) {
          const int j = __ffs(m) - 1;
          if (is_header_line(tc, (int64_t)(p0 + j))) smask &= ~(1u << j);
        }
      }
      cnt += __popc(smask);
      // the packed control-byte mask and the final line-start mask of these 32 bytes, for pass 2 (one bit per text byte
      // each: 1/4 of the text size written here and read there, instead of building both masks twice)
      if (p0 < nbytes_raw) {
        __stcs(&cm_arr[p0 >> 5], cm);  // streaming stores: written once here, read once by pass 2
        __stcs(&ls_arr[p0 >> 5], smask);
      }
    }
    run += __reduce_add_sync(0xffffffffu, cnt);
  }
  if (lane == 0) warp_total[wid] = run;
}

// first row of every tile = base of its warp range + rows of the earlier tiles of the range
__global__ void k_tile_base(const uint32_t* __restrict__ local_prefix, const uint64_t* __restrict__ warp_base,
                            uint32_t tiles_per_warp, uint32_t ntiles, uint64_t* __restrict__ tile_base) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntiles) tile_base[t] = warp_base[t / tiles_per_warp] + local_prefix[t];
}

// decimal field of len digits (1..4) that ENDS at window byte index e (exclusive) -> value (scores of a few digits)
__device__ __forceinline__ uint32_t parse_digits_short(const unsigned char* sm, int e, int len, uint32_t& bad) {
  uint32_t       g0 = ld32u(sm, e - 4);
  const uint32_t m0 = len >= 4 ? 0u : (0xFFFFFFFFu >> (8 * len));  // the bytes in front of the field become '0'
  g0 = (g0 & ~m0) | (0x30303030u & m0);
  return digits4(g0, bad);
}

// ---- one line of a staged tile -----------------------------------------------------------------------------------
// What a tile offers the thread that parses line i: the staged text window, the control-byte mask of [PRE, PRE+TILE+POST),
// the line-start mask of [0, PRE+TILE) (rows only), the window index of line i (q0) and of line i-1 (prev_q0, -1 for the
// tile's first line), the tile's first row.
struct TileView {
  const unsigned char* sm;
  const uint32_t*      ctlp;
  const uint32_t*      lsw;
  int64_t              g0;
  uint64_t             base;
};
template <int NSEP, bool WANT_SCORE>
__device__ __forceinline__ void parse_tile_line(const ParseParams& p, const TileView& tv, const Cursor& cur, uint32_t i, int q0, int prev_q0) {
  const unsigned char* const sm = tv.sm;
  const uint32_t* const      ctlp = tv.ctlp;
  const uint32_t* const      lsw = tv.lsw;
  const int64_t              g0 = tv.g0;
  const uint64_t             base = tv.base;
  uint32_t  v_start = 0, v_end = 0, v_id = 0;
  double    v_score = 0.0;
  int       tok0 = q0, toklen = 0, err = 0;
  uint32_t  linelen = 0xFFFFu;  // bytes up to the NL when the line is canonical (echo = verbatim copy), else 0xFFFF
  bool      head = false;
  // 64 line bytes of control-byte mask, starting at the line start
  const int      b0 = q0 - P_PRE, w0 = b0 >> 5, sh = b0 & 31;
  const uint32_t c0 = ctlp[w0], c1 = ctlp[w0 + 1], c2 = ctlp[w0 + 2];
  uint32_t Wlo = __funnelshift_r(c0, c1, sh), Whi = __funnelshift_r(c1, c2, sh);  // bytes 0..31 / 32..63 of the line
  int  sp[NSEP];  // offsets (from q0) of the control bytes that end the first NSEP fields
  bool fast = true;
#pragma unroll
  for (int f = 0; f < NSEP; f++) {
    if (Wlo) {
      sp[f] = __ffs(Wlo) - 1;
      Wlo &= Wlo - 1;
    } else {
      fast = fast && Whi != 0;
      sp[f] = 31 + __ffs(Whi);
      Whi &= Whi - 1;
    }
  }
  uint32_t bad = 0;
  if (fast) {
    // every separator but the last must be a TAB, the last a TAB or the NL
#pragma unroll
    for (int f = 0; f < NSEP; f++) {
      const unsigned char c = sm[q0 + sp[f]];
      bad |= (c == '\t' || (f + 1 == NSEP && c == '\n')) ? 0u : 1u;
    }
    toklen = sp[0];
    const int l1 = sp[1] - sp[0] - 1, l2 = sp[2] - sp[1] - 1;
    bad |= (toklen < 1 || toklen > 127 || l1 < 1 || l1 > 9 || l2 < 1 || l2 > 9) ? 1u : 0u;
    if (!bad) {
      v_start = parse_digits_swar(sm, q0 + sp[1], l1, bad);
      v_end = parse_digits_swar(sm, q0 + sp[2], l2, bad);
    }
    if (NSEP >= 4) {
      const int idlen = sp[NSEP >= 4 ? 3 : 0] - sp[2] - 1;
      bad |= (idlen < 1 || idlen > 16383) ? 1u : 0u;
      v_id = ((uint32_t)(sp[2] + 1) << 16) | (uint32_t)idlen;
    }
    if (NSEP >= 5) {
      const int s3 = sp[NSEP >= 5 ? 3 : 0], s4 = sp[NSEP >= 5 ? 4 : 0];
      const int l4 = s4 - s3 - 1;
      bad |= l4 < 1 ? 1u : 0u;
      if (WANT_SCORE && !bad) {
        uint32_t sbad = l4 > 9 ? 1u : 0u;
        if (l4 <= 4) v_score = (double)parse_digits_short(sm, q0 + s4, l4, sbad);
        else if (!sbad) v_score = (double)parse_digits_swar(sm, q0 + s4, l4, sbad);
        if (sbad) bad |= parse_score_slow(sm, q0 + s3 + 1, v_score) ? 1u : 0u;  // exact strtod on the field
      }
    }
    fast = bad == 0;
    if (fast && p.line_off) {
      // canonical for echo: single TABs (checked above), no leading zeros -> re-printing the numbers reproduces the
      // input bytes, so the whole line can be copied.  Find the NL among the remaining control bytes.
      int nlpos = -1;
      if (sm[q0 + sp[NSEP - 1]] == '\n') nlpos = sp[NSEP - 1];
      else {
        unsigned long long W = ((unsigned long long)Whi << 32) | Wlo;  // the remaining control bytes
        for (int it = 0; nlpos < 0 && W != 0 && it < 8; it++) {
          const int r = __ffsll((long long)W) - 1;
          if (sm[q0 + r] == '\n') nlpos = r;
          W &= W - 1;
        }
      }
      const bool lz = (sm[q0 + sp[0] + 1] == '0' && l1 > 1) || (sm[q0 + sp[1] + 1] == '0' && l2 > 1);
      if (nlpos >= 0 && !lz) linelen = (uint32_t)nlpos;
    }
  }
  if (!fast) {  // general path: fscanf-equivalent tokeniser (out of line: keeps the fast path's registers low)
    RowOut r;
    parse_line_slow(cur, q0, p.min_fields, p.cols, r);
    err = r.err;
    v_start = (uint32_t)r.start;
    v_end = (uint32_t)r.end;
    v_score = r.score;
    v_id = ((uint32_t)(r.id0 - r.tok0) << 16) | (uint32_t)r.idlen;
    tok0 = (int)r.tok0;
    toklen = r.toklen;
  }
  const uint64_t row = base + i;
  if (p.cols & BK_LOAD_SORTBED) {
    // sort-bed reading: any chromosome order (no run heads), and a row this tokeniser does not take is left to the
    // sorter's own validation (k_sort_validate re-reads every non-canonical row by sort-bed's grammar)
    if (row < p.cap) {
      const bool plain = fast && err == 0;
      p.start[row] = plain ? v_start : 0u;
      p.end[row] = plain ? v_end : 0u;
      p.line_off[row] = ((uint64_t)(plain ? linelen : 0xFFFFu) << 48) | (uint64_t)(g0 + q0);
      if (p.idspan) p.idspan[row] = plain ? (uint32_t)sp[2] : 0u;  // offset of the separator after the end coordinate
    }
    return;
  }
  if (err) {
    dev_set_error(p.scratch, err, row);
    return;
  }
  // chromosome run head?  Compare the token with the previous row's: row i-1 of this tile starts at lstart[i-1];
  // the row before the tile's first one starts after the last NL but one before q0 (NL mask of the head halo).
  {
    bool done = false;
    if (toklen <= 8 && tok0 == q0) {
      int ps = -1;
      if (i > 0) ps = prev_q0;
      else if (q0 >= 1) {  // the last line start in front of q0 (rows only: blank and header lines are not in the mask)
        const int x = q0 - 1;
        int       w = x >> 5;
        uint32_t  pm = lsw[w] & (0xFFFFFFFFu >> (31 - (x & 31)));
        while (!pm && w > 0) pm = lsw[--w];
        if (pm) ps = 32 * w + 31 - __clz(pm);
      }
      if (ps >= 0 && ps + toklen < q0 && sm[ps] > 0x20) {
        const uint2 t8 = token8(sm, q0, toklen), c = token8(sm, ps, toklen);
        head = t8.x != c.x || t8.y != c.y || sm[ps + toklen] > 0x20;
        done = true;
      }
    }
    if (!done) {
      int64_t pt = prev_line_token(cur, q0);
      head = (pt == INT64_MIN) || !same_token(cur, tok0, toklen, pt);
    }
  }
  if (head) {
    uint32_t h = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(&p.scratch[SC_NHEADS]), 1ull);
    if (h < p.heads_cap) {
      HeadRec* hr = &p.heads[h];
      hr->row = row;
      hr->len = toklen;
      for (int k = 0; k < toklen; k++) hr->name[k] = cur.at(tok0 + k);
      hr->name[toklen] = 0;
    }
  }
  if (row < p.cap) {
    p.start[row] = v_start;
    p.end[row] = v_end;
    if (WANT_SCORE) p.score[row] = v_score;
    if (p.line_off) p.line_off[row] = ((uint64_t)linelen << 48) | (uint64_t)(g0 + tok0);
    if (NSEP >= 4 && p.idspan) p.idspan[row] = v_id;
  }
}

// NSEP = min_fields (3|4|5): separators a canonical line must have, one after each of the first min_fields fields.
//
// Per tile (tiles are independent: the first row of each tile comes from pass 1):
//  [A] the text window (tile + halos) arrives by ONE bulk asynchronous copy (cp.async.bulk, completes on an mbarrier),
//      issued by one thread a whole tile ahead into the other buffer -- staging costs the other 255 threads nothing;
//  [B] every thread builds the control-byte / NL bitmasks of its 32 bytes and counts the lines that START there;
//  [C] block scan -> every line start is written to a compact list (lstart[i] = window index of the i-th line);
//  [D] thread i parses line i and writes row base+i of the SoA columns: full warps (the last one excepted) whatever the
//      line length, perfectly coalesced stores, and the previous row -- needed for the chromosome-head test -- is simply
//      lstart[i-1].
// Two block barriers per tile.  The mask arrays are double-buffered so that a warp may start [B] of the next tile while
// others still parse; lstart is written only after barrier [S2], when every thread has left [D] of the previous tile.
template <int NSEP, bool WANT_SCORE>
__global__ void __launch_bounds__(P_THREADS, P_MINBLOCKS) k_parse(ParseParams p) {
  __shared__ __align__(128) unsigned char smbuf[2][P_BUF + 16];  // double-buffered text window
  __shared__ uint32_t                     ctlp2[2][P_NW + 4];
  __shared__ uint32_t                     lsw2[2][P_PRE / 32 + P_TILE / 32];  // line-start mask of window bytes [0, PRE+TILE)
  __shared__ uint32_t                     wsum2[2][P_THREADS / 32];
  __shared__ uint64_t                     base2[2];
  __shared__ uint16_t                     lstart[P_MAXROWS];
  __shared__ __align__(8) uint64_t        mbar[2];
  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t eff = p.scratch[SC_EFFLEN];  // bytes up to and including the last '\n'
  const unsigned char* text = reinterpret_cast<const unsigned char*>(p.text);

  // a tile whose whole window lies inside the text is moved by the copy engine; the first and the last tiles of a file
  // are staged by the threads themselves (zero-filled outside the text)
  auto bulkable = [&](uint32_t tile) { return tile - 1u < p.bulk_tiles; };  // tiles 1 .. bulk_tiles (host: parse_bed)
  auto issue = [&](uint32_t tile, int b) {  // one thread
    fence_proxy_async();                    // the buffer's earlier generic-proxy accesses are ordered before the async write
    bulk_g2s(smbuf[b], text + ((int64_t)tile * P_TILE - P_PRE), P_BUF, &mbar[b]);
  };
  auto stage_edge = [&](uint32_t tile, unsigned char* dst) {
    const int64_t g0 = (int64_t)tile * P_TILE - P_PRE;
    for (int v = tid; v < P_BUF / 16; v += P_THREADS) {
      const int64_t g = g0 + (int64_t)v * 16;
      uint32_t      w[4] = {0, 0, 0, 0};
      if (g >= 0 && (uint64_t)g + 16 <= p.nbytes_raw) {
        const uint4 q = *reinterpret_cast<const uint4*>(text + g);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
      } else if (g + 16 > 0 && (uint64_t)(g < 0 ? 0 : g) < p.nbytes_raw) {
#pragma unroll 1
        for (int i = 0; i < 16; i++) {
          const int64_t gi = g + i;
          if (gi >= 0 && (uint64_t)gi < p.nbytes_raw) w[i >> 2] |= (uint32_t)text[gi] << (8 * (i & 3));
        }
      }
      reinterpret_cast<uint4*>(dst)[v] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  };

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init_fence();
    if (p.line_off && blockIdx.x == 0) p.line_off[p.cap] = eff;  // sentinel: end of the last row's line
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x < p.ntiles && bulkable(blockIdx.x)) issue(blockIdx.x, 0);

  uint32_t phase = 0;  // bit b = parity of the next completion of mbar[b]
  int      buf = 0;
  for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, buf ^= 1) {
    const int64_t        ts = (int64_t)tile * P_TILE;
    const int64_t        g0 = ts - P_PRE;
    unsigned char* const sm = smbuf[buf];
    uint32_t* const      ctlp = ctlp2[buf];
    uint32_t* const      lsw = lsw2[buf];
    // ---- [B] masks of the window from pass 1 (issued before the wait on the text: the loads overlap it) -------------
    // word j of the window covers window bytes [32j, 32j+32); global word index = tile * (P_TILE/32) - P_PRE/32 + j
    const int      off = P_PRE + tid * 32;
    uint32_t       smask, cm;
    {
      const int64_t w0g = (int64_t)tile * (P_TILE / 32) - P_PRE / 32;
      const int64_t gw = w0g + P_PRE / 32 + tid;  // this thread's own 32 bytes
      const bool    in = (uint64_t)gw < p.nwords;
      cm = in ? __ldg(&p.cm_arr[gw]) : 0xFFFFFFFFu;  // beyond the text: zero bytes, i.e. control bytes
      smask = in ? __ldg(&p.ls_arr[gw]) : 0u;
      ctlp[tid] = cm;
      lsw[P_PRE / 32 + tid] = smask;
      if (tid < P_PRE / 32) {  // head halo: line starts only (the row in front of the tile's first row)
        const int64_t g = w0g + tid;
        lsw[tid] = (g >= 0 && (uint64_t)g < p.nwords) ? __ldg(&p.ls_arr[g]) : 0u;
      } else if (tid >= 32 && tid < 32 + P_POST / 32) {  // tail halo: control bytes only
        const int64_t g = w0g + P_PRE / 32 + P_TILE / 32 + (tid - 32);
        ctlp[P_TILE / 32 + tid - 32] = (uint64_t)g < p.nwords ? __ldg(&p.cm_arr[g]) : 0xFFFFFFFFu;
      } else if (tid >= 64 && tid < 68) {
        ctlp[P_TILE / 32 + P_POST / 32 + tid - 64] = 0;  // padding words read by the 64-bit line windows
      }
    }
    // ---- [A] this tile's text -----------------------------------------------------------------------------------
    if (bulkable(tile)) {
      mbar_wait(&mbar[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
    } else {
      stage_edge(tile, sm);  // the buffer is free: every thread passed [S2] of the previous tile after leaving tile-2
      __syncthreads();
    }
    Cursor cur{sm, g0, text, eff};
    const uint32_t cnt = __popc(smask);
    const uint32_t incl = warp_incl_scan(cnt);
    if (lane == 31) wsum2[buf][warp] = incl;
    if (tid == 0) base2[buf] = p.tile_base[tile];  // first row of this tile (pass 1)
    __syncthreads();  // [S2] masks, warp totals and the tile's first row visible; everyone has left [D] of the previous tile
    if (tid == 0 && tile + gridDim.x < p.ntiles && bulkable(tile + gridDim.x)) issue(tile + gridDim.x, buf ^ 1);

    // ---- [C] compact the line starts --------------------------------------------------------------------------
    uint32_t nl;
    {
      const uint32_t v = lane < P_THREADS / 32 ? wsum2[buf][lane] : 0u;
      uint32_t       ex = incl - cnt + __reduce_add_sync(0xffffffffu, lane < warp ? v : 0u);
      nl = __reduce_add_sync(0xffffffffu, v);
      for (uint32_t m = smask; m; m &= m - 1) lstart[ex++] = (uint16_t)(off + __ffs(m) - 1);
    }
    __syncthreads();  // [S3]
    const uint64_t base = base2[buf];

    // ---- [D] thread i parses line i ---------------------------------------------------------------------------
#pragma unroll 1
    for (uint32_t i = tid; i < nl; i += P_THREADS) {
      const TileView tv{sm, ctlp, lsw, g0, base};
      parse_tile_line<NSEP, WANT_SCORE>(p, tv, cur, i, lstart[i], i ? (int)lstart[i - 1] : -1);
    }
    // no barrier here: [S2] of the next tile is the point where every thread has left this tile's [D]
  }
}

// ---- pass 2, warp-specialised form ---------------------------------------------------------------------------------
// The same tiles and the same per-line work as k_parse, without block barriers: warp 0 PRODUCES a tile (issues the bulk
// copy of its text, loads the two masks of pass 1, turns the line-start mask into the list of line starts), warps 1..7
// CONSUME it (thread c of 224 parses lines c, c+224, ...).  Two buffers; full[b] (producer -> consumers: masks and line
// list ready), txt[b] (copy engine -> consumers: text landed), empty[b] (consumers -> producer: buffer free).  The producer
// runs up to two tiles ahead, so the consumers find their next tile ready and never wait for one another.
constexpr int WS_CONSUMERS = P_THREADS - 32;
constexpr int WS_LSCAP = 1024;  // line starts listed per tile; a tile with more (lines of < 8 bytes) selects bits from the mask
#ifndef BK_WS_MINBLOCKS
#define BK_WS_MINBLOCKS 5  // 48 registers: no spills; measured 2.81 ms against 2.91 ms with 6 CTAs of 40 registers
#endif
#ifndef BK_WS_STAGES
#define BK_WS_STAGES 2  // tile buffers per CTA (13.4 KB each): the producer runs up to this many tiles ahead
#endif
constexpr int WS_ST = BK_WS_STAGES;
template <int NSEP, bool WANT_SCORE>
__global__ void __launch_bounds__(P_THREADS, BK_WS_MINBLOCKS) k_parse_ws(ParseParams p) {
  __shared__ __align__(128) unsigned char smbuf[WS_ST][P_BUF + 16];
  __shared__ uint32_t                     ctlp2[WS_ST][P_NW + 4];
  __shared__ uint32_t                     lsw2[WS_ST][P_PRE / 32 + P_TILE / 32];
  __shared__ uint16_t                     lstart2[WS_ST][WS_LSCAP];
  __shared__ uint16_t                     wpre2[WS_ST][P_TILE / 32 + 1];  // line starts in front of every word of the tile
  __shared__ uint32_t                     nl2[WS_ST];
  __shared__ uint64_t                     base2[WS_ST];
  __shared__ __align__(8) uint64_t        txt[WS_ST], full[WS_ST], empty[WS_ST];
  const int            tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t       eff = p.scratch[SC_EFFLEN];
  const unsigned char* text = reinterpret_cast<const unsigned char*>(p.text);
  auto bulkable = [&](uint32_t tile) { return tile - 1u < p.bulk_tiles; };
  if (tid == 0) {
    for (int b = 0; b < WS_ST; b++) {
      mbar_init(&txt[b], 1);
      mbar_init(&full[b], 1);
      mbar_init(&empty[b], WS_CONSUMERS / 32);
    }
    mbar_init_fence();
    if (p.line_off && blockIdx.x == 0) p.line_off[p.cap] = eff;
  }
  __syncthreads();

  if (warp == 0) {
    // ---- producer ---------------------------------------------------------------------------------------------
    int      b = 0;
    uint32_t use = 0;  // tile number it of this CTA uses buffer it % WS_ST for the (it / WS_ST)-th time
    for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, b = b + 1 == WS_ST ? 0 : b + 1, use += b == 0) {
      mbar_wait(&empty[b], (use & 1u) ^ 1u);  // the consumers have left the tile that used this buffer WS_ST tiles ago
      unsigned char* const sm = smbuf[b];
      if (bulkable(tile)) {
        if (lane == 0) {
          fence_proxy_async();
          bulk_g2s(sm, text + ((int64_t)tile * P_TILE - P_PRE), P_BUF, &txt[b]);
        }
      } else {  // first / last tiles of the file: staged by this warp, zero-filled outside the text
        const int64_t g0 = (int64_t)tile * P_TILE - P_PRE;
        for (int v = lane; v < P_BUF / 16; v += 32) {
          const int64_t g = g0 + (int64_t)v * 16;
          uint32_t      w[4] = {0, 0, 0, 0};
          if (g >= 0 && (uint64_t)g + 16 <= p.nbytes_raw) {
            const uint4 q = *reinterpret_cast<const uint4*>(text + g);
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
          } else if (g + 16 > 0 && (uint64_t)(g < 0 ? 0 : g) < p.nbytes_raw) {
#pragma unroll 1
            for (int i = 0; i < 16; i++) {
              const int64_t gi = g + i;
              if (gi >= 0 && (uint64_t)gi < p.nbytes_raw) w[i >> 2] |= (uint32_t)text[gi] << (8 * (i & 3));
            }
          }
          reinterpret_cast<uint4*>(sm)[v] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      // masks of pass 1: lane L owns words 8L .. 8L+7 of the tile (word j covers window bytes [PRE+32j, PRE+32j+32))
      uint32_t* const ctlp = ctlp2[b];
      uint32_t* const lsw = lsw2[b];
      const int64_t   w0g = (int64_t)tile * (P_TILE / 32) - P_PRE / 32;
      uint32_t        ls[8];
      uint32_t        cnt = 0;
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const int     j = lane * 8 + t;
        const int64_t gw = w0g + P_PRE / 32 + j;
        const bool    in = (uint64_t)gw < p.nwords;
        ctlp[j] = in ? __ldg(&p.cm_arr[gw]) : 0xFFFFFFFFu;
        ls[t] = in ? __ldg(&p.ls_arr[gw]) : 0u;
        lsw[P_PRE / 32 + j] = ls[t];
        cnt += __popc(ls[t]);
      }
      if (lane < P_PRE / 32) {
        const int64_t g = w0g + lane;
        lsw[lane] = (g >= 0 && (uint64_t)g < p.nwords) ? __ldg(&p.ls_arr[g]) : 0u;
      } else if (lane < P_PRE / 32 + P_POST / 32) {
        const int64_t g = w0g + P_PRE / 32 + P_TILE / 32 + (lane - P_PRE / 32);
        ctlp[P_TILE / 32 + lane - P_PRE / 32] = (uint64_t)g < p.nwords ? __ldg(&p.cm_arr[g]) : 0xFFFFFFFFu;
      } else if (lane < P_PRE / 32 + P_POST / 32 + 4) {
        ctlp[P_TILE / 32 + P_POST / 32 + lane - (P_PRE / 32 + P_POST / 32)] = 0;
      }
      const uint32_t incl = warp_incl_scan(cnt);
      uint32_t       ex = incl - cnt;
      const uint32_t nl = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const int j = lane * 8 + t;
        wpre2[b][j] = (uint16_t)ex;
        for (uint32_t m = ls[t]; m; m &= m - 1) {
          if (ex < (uint32_t)WS_LSCAP) lstart2[b][ex] = (uint16_t)(P_PRE + 32 * j + __ffs(m) - 1);
          ex++;
        }
      }
      if (lane == 31) wpre2[b][P_TILE / 32] = (uint16_t)ex;
      if (lane == 0) {
        nl2[b] = nl;
        base2[b] = p.tile_base[tile];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[b]);
    }
    return;
  }

  // ---- consumers ------------------------------------------------------------------------------------------------
  const uint32_t ctid = (uint32_t)tid - 32u;
  uint32_t       tphase = 0;  // tphase bit b: parity of the next completion of txt[b]
  int            b = 0;
  uint32_t       use = 0;
  for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, b = b + 1 == WS_ST ? 0 : b + 1, use += b == 0) {
    mbar_wait(&full[b], use & 1u);
    if (bulkable(tile)) {
      mbar_wait(&txt[b], (tphase >> b) & 1u);
      tphase ^= 1u << b;
    }
    const int64_t   g0 = (int64_t)tile * P_TILE - P_PRE;
    const TileView  tv{smbuf[b], ctlp2[b], lsw2[b], g0, base2[b]};
    const Cursor    cur{smbuf[b], g0, text, eff};
    const uint32_t  nl = nl2[b];
    const uint16_t* lst = lstart2[b];
    auto start_of = [&](uint32_t i) -> int {  // window index of line i of the tile
      if (nl <= (uint32_t)WS_LSCAP) return lst[i];
      int lo = 0, hi = P_TILE / 32;  // last word with wpre <= i
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (wpre2[b][mid] <= i) lo = mid; else hi = mid;
      }
      uint32_t m = lsw2[b][P_PRE / 32 + lo];
      for (uint32_t k = i - wpre2[b][lo]; k; k--) m &= m - 1;
      return P_PRE + 32 * lo + __ffs(m) - 1;
    };
#pragma unroll 1
    for (uint32_t i = ctid; i < nl; i += WS_CONSUMERS) parse_tile_line<NSEP, WANT_SCORE>(p, tv, cur, i, start_of(i), i ? start_of(i - 1) : -1);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[b]);
  }
}

// effective length = index of the last '\n' + 1 (an unterminated last line is not a record: the reference's
// iterator tests feof() after the read, AllocateIterator_BED_starch.hpp:172-187).  One warp, backwards.
__global__ void k_efflen(const unsigned char* text, uint64_t nbytes, uint64_t* scratch) {
  const int lane = threadIdx.x;
  int64_t   hi = (int64_t)nbytes;
  while (hi > 0) {
    int64_t  i = hi - 1 - lane;
    bool     hit = i >= 0 && text[i] == '\n';
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      if (lane == 0) scratch[SC_EFFLEN] = (uint64_t)(hi - (__ffs(m) - 1));
      return;
    }
    hi -= 32;
  }
  if (lane == 0) scratch[SC_EFFLEN] = 0;
}

// inclusive running max of `in` within each chromosome run (the index that bounds candidate windows for nested
// intervals, north_star item 3).  max is idempotent, so no carry chain between tiles is needed: rows are cut into
// warp ranges of PM_RANGE rows that never cross a chromosome boundary; pass 1 reduces every range to its maximum,
// pass 2 gives a range the maximum of the earlier ranges of its chromosome as carry (a short reduction over the
// pass-1 array, L2 resident) and scans its own rows, 32 at a time with coalesced loads and stores.
__global__ void __launch_bounds__(PM_THREADS) k_pmax_reduce(const uint32_t* __restrict__ in, const PmRun* __restrict__ runs,
                                                            int nruns, uint64_t nranges, uint32_t* __restrict__ range_max) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * PM_THREADS + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * PM_THREADS) >> 5;
  for (uint64_t r = w0; r < nranges; r += nw) {
    uint64_t a, b, first;
    pm_locate(runs, nruns, r, a, b, first);
    uint32_t m = 0;
    for (uint64_t k = a + lane; k < b; k += 128) {  // four independent loads in flight per lane
      const uint32_t v0 = __ldg(&in[k]);
      const uint32_t v1 = k + 32 < b ? __ldg(&in[k + 32]) : 0u;
      const uint32_t v2 = k + 64 < b ? __ldg(&in[k + 64]) : 0u;
      const uint32_t v3 = k + 96 < b ? __ldg(&in[k + 96]) : 0u;
      m = max(max(m, v0), max(max(v1, v2), v3));
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) range_max[r] = m;
  }
}

__global__ void __launch_bounds__(PM_THREADS) k_pmax(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                     const PmRun* __restrict__ runs, int nruns, uint64_t nranges,
                                                     const uint32_t* __restrict__ range_max) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * PM_THREADS + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * PM_THREADS) >> 5;
  for (uint64_t r = w0; r < nranges; r += nw) {
    uint64_t a, b, first;
    pm_locate(runs, nruns, r, a, b, first);
    uint32_t carry = 0;
    for (uint64_t q = first + lane; q < r; q += 32) carry = max(carry, __ldg(&range_max[q]));
    carry = __reduce_max_sync(0xffffffffu, carry);
    for (uint64_t k0 = a; k0 < b; k0 += 128) {
      uint32_t v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint64_t k = k0 + 32 * u + lane;
        v[u] = k < b ? __ldg(&in[k]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        uint32_t x = v[u];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
          if (lane >= d) x = max(x, y);
        }
        x = max(x, carry);
        const uint64_t k = k0 + 32 * u + lane;
        if (k < b) out[k] = x;
        carry = __shfl_sync(0xffffffffu, x, 31);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------------------
int reset_scratch(bk_ctx* ctx) {
  BK_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, SC_N * sizeof(uint64_t), ctx->stream));
  return BK_OK;
}
int read_scratch(bk_ctx* ctx) {
  BK_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, SC_N * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                               ctx->stream));
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

static int grid_for(const bk_ctx* ctx, const void* kernel, int threads, uint32_t ntiles) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
  if (per_sm < 1) per_sm = 1;
  uint64_t g = (uint64_t)ctx->sms * per_sm;
  return (int)(ntiles < g ? (ntiles ? ntiles : 1) : g);
}

int parse_bed(bk_ctx* ctx, bk_bed* bed, uint64_t nbytes_raw) {
  const unsigned char* text = reinterpret_cast<const unsigned char*>(bed->d_text);
  BK_TRY(reset_scratch(ctx));
  if (nbytes_raw == 0) {
    bed->nrows = 0;
    bed->nbytes = 0;
    return BK_OK;
  }
  const uint32_t ntiles = (uint32_t)((nbytes_raw + P_TILE - 1) / P_TILE);
  const uint32_t heads_cap = 1u << 16;
  constexpr size_t kHeadsInline = 64;
  HeadRec* const heads_pre = reinterpret_cast<HeadRec*>(ctx->h_scratch + SC_N);  // pinned
  HeadRec*       d_heads = dalloc<HeadRec>(ctx, heads_cap);
  // pass 1 geometry: one warp per contiguous range of tiles
  const uint32_t max_warps = (uint32_t)ctx->sms * 8 * (CR_THREADS / 32);
  const uint32_t tiles_per_warp = (ntiles + max_warps - 1) / max_warps;
  const uint32_t nwarps = (ntiles + tiles_per_warp - 1) / tiles_per_warp;
  uint32_t*      d_lpre = dalloc<uint32_t>(ctx, ntiles);
  uint64_t*      d_wtot = dalloc<uint64_t>(ctx, nwarps);
  uint64_t*      d_wbase = dalloc<uint64_t>(ctx, (size_t)nwarps + 1);
  const uint64_t nwords = (nbytes_raw + 31) / 32;
  uint32_t*      d_cm = dalloc<uint32_t>(ctx, nwords);
  uint32_t*      d_ls = dalloc<uint32_t>(ctx, nwords);
  if (!d_heads || !d_lpre || !d_wtot || !d_wbase || !d_cm || !d_ls) return BK_ERR_NOMEM;
  // pass 1: effective length, rows per tile, exclusive scan -> exact row count
  prof_begin(ctx, "k_efflen");
  k_efflen<<<1, 32, 0, ctx->stream>>>(text, nbytes_raw, ctx->d_scratch);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  prof_begin(ctx, "k_count_rows");
  k_count_rows<<<(nwarps * 32 + CR_THREADS - 1) / CR_THREADS, CR_THREADS, 0, ctx->stream>>>(
      text, nbytes_raw, ctx->d_scratch, d_lpre, d_wtot, ntiles, tiles_per_warp, nwarps, (bed->cols & BK_LOAD_HEADERS) ? 1 : 0,
      (bed->cols & BK_LOAD_SORTBED) ? 1 : 0, d_cm, d_ls);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  prof_begin(ctx, "k_scan_warps");
  k_scan_totals<SC_NROWS><<<1, 1024, 0, ctx->stream>>>(d_wtot, d_wbase, nwarps, ctx->d_scratch);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  uint64_t* d_tbase = dalloc<uint64_t>(ctx, ntiles);
  if (!d_tbase) return BK_ERR_NOMEM;
  k_tile_base<<<(ntiles + 255) / 256, 256, 0, ctx->stream>>>(d_lpre, d_wbase, tiles_per_warp, ntiles, d_tbase);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));
  bed->nrows = ctx->h_scratch[SC_NROWS];
  bed->nbytes = ctx->h_scratch[SC_EFFLEN];
  dfree(ctx, d_wtot);
  dfree(ctx, d_lpre);
  dfree(ctx, d_wbase);
  const uint64_t cap = bed->nrows;
  {
    bed->start = dalloc<uint32_t>(ctx, cap);
    bed->end = dalloc<uint32_t>(ctx, cap);
    if (bed->cols & BK_COL_SCORE) bed->score = dalloc<double>(ctx, cap);
    if (bed->cols & BK_COL_LINE) bed->line_off = dalloc<uint64_t>(ctx, cap + 1);
    if (bed->cols & BK_COL_ID) bed->idspan = dalloc<uint32_t>(ctx, cap);
    if (!bed->start || !bed->end || ((bed->cols & BK_COL_SCORE) && !bed->score) ||
        ((bed->cols & BK_COL_LINE) && !bed->line_off) || ((bed->cols & BK_COL_ID) && !bed->idspan))
      return BK_ERR_NOMEM;

    // pass 2: tokenise
    ParseParams p{};
    p.text = bed->d_text;
    p.nbytes_raw = nbytes_raw;
    p.min_fields = bed->min_fields;
    p.cols = bed->cols;
    p.start = bed->start;
    p.end = bed->end;
    p.score = bed->score;
    p.line_off = bed->line_off;
    p.idspan = bed->idspan;
    p.cap = cap;
    p.ntiles = ntiles;
    p.scratch = ctx->d_scratch;
    p.heads = d_heads;
    p.heads_cap = heads_cap;
    p.tile_base = d_tbase;
    p.cm_arr = d_cm;
    p.ls_arr = d_ls;
    p.nwords = nwords;
    // tiles whose whole window [tile*P_TILE - P_PRE, +P_BUF) lies inside the text travel by bulk copy: tiles 1 .. bulk_tiles
    p.bulk_tiles = nbytes_raw >= (uint64_t)(P_TILE + P_POST) ? (uint32_t)((nbytes_raw - P_TILE - P_POST) / P_TILE) : 0u;
    prof_begin(ctx, "k_parse");
    {
      const bool sc = (p.cols & BK_COL_SCORE) != 0;
      static const bool ws = getenv("BEDKIT_PARSE_BLOCK") == nullptr;  // BEDKIT_PARSE_BLOCK=1: the block-barrier form (A/B measurements)
#define BK_PARSE(N, S)                                                                                                        \
  do {                                                                                                                        \
    if (ws) k_parse_ws<N, S><<<grid_for(ctx, (const void*)k_parse_ws<N, S>, P_THREADS, p.ntiles), P_THREADS, 0, ctx->stream>>>(p); \
    else k_parse<N, S><<<grid_for(ctx, (const void*)k_parse<N, S>, P_THREADS, p.ntiles), P_THREADS, 0, ctx->stream>>>(p);          \
  } while (0)
      if (p.min_fields == 3) BK_PARSE(3, false);
      else if (p.min_fields == 4) BK_PARSE(4, false);
      else if (sc) BK_PARSE(5, true);
      else BK_PARSE(5, false);
#undef BK_PARSE
    }
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    // one host round trip: the first heads travel with the scratch words (a genome has a few dozen chromosome runs)
    static_assert(kHeadsInline * sizeof(HeadRec) <= kHostScratchExtra, "inline heads must fit the pinned scratch tail");
    BK_CUDA(ctx, cudaMemcpyAsync(heads_pre, d_heads, kHeadsInline * sizeof(HeadRec), cudaMemcpyDeviceToHost, ctx->stream));
    BK_TRY(read_scratch(ctx));
    dfree(ctx, d_tbase);
    dfree(ctx, d_cm);
    dfree(ctx, d_ls);
    const uint64_t* h = ctx->h_scratch;
    if (h[SC_ERR_CODE]) {
      int code = (int)h[SC_ERR_CODE];
      dfree(ctx, d_heads);
      const char* what = code == BK_ERR_COORD_RANGE ? "coordinate does not fit the 32-bit device layout"
                         : code == BK_ERR_UNSUPPORTED ? "score literal outside the exact device strtod path"
                                                      : "line is not chrom<ws>start<ws>end[...] with enough fields";
      return fail(ctx, code, "BED parse error at row %llu: %s", (unsigned long long)h[SC_ERR_ROW] + 1, what);
    }
  }
  // (line_off[nrows] = effective length is written by k_parse itself)
  // chromosome runs
  uint64_t nheads = ctx->h_scratch[SC_NHEADS];
  if (nheads > heads_cap) {
    dfree(ctx, d_heads);
    return fail(ctx, BK_ERR_UNSUPPORTED, "more than %u chromosome runs in one file", heads_cap);
  }
  std::vector<HeadRec> heads(nheads);
  if (nheads <= kHeadsInline) {
    std::copy(heads_pre, heads_pre + nheads, heads.begin());
  } else {
    BK_CUDA(ctx, cudaMemcpyAsync(heads.data(), d_heads, nheads * sizeof(HeadRec), cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  dfree(ctx, d_heads);
  std::sort(heads.begin(), heads.end(), [](const HeadRec& a, const HeadRec& b) { return a.row < b.row; });
  bed->runs.clear();
  for (size_t i = 0; i < heads.size(); i++) {
    ChromRun r;
    r.name.assign(heads[i].name, heads[i].len);
    r.row_begin = heads[i].row;
    r.row_end = (i + 1 < heads.size()) ? heads[i + 1].row : bed->nrows;
    bed->runs.push_back(r);
  }
  for (size_t i = 1; i < bed->runs.size(); i++) {
    if (strcmp(bed->runs[i - 1].name.c_str(), bed->runs[i].name.c_str()) >= 0)
      return fail(ctx, BK_ERR_UNSORTED, "chromosome '%s' (row %llu) follows '%s': input is not sorted per sort-bed",
                  bed->runs[i].name.c_str(), (unsigned long long)bed->runs[i].row_begin + 1,
                  bed->runs[i - 1].name.c_str());
  }
  return BK_OK;
}

// out[k] = max(in[j] : j <= k, j in the same chromosome run as k)
int seg_prefix_max(bk_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n, const std::vector<ChromRun>& runs) {
  if (n == 0) return BK_OK;
  std::vector<PmRun> pr;
  uint64_t           nranges = 0;
  for (const ChromRun& r : runs) {
    pr.push_back({r.row_begin, r.row_end, nranges});
    nranges += (r.row_end - r.row_begin + PM_RANGE - 1) / PM_RANGE;
  }
  if (nranges == 0) return BK_OK;
  PmRun*    d_runs = dalloc<PmRun>(ctx, pr.size());
  uint32_t* d_rmax = dalloc<uint32_t>(ctx, nranges);
  if (!d_runs || !d_rmax) return BK_ERR_NOMEM;
  BK_TRY(upload_params(ctx, d_runs, pr.data(), pr.size() * sizeof(PmRun)));
  const uint64_t want = (nranges + PM_THREADS / 32 - 1) / (PM_THREADS / 32);
  const uint32_t ga = (uint32_t)std::min<uint64_t>(want, (uint64_t)grid_for(ctx, (const void*)k_pmax_reduce, PM_THREADS, 0xFFFFFFFFu));
  const uint32_t gb = (uint32_t)std::min<uint64_t>(want, (uint64_t)grid_for(ctx, (const void*)k_pmax, PM_THREADS, 0xFFFFFFFFu));
  prof_begin(ctx, "k_pmax_reduce");
  k_pmax_reduce<<<ga, PM_THREADS, 0, ctx->stream>>>(in, d_runs, (int)pr.size(), nranges, d_rmax);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  prof_begin(ctx, "k_pmax");
  k_pmax<<<gb, PM_THREADS, 0, ctx->stream>>>(in, out, d_runs, (int)pr.size(), nranges, d_rmax);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  dfree(ctx, d_runs);  // stream-ordered reuse; the table went through the pinned ring
  dfree(ctx, d_rmax);
  return BK_OK;
}

// max end of every block of 32 consecutive rows (blocks on the global row index, chromosome runs ignored: a block that
// straddles two chromosomes is merely less often skippable).  The window scans skip blocks no row of which reaches the
// reference row.
__global__ void __launch_bounds__(256) k_block_max(const uint32_t* __restrict__ end, uint64_t n, uint32_t* __restrict__ bmax) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * 256) >> 5;
  const uint64_t nblocks = (n + 31) >> 5;
  for (uint64_t b = w0; b < nblocks; b += nw) {
    const uint64_t k = (b << 5) + lane;
    const uint32_t v = __reduce_max_sync(0xffffffffu, k < n ? __ldg(&end[k]) : 0u);
    if (lane == 0) bmax[b] = v;
  }
}

int ensure_pmax(bk_ctx* ctx, const bk_bed* cbed) {
  bk_bed* bed = const_cast<bk_bed*>(cbed);
  if (bed->pmax_end || bed->nrows == 0) return BK_OK;
  bed->pmax_end = dalloc<uint32_t>(ctx, bed->nrows);
  if (!bed->pmax_end) return BK_ERR_NOMEM;
  return seg_prefix_max(ctx, bed->end, bed->pmax_end, bed->nrows, bed->runs);
}

// the block-max index is only worth its pass for dense map files (bk_bedmap decides)
int ensure_bmax(bk_ctx* ctx, const bk_bed* cbed) {
  bk_bed* bed = const_cast<bk_bed*>(cbed);
  if (bed->bmax_end || bed->nrows == 0) return BK_OK;
  const uint64_t nblocks = (bed->nrows + 31) / 32;
  bed->bmax_end = dalloc<uint32_t>(ctx, nblocks + 2);
  if (!bed->bmax_end) return BK_ERR_NOMEM;
  const uint64_t want = (nblocks + 7) / 8;
  prof_begin(ctx, "k_block_max");
  k_block_max<<<grid_for(ctx, (const void*)k_block_max, 256, (uint32_t)std::min<uint64_t>(want, 0xFFFFFFFFu)), 256, 0, ctx->stream>>>(
      bed->end, bed->nrows, bed->bmax_end);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  return BK_OK;
}

}  // namespace bk
