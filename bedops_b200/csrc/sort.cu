// sort.cu -- sort-bed on the device and the radix sort the other operators borrow.
//
// Replaces processData / lexSortBedData / printBed of applications/bed/sort-bed/src/SortDetails.cpp:530-1208: read
// BED rows in any order, order them by chromosome (strcmp), start, end, then the rest of the line (strcmp, a row
// without a rest first; BedCoordData::bcd_cmp, Structures.hpp:47-76) and print "chrom\tstart\tend[\trest]\n".
//
// Design (B200): the text is parsed by the ordinary reader in its sort-bed mode (rows in any order, only empty lines
// skipped); k_sort_validate re-reads by sort-bed's own grammar every row the reader did not find canonical and checks
// end > start; chromosome names go through a device hash table (a genome has dozens of names, not millions: one
// atomic per first sighting), the host ranks the few names by strcmp; rows are ordered by ONE packed key
// (rank | start | length) with a least-significant-digit radix sort of exactly the bits in use (8 bits per pass; count
// per tile -> one-CTA scan -> stable scatter with a shared-memory local sort so that every digit leaves as a coalesced
// run); rows with equal coordinates are then ordered by their rest; the writer echoes row perm[i].
#include <algorithm>
#include "common.cuh"
#include "emit.cuh"
#include "parse.cuh"
#include "sort.cuh"

namespace bk {

// ---------------------------------------------------------------------------------------------------------
// radix sort of (u64 key, u32 value) pairs by key bits [0, nbits)
// ---------------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_SUB = 2048;    // keys ranked and locally sorted at a time (8 per thread)
constexpr int RS_TILE = 16384;  // keys per tile = one column of the digit table

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint64_t* __restrict__ keys, uint64_t n, int shift,
                                                           uint32_t* __restrict__ table, uint32_t ntiles) {
  __shared__ uint32_t hist[256];
  const int tid = threadIdx.x;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    hist[tid] = 0;
    __syncthreads();
    const uint64_t t0 = (uint64_t)tile * RS_TILE, t1 = t0 + RS_TILE < n ? t0 + RS_TILE : n;
    for (uint64_t j = t0 + tid; j < t1; j += RS_THREADS) atomicAdd(&hist[(uint32_t)(keys[j] >> shift) & 255u], 1u);
    __syncthreads();
    table[(size_t)tid * ntiles + tile] = hist[tid];  // digit-major: the scan order is (digit, tile)
    __syncthreads();
  }
}

// in-place exclusive scan of a u32 array by one CTA (the digit table: 256 x ntiles entries)
__global__ void __launch_bounds__(1024) k_scan_u32(uint32_t* __restrict__ a, uint64_t n) {
  __shared__ uint32_t part[1024];
  const uint32_t tid = threadIdx.x;
  const uint64_t per = (n + 1023) / 1024;
  const uint64_t b = (uint64_t)tid * per < n ? (uint64_t)tid * per : n, e = b + per < n ? b + per : n;
  uint32_t       s = 0;
  for (uint64_t i = b; i < e; i++) s += a[i];
  part[tid] = s;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    const uint32_t v = tid >= d ? part[tid - d] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  uint32_t run = tid ? part[tid - 1] : 0;
  for (uint64_t i = b; i < e; i++) {
    const uint32_t v = a[i];
    a[i] = run;
    run += v;
  }
}

// Stable scatter of one pass.  A tile is walked in sub-tiles of RS_SUB keys; inside a sub-tile warp w owns keys
// [256w, 256w+256) and reads them 32 at a time in index order, so (warp, round, lane) order IS key order.  The rank
// of a key among the equal digits of its warp comes from MATCH.ANY (peers with the same digit) and a per-warp digit
// counter; a prefix over the warps and over the digits gives its position in the locally sorted sub-tile (shared
// memory), from which every digit's keys leave as one contiguous run at the tile's running cursor of that digit.
template <bool HAS_VAL>
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                              uint64_t* __restrict__ kout, uint32_t* __restrict__ vout, uint64_t n,
                                                              int shift, const uint32_t* __restrict__ table, uint32_t ntiles) {
  __shared__ uint32_t wcnt[RS_THREADS / 32][256];
  __shared__ uint32_t dbase[256], tot[256], cursor[256], wsum[RS_THREADS / 32];
  __shared__ uint64_t skey[RS_SUB];
  __shared__ uint32_t sval[HAS_VAL ? RS_SUB : 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    cursor[tid] = table[(size_t)tid * ntiles + tile];
    const uint64_t t0 = (uint64_t)tile * RS_TILE, t1 = t0 + RS_TILE < n ? t0 + RS_TILE : n;
    for (uint64_t s0 = t0; s0 < t1; s0 += RS_SUB) {
      const uint32_t cntk = (uint32_t)(t1 - s0 < RS_SUB ? t1 - s0 : RS_SUB);
#pragma unroll
      for (int w = 0; w < RS_THREADS / 32; w++) wcnt[w][tid] = 0;
      __syncthreads();
      uint64_t k[8];
      uint32_t v[8], lr[8];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const uint32_t idx = (uint32_t)warp * 256u + (uint32_t)r * 32u + (uint32_t)lane;
        const bool     valid = idx < cntk;
        k[r] = valid ? kin[s0 + idx] : 0ull;
        v[r] = (HAS_VAL && valid) ? vin[s0 + idx] : 0u;
        const uint32_t d = valid ? ((uint32_t)(k[r] >> shift) & 255u) : 256u;  // lanes past the end form their own group
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int      leader = __ffs(peers) - 1;
        uint32_t       old = 0;
        if (valid && lane == leader) {
          old = wcnt[warp][d];
          wcnt[warp][d] = old + (uint32_t)__popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        lr[r] = old + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncwarp();  // the counter update is visible to the next round's leaders
      }
      __syncthreads();
      {  // thread d: exclusive prefix over the warps, then over the digits
        uint32_t acc = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) {
          const uint32_t t = wcnt[w][tid];
          wcnt[w][tid] = acc;
          acc += t;
        }
        tot[tid] = acc;
        const uint32_t incl = warp_incl_scan(acc);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += wsum[w];
        dbase[tid] = base + incl - acc;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const uint32_t idx = (uint32_t)warp * 256u + (uint32_t)r * 32u + (uint32_t)lane;
        if (idx < cntk) {
          const uint32_t d = (uint32_t)(k[r] >> shift) & 255u;
          const uint32_t pos = dbase[d] + wcnt[warp][d] + lr[r];
          skey[pos] = k[r];
          if (HAS_VAL) sval[pos] = v[r];
        }
      }
      __syncthreads();
      for (uint32_t j = tid; j < cntk; j += RS_THREADS) {
        const uint64_t key = skey[j];
        const uint32_t d = (uint32_t)(key >> shift) & 255u;
        const uint64_t g = (uint64_t)cursor[d] + (j - dbase[d]);
        kout[g] = key;
        if (HAS_VAL) vout[g] = sval[j];
      }
      __syncthreads();
      cursor[tid] += tot[tid];
    }
    __syncthreads();  // cursor is reloaded for the next tile
  }
}

// Sorts n pairs by key bits [0, nbits).  keys/vals and the two scratch buffers are swapped as needed; on return *keys
// and *vals point at the sorted data (either pair may be the one the caller allocated: free all four).
int radix_sort_pairs(bk_ctx* ctx, uint64_t** keys, uint32_t** vals, uint64_t** keys_alt, uint32_t** vals_alt, uint64_t n,
                     int nbits) {
  if (n < 2 || nbits <= 0) return BK_OK;
  if (n >= 0xFFFFFFFFull) return fail(ctx, BK_ERR_UNSUPPORTED, "more than 2^32-1 rows in one sort");
  const uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
  uint32_t*      table = dalloc<uint32_t>(ctx, (size_t)256 * ntiles);
  if (!table) return BK_ERR_NOMEM;
  const bool has_val = vals && *vals;
  const int  gh = grid_for_kernel((const void*)k_radix_hist, RS_THREADS, ntiles);
  const int  gs = has_val ? grid_for_kernel((const void*)k_radix_scatter<true>, RS_THREADS, ntiles)
                          : grid_for_kernel((const void*)k_radix_scatter<false>, RS_THREADS, ntiles);
  for (int shift = 0; shift < nbits; shift += 8) {
    prof_begin(ctx, "k_radix_hist");
    k_radix_hist<<<gh, RS_THREADS, 0, ctx->stream>>>(*keys, n, shift, table, ntiles);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    k_scan_u32<<<1, 1024, 0, ctx->stream>>>(table, (uint64_t)256 * ntiles);
    BK_LAUNCHED(ctx);
    prof_begin(ctx, "k_radix_scatter");
    if (has_val)
      k_radix_scatter<true><<<gs, RS_THREADS, 0, ctx->stream>>>(*keys, *vals, *keys_alt, *vals_alt, n, shift, table, ntiles);
    else
      k_radix_scatter<false><<<gs, RS_THREADS, 0, ctx->stream>>>(*keys, nullptr, *keys_alt, nullptr, n, shift, table, ntiles);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    std::swap(*keys, *keys_alt);
    if (has_val) std::swap(*vals, *vals_alt);
  }
  dfree(ctx, table);
  return BK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// sort-bed
// ---------------------------------------------------------------------------------------------------------
constexpr uint32_t kVerbatim = 0x80000000u;  // dataoff flag: the whole input line is the output line
constexpr uint32_t HT_SIZE = 1u << 20;       // chromosome-name table slots (at most half are used)
constexpr uint64_t HT_EMPTY = ~0ull;
struct HtSlot {
  unsigned long long hash, rep;  // rep: a row that carries the name
};

__device__ __forceinline__ bool sb_sep(char c) { return c == '\t' || c == ' '; }       // strpbrk(..., "\t ")
__device__ __forceinline__ bool sb_space(char c) { return is_ws((unsigned char)c); }    // sscanf's white space (the NL ends the line)

// sort-bed's own reading of one line that starts at p (SortDetails.cpp:638-779, :833-853); 0 ok, 1 = a line the
// reference rejects, 2 = a coordinate outside the 32-bit device layout.  doff = offset of the rest (0: none), e3 = offset
// of the byte after the end coordinate.
__device__ __noinline__ int sortbed_strict(const char* __restrict__ p, uint32_t& st, uint32_t& en, uint32_t& doff) {
  doff = 0;
  if (sb_sep(p[0])) return 1;
  uint32_t c = 0;
  while (!sb_sep(p[c]) && p[c] != '\n') c++;
  if (p[c] == '\n' || c > 127) return 1;
  uint64_t val[2];
  uint32_t q = c + 1;
  for (int f = 0; f < 2; f++) {
    uint32_t d = q;
    while (!sb_sep(p[d]) && p[d] != '\n') d++;
    if (f == 0 && p[d] == '\n') return 1;  // no separator after the start coordinate
    const uint32_t len = d - q;
    if (len == 0 || len > 12) return 1;
    uint64_t acc = 0;
    for (uint32_t k = q; k < d; k++) {
      if (!is_digit((unsigned char)p[k])) return 1;
      acc = acc * 10 + (uint64_t)(p[k] - '0');
    }
    val[f] = acc;
    q = d + (f == 0 ? 1 : 0);
  }
  if (val[1] <= val[0]) return 1;
  while (sb_space(p[q])) q++;
  if (p[q] != '\n') {
    doff = q;
    uint32_t k = q;
    while (!sb_sep(p[k]) && p[k] != '\n') k++;
    if (k - q > 16383) return 1;  // ID_NAME_LEN
  }
  if (val[0] >= 0xFFFFFFFFull || val[1] >= 0xFFFFFFFFull) return 2;
  st = (uint32_t)val[0];
  en = (uint32_t)val[1];
  return 0;
}

struct SortParams {
  const char*     text;
  uint64_t        n;
  uint32_t*       start;
  uint32_t*       end;
  const uint64_t* line;
  const uint32_t* e3;       // reader: offset of the separator after the end coordinate (canonical rows)
  uint32_t*       dataoff;  // offset of the rest from the line start (0: none) | kVerbatim
  uint64_t*       scratch;
  HtSlot*         table;
  uint32_t*       slot;     // [n] table slot of the row's chromosome
};

// scratch: SC_COUNT_A = max start, SC_COUNT_B = max length, SC_COUNT_C = ~(first row the reference rejects),
// SC_COUNT_D = ~(first row outside the 32-bit layout)   (~row so that atomicMax finds the FIRST row; 0 = none)
__global__ void __launch_bounds__(256) k_sort_validate(SortParams p) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  uint32_t       st = 0, len = 0;
  if (i < p.n) {
    const uint64_t lo = p.line[i];
    const char*    ln = p.text + (lo & kLineOffMask);
    const uint32_t l16 = (uint32_t)(lo >> 48);
    uint32_t       en = 0, doff = 0;
    int            err = 0;
    if (l16 <= 16000u) {  // canonical for the reader: single TABs, plain digits; what is left is the rest and end > start
      st = p.start[i];
      en = p.end[i];
      const uint32_t e = p.e3[i];
      if (en <= st) err = 1;
      if (ln[e] == '\n') {
        doff = kVerbatim;
      } else {
        uint32_t q = e + 1;
        while (sb_space(ln[q])) q++;
        if (ln[q] != '\n') doff = q | (q == e + 1 ? kVerbatim : 0u);
      }
    } else {
      err = sortbed_strict(ln, st, en, doff);
      if (!err) {
        p.start[i] = st;
        p.end[i] = en;
      }
    }
    p.dataoff[i] = doff;
    if (err) atomicMax(reinterpret_cast<unsigned long long*>(&p.scratch[err == 1 ? SC_COUNT_C : SC_COUNT_D]), ~(unsigned long long)i);
    len = err ? 0u : en - st;
    if (err) st = 0;
  }
  const uint32_t ms = __reduce_max_sync(0xffffffffu, st), ml = __reduce_max_sync(0xffffffffu, len);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(&p.scratch[SC_COUNT_A]), (unsigned long long)ms);
    atomicMax(reinterpret_cast<unsigned long long*>(&p.scratch[SC_COUNT_B]), (unsigned long long)ml);
  }
}

// chromosome name of every row -> slot of a device hash table (FNV-1a of the name, linear probing).  Only the first
// sighting of a name costs an atomic; everybody else finds the hash already in place.
__global__ void __launch_bounds__(256) k_chrom_insert(SortParams p) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const char* ln = p.text + (p.line[i] & kLineOffMask);
  uint64_t    h = 1469598103934665603ull;
  for (uint32_t k = 0; !sb_sep(ln[k]) && ln[k] != '\n'; k++) h = (h ^ (unsigned char)ln[k]) * 1099511628211ull;
  if (h == HT_EMPTY) h = 1;
  uint32_t s = (uint32_t)(h ^ (h >> 32)) & (HT_SIZE - 1);
  for (uint32_t probe = 0; probe < HT_SIZE / 2; probe++, s = (s + 1) & (HT_SIZE - 1)) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&p.table[s].hash);
    if (cur == HT_EMPTY) {
      cur = atomicCAS(&p.table[s].hash, HT_EMPTY, (unsigned long long)h);
      if (cur == HT_EMPTY) {
        p.table[s].rep = i;
        atomicAdd(reinterpret_cast<unsigned long long*>(&p.scratch[SC_NHEADS]), 1ull);
        cur = h;
      }
    }
    if (cur == h) {
      p.slot[i] = s;
      return;
    }
  }
  dev_set_error(p.scratch, BK_ERR_UNSUPPORTED, i);  // more distinct names than the table takes
}

// key = rank | start | length; equal hashes are checked to be equal names
__global__ void __launch_bounds__(256) k_sort_keys(SortParams p, const uint32_t* __restrict__ slot_rank, int start_bits, int len_bits,
                                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ rank_out) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const uint32_t s = p.slot[i];
  const uint64_t rep = p.table[s].rep;
  if (rep != i) {
    const char* a = p.text + (p.line[i] & kLineOffMask);
    const char* b = p.text + (p.line[rep] & kLineOffMask);
    uint32_t    k = 0;
    while (!sb_sep(a[k]) && a[k] != '\n' && a[k] == b[k]) k++;
    const bool ea = sb_sep(a[k]) || a[k] == '\n', eb = sb_sep(b[k]) || b[k] == '\n';
    if (!(ea && eb)) dev_set_error(p.scratch, BK_ERR_UNSUPPORTED, i);  // two names, one 64-bit hash
  }
  const uint32_t r = slot_rank[s];
  rank_out[i] = r;
  const uint32_t st = p.start[i], len = p.end[i] - st;
  keys[i] = ((uint64_t)r << (start_bits + len_bits)) | ((uint64_t)st << len_bits) | (uint64_t)len;
  vals[i] = (uint32_t)i;
}

// the rest of row a against the rest of row b: strcmp, a row without a rest first (bcd_cmp, Structures.hpp:61-75)
__device__ __forceinline__ int cmp_rest(const SortParams& p, uint32_t a, uint32_t b) {
  const uint32_t da = p.dataoff[a] & ~kVerbatim, db = p.dataoff[b] & ~kVerbatim;
  if (!da || !db) return da ? 1 : (db ? -1 : 0);
  const char* x = p.text + (p.line[a] & kLineOffMask) + da;
  const char* y = p.text + (p.line[b] & kLineOffMask) + db;
  for (uint32_t k = 0;; k++) {
    const unsigned char cx = (unsigned char)x[k], cy = (unsigned char)y[k];
    const bool          ex = cx == '\n', ey = cy == '\n';
    if (ex || ey) return ex ? (ey ? 0 : -1) : 1;
    if (cx != cy) return cx < cy ? -1 : 1;
  }
}

// rows with equal (chromosome, start, end) are adjacent after the key sort: the thread at the head of such a run orders
// it by the rest of the line (insertion sort for the usual pair or triple, heap sort beyond 16 rows).  Equal rows are
// indistinguishable in the output, so stability does not matter.
__global__ void __launch_bounds__(256) k_tie_fix(SortParams p, const uint64_t* __restrict__ keys, uint32_t* __restrict__ perm) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const uint64_t key = keys[i];
  if (i > 0 && keys[i - 1] == key) return;
  if (i + 1 >= p.n || keys[i + 1] != key) return;
  uint64_t j = i + 2;
  while (j < p.n && keys[j] == key) j++;
  uint32_t*      a = perm + i;
  const uint64_t k = j - i;
  if (k <= 16) {
    for (uint64_t x = 1; x < k; x++) {
      const uint32_t v = a[x];
      uint64_t       y = x;
      while (y > 0 && cmp_rest(p, a[y - 1], v) > 0) {
        a[y] = a[y - 1];
        y--;
      }
      a[y] = v;
    }
    return;
  }
  auto sift = [&](uint64_t root, uint64_t end) {
    while (true) {
      uint64_t child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && cmp_rest(p, a[child], a[child + 1]) < 0) child++;
      if (cmp_rest(p, a[root], a[child]) >= 0) break;
      const uint32_t t = a[root];
      a[root] = a[child];
      a[child] = t;
      root = child;
    }
  };
  for (uint64_t s = k / 2; s-- > 0;) sift(s, k);
  for (uint64_t e = k - 1; e > 0; e--) {
    const uint32_t t = a[0];
    a[0] = a[e];
    a[e] = t;
    sift(0, e);
  }
}

struct SortRow {  // printBed, SortDetails.cpp:1120-1140
  SortParams      p;
  const uint32_t* perm;
  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& s) const {
    const uint32_t r = perm[i];
    const uint64_t lo = p.line[r];
    const char*    ln = p.text + (lo & kLineOffMask);
    const uint32_t d = p.dataoff[r];
    if ((d & kVerbatim) && (uint32_t)(lo >> 48) != 0xFFFFu) {
      s.copy(ln, (uint32_t)(lo >> 48));
      s.put('\n');
      return;
    }
    uint32_t c = 0;
    while (!sb_sep(ln[c])) c++;
    s.copy(ln, c);
    s.put('\t');
    s.put_u32(p.start[r]);
    s.put('\t');
    s.put_u32(p.end[r]);
    const uint32_t doff = d & ~kVerbatim;
    if (doff) {
      uint32_t m = 0;
      while (ln[doff + m] != '\n') m++;
      s.put('\t');
      s.copy(ln + doff, m);
    }
    s.put('\n');
  }
};

static int bits_for(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b)) b++;
  return b;
}

int sort_bed_device(bk_ctx* ctx, const char* d_text, uint64_t nbytes, int on_device, bk_text* out, uint64_t* bad_offset) {
  if (bad_offset) *bad_offset = ~0ull;
  bk_bed* bed = nullptr;
  int     rc = bk_load_bed_device(ctx, d_text, nbytes, 3, BK_COL_LINE | BK_COL_ID | BK_LOAD_SORTBED, &bed);
  if (rc != BK_OK) return rc;
  struct Guard {
    bk_ctx*            ctx;
    bk_bed*            bed;
    std::vector<void*> blocks;
    ~Guard() {
      for (void* b : blocks) dfree(ctx, b);
      bk_free_bed(ctx, bed);
    }
  } g{ctx, bed, {}};
  const uint64_t n = bed->nrows;
  if (n == 0) return finish_text(ctx, nullptr, 0, 0, on_device, out);
  if (n >= 0xFFFFFFFFull) return fail(ctx, BK_ERR_UNSUPPORTED, "more than 2^32-1 rows in one sort");

  SortParams p{};
  p.text = bed->d_text;
  p.n = n;
  p.start = bed->start;
  p.end = bed->end;
  p.line = bed->line_off;
  p.e3 = bed->idspan;
  p.scratch = ctx->d_scratch;
  p.dataoff = dalloc<uint32_t>(ctx, n);
  p.slot = dalloc<uint32_t>(ctx, n);
  p.table = dalloc<HtSlot>(ctx, HT_SIZE);
  g.blocks = {p.dataoff, p.slot, p.table};
  if (!p.dataoff || !p.slot || !p.table) return BK_ERR_NOMEM;
  const unsigned grid = (unsigned)((n + 255) / 256);

  BK_TRY(reset_scratch(ctx));
  BK_CUDA(ctx, cudaMemsetAsync(p.table, 0xFF, sizeof(HtSlot) * HT_SIZE, ctx->stream));
  prof_begin(ctx, "k_sort_validate");
  k_sort_validate<<<grid, 256, 0, ctx->stream>>>(p);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  prof_begin(ctx, "k_chrom_insert");
  k_chrom_insert<<<grid, 256, 0, ctx->stream>>>(p);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  std::vector<HtSlot> table(HT_SIZE);
  BK_CUDA(ctx, cudaMemcpyAsync(table.data(), p.table, sizeof(HtSlot) * HT_SIZE, cudaMemcpyDeviceToHost, ctx->stream));
  BK_TRY(read_scratch(ctx));
  const uint64_t* h = ctx->h_scratch;
  if (h[SC_COUNT_C] || h[SC_COUNT_D]) {  // the first offending row, whichever rule it breaks
    const uint64_t r1 = h[SC_COUNT_C] ? ~h[SC_COUNT_C] : ~0ull, r2 = h[SC_COUNT_D] ? ~h[SC_COUNT_D] : ~0ull;
    const uint64_t row = r1 < r2 ? r1 : r2;
    uint64_t       lo = 0;
    BK_CUDA(ctx, cudaMemcpyAsync(&lo, bed->line_off + row, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad_offset) *bad_offset = lo & kLineOffMask;
    if (r1 < r2) return fail(ctx, BK_ERR_PARSE, "row %llu is not a BED row sort-bed accepts", (unsigned long long)row + 1);
    return fail(ctx, BK_ERR_COORD_RANGE, "row %llu: coordinate does not fit the 32-bit device layout", (unsigned long long)row + 1);
  }
  if (h[SC_ERR_CODE]) return fail(ctx, BK_ERR_UNSUPPORTED, "more than %u distinct chromosome names", HT_SIZE / 2);

  // the names, ranked by strcmp on the host (dozens of them); the text of a representative row comes back for each
  struct Name {
    std::string s;
    uint32_t    slot;
  };
  std::vector<Name> names;
  {
    std::vector<uint32_t> used;
    for (uint32_t s = 0; s < HT_SIZE; s++)
      if (table[s].hash != HT_EMPTY) used.push_back(s);
    std::vector<uint64_t> reps(used.size()), offs(used.size());
    for (size_t k = 0; k < used.size(); k++) reps[k] = table[used[k]].rep;
    // line offsets of the representatives: one gather through the pinned scratch would do; the count is small
    for (size_t k = 0; k < used.size(); k++)
      BK_CUDA(ctx, cudaMemcpyAsync(&offs[k], bed->line_off + reps[k], 8, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<char> buf(used.size() * 129);
    for (size_t k = 0; k < used.size(); k++) {
      const uint64_t o = offs[k] & kLineOffMask;
      const uint64_t take = std::min<uint64_t>(128, bed->nbytes - o);
      BK_CUDA(ctx, cudaMemcpyAsync(&buf[k * 129], bed->d_text + o, take, cudaMemcpyDeviceToHost, ctx->stream));
      buf[k * 129 + take] = '\t';
    }
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < used.size(); k++) {
      const char* b = &buf[k * 129];
      size_t      l = 0;
      while (l < 128 && b[l] != '\t' && b[l] != ' ' && b[l] != '\n') l++;
      names.push_back({std::string(b, l), used[k]});
    }
  }
  std::sort(names.begin(), names.end(), [](const Name& a, const Name& b) { return strcmp(a.s.c_str(), b.s.c_str()) < 0; });
  for (size_t k = 1; k < names.size(); k++)
    if (names[k].s == names[k - 1].s) return fail(ctx, BK_ERR_UNSUPPORTED, "chromosome name table inconsistency ('%s')", names[k].s.c_str());
  std::vector<uint32_t> slot_rank(HT_SIZE, 0);
  for (size_t k = 0; k < names.size(); k++) slot_rank[names[k].slot] = (uint32_t)k;
  uint32_t* d_slot_rank = dalloc<uint32_t>(ctx, HT_SIZE);
  if (!d_slot_rank) return BK_ERR_NOMEM;
  g.blocks.push_back(d_slot_rank);
  BK_CUDA(ctx, cudaMemcpyAsync(d_slot_rank, slot_rank.data(), sizeof(uint32_t) * HT_SIZE, cudaMemcpyHostToDevice, ctx->stream));

  const int rank_bits = bits_for(names.size() ? names.size() - 1 : 0), start_bits = bits_for(h[SC_COUNT_A]),
            len_bits = bits_for(h[SC_COUNT_B]);
  if (rank_bits + start_bits + len_bits > 64)
    return fail(ctx, BK_ERR_UNSUPPORTED, "sort key of %d bits (chromosomes %d, start %d, length %d) exceeds 64", rank_bits + start_bits + len_bits,
                rank_bits, start_bits, len_bits);
  uint64_t* keys = dalloc<uint64_t>(ctx, n);
  uint64_t* keys2 = dalloc<uint64_t>(ctx, n);
  uint32_t* vals = dalloc<uint32_t>(ctx, n);
  uint32_t* vals2 = dalloc<uint32_t>(ctx, n);
  uint32_t* rank = dalloc<uint32_t>(ctx, n);
  g.blocks.insert(g.blocks.end(), {keys, keys2, vals, vals2, rank});
  if (!keys || !keys2 || !vals || !vals2 || !rank) return BK_ERR_NOMEM;
  BK_TRY(reset_scratch(ctx));
  prof_begin(ctx, "k_sort_keys");
  k_sort_keys<<<grid, 256, 0, ctx->stream>>>(p, d_slot_rank, start_bits, len_bits, keys, vals, rank);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  // g.blocks holds the four buffers whichever way the sort swaps them
  BK_TRY(radix_sort_pairs(ctx, &keys, &vals, &keys2, &vals2, n, rank_bits + start_bits + len_bits));
  prof_begin(ctx, "k_tie_fix");
  k_tie_fix<<<grid, 256, 0, ctx->stream>>>(p, keys, vals);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));  // also keeps slot_rank (host vector) alive until its copy has run
  if (ctx->h_scratch[SC_ERR_CODE]) return fail(ctx, BK_ERR_UNSUPPORTED, "two chromosome names share one 64-bit hash (row %llu)",
                                               (unsigned long long)ctx->h_scratch[SC_ERR_ROW] + 1);
  SortRow fn{p, vals};
  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  rc = run_emit(ctx, fn, n, 0, &d_out, &bytes, &rows);
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return rc;
  }
  return finish_text(ctx, d_out, bytes, rows, on_device, out);
}

}  // namespace bk

using namespace bk;

// the radix sort by itself on caller-owned device arrays (tests; hosts that need an order the input does not have)
extern "C" int bk_radix_sort_pairs(bk_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, uint64_t n, int nbits) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !d_keys || nbits < 1 || nbits > 64) return BK_ERR_ARG;
  ctx->last_error.clear();
  if (n < 2) return BK_OK;
  uint64_t* k2 = dalloc<uint64_t>(ctx, n);
  uint32_t* v2 = d_vals ? dalloc<uint32_t>(ctx, n) : nullptr;
  if (!k2 || (d_vals && !v2)) return BK_ERR_NOMEM;
  uint64_t *ka = d_keys, *kb = k2;
  uint32_t *va = d_vals, *vb = v2;
  int       rc = radix_sort_pairs(ctx, &ka, d_vals ? &va : nullptr, &kb, d_vals ? &vb : nullptr, n, nbits);
  if (rc == BK_OK && ka != d_keys) {  // an odd number of passes left the result in the scratch pair
    if (cudaMemcpyAsync(d_keys, ka, n * 8, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) rc = BK_ERR_CUDA;
    if (d_vals && cudaMemcpyAsync(d_vals, va, n * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) rc = BK_ERR_CUDA;
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = BK_ERR_CUDA;
  dfree(ctx, k2);
  dfree(ctx, v2);
  return rc;
}

extern "C" int bk_sort_bed_device(bk_ctx* ctx, const char* dev_text, size_t nbytes, int out_on_device, bk_text* out,
                                  uint64_t* bad_offset) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !out || (!dev_text && nbytes)) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  return sort_bed_device(ctx, dev_text, nbytes, out_on_device, out, bad_offset);
}

extern "C" int bk_sort_bed(bk_ctx* ctx, const char* host_text, size_t nbytes, int out_on_device, bk_text* out, uint64_t* bad_offset) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !out || (!host_text && nbytes)) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  char* d = reinterpret_cast<char*>(dmalloc(ctx, nbytes + 64));
  if (!d) return BK_ERR_NOMEM;
  int rc = upload(ctx, d, host_text, nbytes, ctx->stream);
  if (rc == BK_OK) rc = sort_bed_device(ctx, d, nbytes, out_on_device, out, bad_offset);
  dfree(ctx, d);
  return rc;
}
