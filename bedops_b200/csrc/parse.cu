// parse.cu -- the BED reader as one single-pass sm_100a kernel.
//
// Replaces Bed::allocate_iterator_starch_bed<T*>::operator++ -> T::readline(FILE*) with fscanf formats
// "%s\t%lu\t%lu%[^\n]s\n" (B3Rest, Bed.hpp:380-382), "...\t%s%[^\n]s\n" (B4Rest, :644-646) and
// "...\t%s\t%lf%[^\n]s\n" (B5Rest, :901-903).
//
// Design (B200): a persistent grid (multiple of 148 CTAs) takes 8 KiB text tiles by dynamic ticket.  A tile is
// staged into shared memory with 16-byte coalesced streaming loads (plus a small halo either side), newlines
// are found with SWAR compares, the number of rows that START in the tile is published through a decoupled
// look-back chain (one 64-bit word per tile) to obtain the global row index, and then one thread per line
// tokenises its line out of shared memory and writes the SoA columns with coalesced 4/8-byte stores.
// Text is read from HBM exactly once; nothing but the SoA columns is written.
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

// compare the chromosome token at window index a (length la) with the token starting at window index b
template <class C>
__device__ __forceinline__ bool same_token(const C& c, int64_t a, int la, int64_t b) {
  for (int i = 0; i < la; i++)
    if (c.at(a + i) != c.at(b + i)) return false;
  return !is_tok(c.at(b + la));
}

// window index of the first token of the last non-blank line that ends before window index p0 (p0 = a line start);
// INT64_MIN if there is none.  Rare path (first row of a tile only): walks backwards through the window/global text.
__device__ int64_t prev_line_token(const Cursor& c, int64_t p0) {
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

__global__ void __launch_bounds__(P_THREADS) k_parse(ParseParams p) {
  __shared__ __align__(16) unsigned char sm[P_BUF];
  __shared__ uint16_t                    lstart[P_MAXROWS];
  __shared__ uint32_t                    scan_sm[34];
  __shared__ uint32_t                    ticket_sm;
  __shared__ uint64_t                    base_sm;

  const int      tid = threadIdx.x;
  const uint64_t eff = p.scratch[SC_EFFLEN];  // bytes up to and including the last '\n'
  const unsigned char* text = reinterpret_cast<const unsigned char*>(p.text);

  while (true) {
    const uint32_t tile = next_ticket(p.scratch, &ticket_sm);
    if (tile >= p.ntiles) break;
    const int64_t ts = (int64_t)tile * P_TILE;
    const int64_t g0 = ts - P_PRE;

    // ---- stage [ts-PRE, ts+TILE+POST) into shared memory, 16 bytes per request ---------------------------
    for (int v = tid; v < P_BUF / 16; v += P_THREADS) {
      int64_t g = g0 + (int64_t)v * 16;
      uint4   w = make_uint4(0, 0, 0, 0);
      if (g >= 0 && (uint64_t)g + 16 <= p.nbytes_raw) {
        w = ldg_stream16(text + g);
      } else if (g + 16 > 0 && (uint64_t)(g < 0 ? 0 : g) < p.nbytes_raw) {
        unsigned char b[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          int64_t gi = g + i;
          b[i] = (gi >= 0 && (uint64_t)gi < p.nbytes_raw) ? text[gi] : 0;
        }
        w.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
        w.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
        w.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
        w.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
      }
      reinterpret_cast<uint4*>(sm)[v] = w;
    }
    __syncthreads();

    // ---- line starts: position q starts a line iff q == 0 or byte q-1 is '\n' -----------------------------
    Cursor cur{sm, g0, text, eff};
    const int      off = P_PRE + tid * 32;
    const uint32_t* wv = reinterpret_cast<const uint32_t*>(sm + off);
    uint32_t       nl = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      uint32_t m = __vcmpeq4(wv[w], 0x0A0A0A0Au) & 0x01010101u;
      nl |= ((m * 0x01020408u) >> 24) << (4 * w);
    }
    const int64_t p0 = ts + tid * 32;  // global offset of this thread's first byte
    uint32_t      prevnl = (p0 == 0) ? 1u : (sm[off - 1] == '\n');
    uint32_t      smask = (nl << 1) | prevnl;
    if ((uint64_t)p0 >= eff) smask = 0;
    else if ((uint64_t)p0 + 32 > eff) smask &= (1u << (int)(eff - (uint64_t)p0)) - 1u;
    // drop blank lines (fscanf's %s skips them: whitespace, including '\n', is not a record)
    for (uint32_t m = smask; m; m &= m - 1) {
      int     j = __ffs(m) - 1;
      int64_t q = off + j;
      while (is_ws(cur.at(q))) q++;
      if (cur.at(q) == '\n') smask &= ~(1u << j);
    }
    uint32_t cnt = __popc(smask), nrow;
    uint32_t ex = block_excl_scan(cnt, scan_sm, &nrow);
    for (uint32_t m = smask; m; m &= m - 1) lstart[ex++] = (uint16_t)(off + __ffs(m) - 1);

    // ---- global row index of the tile's first row -----------------------------------------------------------
    if (tid < 32) {
      uint64_t b = lookback_sum(p.tile_state, tile, nrow);
      if (tid == 0) base_sm = b;
    }
    __syncthreads();
    const uint64_t base = base_sm;

    // ---- one thread per line -----------------------------------------------------------------------------
    for (uint32_t k = tid; k < nrow; k += P_THREADS) {
      const int64_t q0 = lstart[k];
      RowOut        r;
      bool          head;
      if (k + 1 < nrow) {  // line ends before the next line start: entirely inside the window
        SmCursor sc{sm};
        parse_line(sc, q0, p.min_fields, p.cols, r);
      } else {
        parse_line(cur, q0, p.min_fields, p.cols, r);
      }
      const uint64_t row = base + k;
      if (r.err) {
        dev_set_error(p.scratch, r.err, row);
        continue;
      }
      if (k > 0) {  // line k-1 lies wholly inside the window
        int64_t pq = lstart[k - 1];
        while (is_ws(sm[pq])) pq++;
        head = !same_token(cur, r.tok0, r.toklen, pq);
      } else {
        int64_t pq = prev_line_token(cur, q0);
        head = (pq == INT64_MIN) || !same_token(cur, r.tok0, r.toklen, pq);
      }
      if (head) {
        uint32_t h = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(&p.scratch[SC_NHEADS]), 1ull);
        if (h < p.heads_cap) {
          HeadRec* hr = &p.heads[h];
          hr->row = row;
          hr->len = r.toklen;
          for (int i = 0; i < r.toklen; i++) hr->name[i] = cur.at(r.tok0 + i);
          hr->name[r.toklen] = 0;
        }
      }
      if (row < p.cap) {
        p.start[row] = (uint32_t)r.start;
        p.end[row] = (uint32_t)r.end;
        if (p.score) p.score[row] = r.score;
        if (p.line_off) p.line_off[row] = (uint64_t)(g0 + r.tok0);
        if (p.idspan) p.idspan[row] = ((uint32_t)(r.id0 - r.tok0) << 16) | (uint32_t)r.idlen;
      }
    }
    if (tile == p.ntiles - 1 && tid == 0) p.scratch[SC_NROWS] = base + nrow;
    __syncthreads();
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

// count '\n' in a prefix of the text (row-capacity estimate)
__global__ void k_count_nl(const unsigned char* text, uint64_t n, uint64_t* scratch) {
  uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  uint32_t c = 0;
  if (i + 16 <= n) {
    uint4 w = ldg_stream16(text + i);
    c = __popc(__vcmpeq4(w.x, 0x0A0A0A0Au) & 0x01010101u) + __popc(__vcmpeq4(w.y, 0x0A0A0A0Au) & 0x01010101u) +
        __popc(__vcmpeq4(w.z, 0x0A0A0A0Au) & 0x01010101u) + __popc(__vcmpeq4(w.w, 0x0A0A0A0Au) & 0x01010101u);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(&scratch[SC_COUNT_A]), (unsigned long long)c);
}

// inclusive running max of end within each chromosome run (the index that bounds candidate windows for nested
// intervals, north_star item 3).  Tiles of 256*8 rows; segmented-max look-back across tiles.
constexpr int PM_THREADS = 256, PM_ITEMS = 8, PM_TILE = PM_THREADS * PM_ITEMS;
__global__ void __launch_bounds__(PM_THREADS) k_pmax(const uint32_t* __restrict__ end, uint32_t* __restrict__ pmax,
                                                     uint64_t n, const uint64_t* __restrict__ run_begin, int nruns,
                                                     uint64_t* tile_state, uint32_t ntiles, uint64_t* scratch) {
  __shared__ uint32_t ticket_sm;
  __shared__ uint64_t wmax[PM_THREADS / 32];
  __shared__ uint32_t whead[PM_THREADS / 32];
  __shared__ uint64_t carry_sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  while (true) {
    const uint32_t tile = next_ticket(scratch, &ticket_sm);
    if (tile >= ntiles) break;
    const uint64_t r0 = (uint64_t)tile * PM_TILE + (uint64_t)tid * PM_ITEMS;
    // which rows of this thread's 8 are run heads?  run_begin is sorted; find the first run_begin >= r0
    int lo = 0, hi = nruns;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (run_begin[mid] < r0) lo = mid + 1; else hi = mid;
    }
    uint32_t v[PM_ITEMS];
    uint32_t headmask = 0;
#pragma unroll
    for (int i = 0; i < PM_ITEMS; i++) {
      uint64_t r = r0 + i;
      v[i] = r < n ? end[r] : 0;
      while (lo < nruns && run_begin[lo] < r) lo++;
      if (lo < nruns && run_begin[lo] == r) headmask |= 1u << i;
    }
    // thread-local segmented inclusive max
    uint32_t run = 0;
    bool     seen = false;
#pragma unroll
    for (int i = 0; i < PM_ITEMS; i++) {
      if (headmask & (1u << i)) { run = 0; seen = true; }
      run = v[i] > run ? v[i] : run;
      v[i] = run;
    }
    // warp-level segmented scan of (seen, run): value flowing into each thread
    uint32_t agg = run;
    bool     aggh = seen;
    uint32_t inflow = 0;  // max flowing into this thread from earlier threads of the warp (until a head)
    {
      uint32_t a = agg;
      bool     h = aggh;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t oa = __shfl_up_sync(0xffffffffu, a, d);
        bool     oh = __shfl_up_sync(0xffffffffu, (int)h, d);
        if (lane >= d) {
          if (!h) a = oa > a ? oa : a;
          h = h || oh;
        }
      }
      // a,h = inclusive segmented aggregate up to this thread; exclusive = previous lane's
      uint32_t pa = __shfl_up_sync(0xffffffffu, a, 1);
      bool     ph = __shfl_up_sync(0xffffffffu, (int)h, 1);
      inflow = lane ? pa : 0;
      bool inflow_h = lane ? ph : false;
      if (lane == 31) { wmax[warp] = a; whead[warp] = h; }
      __syncthreads();
      // cross-warp: combine earlier warps (right to left until a head)
      uint32_t winflow = 0;
      bool     wh = false;
      for (int w = warp - 1; w >= 0 && !wh; w--) {
        winflow = wmax[w] > winflow ? (uint32_t)wmax[w] : winflow;
        wh = whead[w];
      }
      // tile-level look-back (warp 0 publishes the tile aggregate)
      if (warp == 0) {
        uint32_t ta = 0;
        bool     th = false;
        for (int w = PM_THREADS / 32 - 1; w >= 0 && !th; w--) {
          ta = wmax[w] > ta ? (uint32_t)wmax[w] : ta;
          th = whead[w];
        }
        uint64_t c = lookback_segmax(tile_state, tile, th, ta);
        if (lane == 0) carry_sm = c;
      }
      __syncthreads();
      uint32_t carry = (uint32_t)carry_sm;
      // total inflow for this thread
      if (!inflow_h) {
        inflow = winflow > inflow ? winflow : inflow;
        if (!wh) inflow = carry > inflow ? carry : inflow;
      }
    }
    // apply inflow to the items before this thread's first head
#pragma unroll
    for (int i = 0; i < PM_ITEMS; i++) {
      if (headmask & (1u << i)) inflow = 0;
      uint32_t o = v[i] > inflow ? v[i] : inflow;
      if (r0 + i < n) pmax[r0 + i] = o;
    }
    __syncthreads();
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

static int grid_for(const void* kernel, int threads, uint32_t ntiles) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
  if (per_sm < 1) per_sm = 1;
  uint64_t g = (uint64_t)kSMs * per_sm;
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
  // row-capacity estimate: exact bound for small inputs, sampled line length for large ones
  uint64_t cap;
  if (nbytes_raw <= (64ull << 20)) {
    cap = nbytes_raw / 6 + 2;  // shortest legal line "c\t0\t1\n"
  } else {
    uint64_t sample = 16ull << 20;
    prof_begin(ctx, "k_count_nl");
    k_count_nl<<<(unsigned)(sample / 16 / 256), 256, 0, ctx->stream>>>(text, sample, ctx->d_scratch);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    BK_TRY(read_scratch(ctx));
    uint64_t nl = ctx->h_scratch[SC_COUNT_A];
    if (nl < 16) nl = 16;
    double per = (double)sample / (double)nl;
    cap = (uint64_t)((double)nbytes_raw / per * 1.10) + 4096;
    if (cap > nbytes_raw / 6 + 2) cap = nbytes_raw / 6 + 2;
    BK_TRY(reset_scratch(ctx));
  }
  const uint32_t heads_cap = 1u << 16;
  HeadRec*       d_heads = dalloc<HeadRec>(ctx, heads_cap);
  if (!d_heads) return BK_ERR_NOMEM;

  for (int attempt = 0; attempt < 2; attempt++) {
    bed->start = dalloc<uint32_t>(ctx, cap);
    bed->end = dalloc<uint32_t>(ctx, cap);
    if (bed->cols & BK_COL_SCORE) bed->score = dalloc<double>(ctx, cap);
    if (bed->cols & BK_COL_LINE) bed->line_off = dalloc<uint64_t>(ctx, cap + 1);
    if (bed->cols & BK_COL_ID) bed->idspan = dalloc<uint32_t>(ctx, cap);
    if (!bed->start || !bed->end || ((bed->cols & BK_COL_SCORE) && !bed->score) ||
        ((bed->cols & BK_COL_LINE) && !bed->line_off) || ((bed->cols & BK_COL_ID) && !bed->idspan))
      return BK_ERR_NOMEM;

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
    p.ntiles = (uint32_t)((nbytes_raw + P_TILE - 1) / P_TILE);
    p.scratch = ctx->d_scratch;
    p.heads = d_heads;
    p.heads_cap = heads_cap;
    p.tile_state = dalloc<uint64_t>(ctx, p.ntiles);
    if (!p.tile_state) return BK_ERR_NOMEM;
    BK_CUDA(ctx, cudaMemsetAsync(p.tile_state, 0, (size_t)p.ntiles * 8, ctx->stream));
    prof_begin(ctx, "k_efflen");
    k_efflen<<<1, 32, 0, ctx->stream>>>(text, nbytes_raw, ctx->d_scratch);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    prof_begin(ctx, "k_parse");
    k_parse<<<grid_for((const void*)k_parse, P_THREADS, p.ntiles), P_THREADS, 0, ctx->stream>>>(p);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    BK_TRY(read_scratch(ctx));
    dfree(ctx, p.tile_state);
    const uint64_t* h = ctx->h_scratch;
    if (h[SC_ERR_CODE]) {
      int code = (int)h[SC_ERR_CODE];
      dfree(ctx, d_heads);
      const char* what = code == BK_ERR_COORD_RANGE ? "coordinate does not fit the 32-bit device layout"
                         : code == BK_ERR_UNSUPPORTED ? "score literal outside the exact device strtod path"
                                                      : "line is not chrom<ws>start<ws>end[...] with enough fields";
      return fail(ctx, code, "BED parse error at row %llu: %s", (unsigned long long)h[SC_ERR_ROW] + 1, what);
    }
    bed->nrows = h[SC_NROWS];
    bed->nbytes = h[SC_EFFLEN];
    if (bed->nrows <= cap) break;
    // the estimate was too small: the exact row count is now known; free and redo once
    dfree(ctx, bed->start); dfree(ctx, bed->end); dfree(ctx, bed->score); dfree(ctx, bed->line_off); dfree(ctx, bed->idspan);
    bed->start = bed->end = nullptr; bed->score = nullptr; bed->line_off = nullptr; bed->idspan = nullptr;
    cap = bed->nrows;
    BK_TRY(reset_scratch(ctx));
  }
  if (bed->line_off && bed->nrows) {
    uint64_t endoff = bed->nbytes;
    BK_CUDA(ctx, cudaMemcpyAsync(bed->line_off + bed->nrows, &endoff, 8, cudaMemcpyHostToDevice, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  // chromosome runs
  uint64_t nheads = ctx->h_scratch[SC_NHEADS];
  if (nheads > heads_cap) {
    dfree(ctx, d_heads);
    return fail(ctx, BK_ERR_UNSUPPORTED, "more than %u chromosome runs in one file", heads_cap);
  }
  std::vector<HeadRec> heads(nheads);
  if (nheads) {
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

int ensure_pmax(bk_ctx* ctx, const bk_bed* cbed) {
  bk_bed* bed = const_cast<bk_bed*>(cbed);
  if (bed->pmax_end || bed->nrows == 0) return BK_OK;
  bed->pmax_end = dalloc<uint32_t>(ctx, bed->nrows);
  if (!bed->pmax_end) return BK_ERR_NOMEM;
  int                   nruns = (int)bed->runs.size();
  std::vector<uint64_t> rb(nruns);
  for (int i = 0; i < nruns; i++) rb[i] = bed->runs[i].row_begin;
  uint64_t* d_rb = dalloc<uint64_t>(ctx, nruns);
  uint32_t  ntiles = (uint32_t)((bed->nrows + PM_TILE - 1) / PM_TILE);
  uint64_t* state = dalloc<uint64_t>(ctx, ntiles);
  if (!d_rb || !state) return BK_ERR_NOMEM;
  BK_CUDA(ctx, cudaMemcpyAsync(d_rb, rb.data(), nruns * 8, cudaMemcpyHostToDevice, ctx->stream));
  BK_CUDA(ctx, cudaMemsetAsync(state, 0, (size_t)ntiles * 8, ctx->stream));
  BK_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch + SC_TICKET, 0, 8, ctx->stream));
  prof_begin(ctx, "k_pmax");
  k_pmax<<<grid_for((const void*)k_pmax, PM_THREADS, ntiles), PM_THREADS, 0, ctx->stream>>>(
      bed->end, bed->pmax_end, bed->nrows, d_rb, nruns, state, ntiles, ctx->d_scratch);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // rb (host vector) must outlive the copy
  dfree(ctx, d_rb);
  dfree(ctx, state);
  return BK_OK;
}

}  // namespace bk
