// pad.cu -- bedops --range L:R: the padded view of a parsed BED file (BedPadReader.hpp:116-277).
//
// Every start moves by L, every end by R.  What makes this more than two additions is what the reference's reader does
// around zero and with rows that stop being intervals:
//   * rpad < 0 or lpad > 0 (:133-141): a row whose padded end does not pass its padded start vaporises; a start that
//     would go below zero wraps (unsigned) and the row vaporises too -- except in the zone the constructor's getFirst()
//     consumed when lpad < 0 (:194-277): from the first row of the file up to the first surviving row with
//     start > |lpad|, where starts are clamped to 0 and the rows re-sorted;
//   * lpad < 0 otherwise (:142-155): at every chromosome the starts <= |lpad| are clamped to 0 and those rows re-ordered
//     by their end (input order on ties); the other rows shift;
//   * rpad > 0 (:156-160): ends shift.
// Device form: one pass computes the padded coordinates and a keep / clamped flag per row; two flag scans give every
// kept row its slot; the clamped rows -- a prefix of their chromosome -- are ordered by (chromosome, end) with the radix
// sort of sort.cu (stable) and dealt into the slots the clamped rows occupy; a gather builds the new columns.  Echoed
// rows re-print their numbers (line length marker 0xFFFF).
#include <algorithm>
#include "common.cuh"
#include "parse.cuh"
#include "sort.cuh"

namespace bk {

constexpr int FR_RANGE = 2048;  // rows per warp range of the flag scans

struct PadParams {
  const uint32_t* s;
  const uint32_t* e;
  uint64_t        n;
  long long       lpad, rpad;
  uint64_t        zone_end;   // mode A with lpad < 0: index of the row that ends getFirst()'s zone (n: the whole file)
  const uint64_t* run_begin;  // [nruns]
  int             nruns;
  uint32_t*       ns;
  uint32_t*       ne;
  uint8_t*        flags;      // bit 0 keep, bit 1 kept and clamped to start 0
  uint64_t*       scratch;
};

// first row with start > |lpad| that survives the padding (ends getFirst()'s loop, BedPadReader.hpp:206-213);
// the end test is unsigned there: an end below |rpad| wraps and passes
__global__ void __launch_bounds__(256) k_pad_zone(PadParams p) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const long long lpd = -p.lpad, s = p.s[i], e = p.e[i];
  if (s > lpd && (e + p.rpad < 0 || e + p.rpad > s - lpd))
    atomicMax(reinterpret_cast<unsigned long long*>(&p.scratch[SC_COUNT_A]), ~(unsigned long long)i);
}

__global__ void __launch_bounds__(256) k_pad_flags(PadParams p) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const long long s = p.s[i], e = p.e[i], lpd = p.lpad < 0 ? -p.lpad : p.lpad;
  long long       ns = s, ne = e + p.rpad;
  bool            keep = true, clamped = false;
  if (p.rpad < 0 || p.lpad > 0) {
    if (p.lpad < 0 && i <= p.zone_end) {
      if (s > lpd) {
        ns = s - lpd;
        keep = i == p.zone_end;  // the rows before it with start > |lpad| did not survive
      } else {
        keep = ne > 0;
        ns = 0;
        clamped = keep;
      }
    } else {
      ns = s + p.lpad;
      keep = ns >= 0 && ne > ns;  // a negative start wraps to a huge unsigned value in the reference: never an interval
    }
  } else if (p.lpad < 0) {
    if (s <= lpd) {
      ns = 0;
      clamped = true;
    } else {
      ns = s - lpd;
    }
  }
  if (keep && (ne < 0 || ne >= 0xFFFFFFFFll || ns >= 0xFFFFFFFFll)) dev_set_error(p.scratch, BK_ERR_COORD_RANGE, i);
  p.ns[i] = (uint32_t)ns;
  p.ne[i] = (uint32_t)ne;
  p.flags[i] = (keep ? 1 : 0) | (clamped ? 2 : 0);
}

// exclusive rank of the rows whose flag has `bit`: warp ranges of FR_RANGE rows -> totals -> k_scan_totals -> ranks
__global__ void __launch_bounds__(256) k_flag_totals(const uint8_t* __restrict__ flags, uint64_t n, uint8_t bit, uint64_t* __restrict__ tot) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * FR_RANGE, b = a + FR_RANGE < n ? a + FR_RANGE : n;
  if (a >= n) return;
  uint32_t c = 0;
  for (uint64_t k = a + lane; k < b; k += 32) c += (flags[k] & bit) ? 1u : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) tot[w] = c;
}
__global__ void __launch_bounds__(256) k_flag_ranks(const uint8_t* __restrict__ flags, uint64_t n, uint8_t bit,
                                                    const uint64_t* __restrict__ base, uint32_t* __restrict__ rank) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * FR_RANGE, b = a + FR_RANGE < n ? a + FR_RANGE : n;
  if (a >= n) return;
  uint32_t run = (uint32_t)base[w];
  for (uint64_t k0 = a; k0 < b; k0 += 32) {
    const uint64_t k = k0 + lane;
    const bool     f = k < b && (flags[k] & bit);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (k < b) rank[k] = run + (uint32_t)__popc(m & ((1u << lane) - 1u));
    run += (uint32_t)__popc(m);
  }
}

struct PermParams {
  const uint8_t*  flags;
  const uint32_t* keep_rank;
  const uint32_t* clamp_rank;
  const uint32_t* ne;
  const uint64_t* run_begin;
  int             nruns;
  uint64_t        n;
  uint32_t*       perm;   // [kept] source row of every output slot
  uint64_t*       ckey;   // [clamped] (run << 32) | padded end
  uint32_t*       crow;   // [clamped] source row
  uint32_t*       cslot;  // [clamped] output slot of the j-th clamped row in file order
};
__global__ void __launch_bounds__(256) k_pad_perm(PermParams p) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const uint8_t f = p.flags[i];
  if (!(f & 1)) return;
  if (!(f & 2)) {
    p.perm[p.keep_rank[i]] = (uint32_t)i;
    return;
  }
  int l = 0, h = p.nruns;
  while (h - l > 1) {
    const int mid = (l + h) >> 1;
    if (p.run_begin[mid] <= i) l = mid; else h = mid;
  }
  const uint32_t j = p.clamp_rank[i];
  p.ckey[j] = ((uint64_t)l << 32) | p.ne[i];
  p.crow[j] = (uint32_t)i;
  p.cslot[j] = p.keep_rank[i];
}
__global__ void __launch_bounds__(256) k_pad_deal(const uint32_t* __restrict__ crow_sorted, const uint32_t* __restrict__ cslot, uint64_t nc,
                                                  uint32_t* __restrict__ perm) {
  const uint64_t j = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (j < nc) perm[cslot[j]] = crow_sorted[j];
}

struct GatherParams {
  const uint32_t* perm;
  uint64_t        nk;
  const uint32_t *ns, *ne;
  const uint64_t* line;
  const double*   score;
  const uint32_t* idspan;
  uint32_t *      os, *oe;
  uint64_t*       oline;
  double*         oscore;
  uint32_t*       oidspan;
  uint64_t        end_sentinel;
};
__global__ void __launch_bounds__(256) k_pad_gather(GatherParams g) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i == 0 && g.oline) g.oline[g.nk] = g.end_sentinel;
  if (i >= g.nk) return;
  const uint32_t r = g.perm[i];
  g.os[i] = g.ns[r];
  g.oe[i] = g.ne[r];
  if (g.oline) g.oline[i] = (g.line[r] & kLineOffMask) | (0xFFFFull << 48);  // the numbers changed: echo re-prints them
  if (g.oscore) g.oscore[i] = g.score[r];
  if (g.oidspan) g.oidspan[i] = g.idspan[r];
}

// the rest of a line after its third field, and strcmp of two '\n'-terminated strings (as setops.cu's --everything)
__device__ __forceinline__ const char* pad_rest_of_line(const char* p) {
  while (is_tok((unsigned char)*p)) p++;
  for (int f = 0; f < 2; f++) {
    while (is_ws((unsigned char)*p)) p++;
    if (*p == '+') p++;
    while (is_digit((unsigned char)*p)) p++;
  }
  return p;
}
// rows that became equal in (start, end) keep their input order; --everything over several files merges by the rest of
// the line too and needs it ascending inside such a group: flag a file where it is not
__global__ void __launch_bounds__(256) k_pad_tie_order(const uint32_t* __restrict__ s, const uint32_t* __restrict__ e,
                                                       const uint64_t* __restrict__ line, const char* __restrict__ text, uint64_t n,
                                                       uint64_t* scratch) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x + 1;
  if (i >= n || s[i] != 0 || s[i - 1] != 0 || e[i] != e[i - 1]) return;
  const char* a = pad_rest_of_line(text + (line[i - 1] & kLineOffMask));
  const char* b = pad_rest_of_line(text + (line[i] & kLineOffMask));
  while (true) {
    const unsigned char x = *a == '\n' ? 0 : (unsigned char)*a, y = *b == '\n' ? 0 : (unsigned char)*b;
    if (x != y) {
      if (x > y) scratch[SC_COUNT_D] = 1;
      return;
    }
    if (x == 0) return;
    a++;
    b++;
  }
}

static int flag_ranks(bk_ctx* ctx, const uint8_t* flags, uint64_t n, uint8_t bit, uint32_t* rank, uint64_t* total) {
  const uint32_t nw = (uint32_t)((n + FR_RANGE - 1) / FR_RANGE);
  uint64_t*      tot = dalloc<uint64_t>(ctx, nw);
  uint64_t*      base = dalloc<uint64_t>(ctx, (size_t)nw + 1);
  if (!tot || !base) return BK_ERR_NOMEM;
  const unsigned grid = (unsigned)(((uint64_t)nw * 32 + 255) / 256);
  k_flag_totals<<<grid, 256, 0, ctx->stream>>>(flags, n, bit, tot);
  BK_LAUNCHED(ctx);
  k_scan_totals<SC_COUNT_B><<<1, 1024, 0, ctx->stream>>>(tot, base, nw, ctx->d_scratch);
  BK_LAUNCHED(ctx);
  k_flag_ranks<<<grid, 256, 0, ctx->stream>>>(flags, n, bit, base, rank);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));
  *total = ctx->h_scratch[SC_COUNT_B];
  dfree(ctx, tot);
  dfree(ctx, base);
  return BK_OK;
}

}  // namespace bk

using namespace bk;

extern "C" int bk_bed_pad(bk_ctx* ctx, const bk_bed* src, long long lpad, long long rpad, bk_bed** out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !src || !out) return BK_ERR_ARG;
  *out = nullptr;
  ctx->last_error.clear();
  const uint64_t n = src->nrows;
  if (n >= 0xFFFFFFFFull) return fail(ctx, BK_ERR_UNSUPPORTED, "--range over more than 2^32-1 rows");
  if (lpad > 0x7FFFFFFFll || lpad < -0x7FFFFFFFll || rpad > 0x7FFFFFFFll || rpad < -0x7FFFFFFFll) return fail(ctx, BK_ERR_ARG, "--range value outside int");
  bk_bed* b = new bk_bed();
  b->min_fields = src->min_fields;
  b->cols = src->cols;
  b->d_text = src->d_text;  // borrowed: the source file must outlive its padded view
  b->owns_text = false;
  b->nbytes = src->nbytes;
  std::vector<void*> tmp;
  auto done = [&](int rc) {
    for (void* q : tmp) dfree(ctx, q);
    if (rc != BK_OK) bk_free_bed(ctx, b);
    else *out = b;
    return rc;
  };
  if (n == 0) {
    b->runs = src->runs;
    return done(BK_OK);
  }
  std::vector<uint64_t> rb;
  for (auto& r : src->runs) rb.push_back(r.row_begin);
  if (rb.empty()) rb.push_back(0);
  PadParams p{};
  p.s = src->start; p.e = src->end; p.n = n; p.lpad = lpad; p.rpad = rpad; p.zone_end = n; p.nruns = (int)rb.size();
  p.scratch = ctx->d_scratch;
  uint64_t* d_rb = dalloc<uint64_t>(ctx, rb.size());
  p.ns = dalloc<uint32_t>(ctx, n); p.ne = dalloc<uint32_t>(ctx, n); p.flags = dalloc<uint8_t>(ctx, n);
  uint32_t* keep_rank = dalloc<uint32_t>(ctx, n);
  uint32_t* clamp_rank = dalloc<uint32_t>(ctx, n);
  tmp = {d_rb, p.ns, p.ne, p.flags, keep_rank, clamp_rank};
  if (!d_rb || !p.ns || !p.ne || !p.flags || !keep_rank || !clamp_rank) return done(BK_ERR_NOMEM);
  p.run_begin = d_rb;
  const unsigned grid = (unsigned)((n + 255) / 256);
  int            rc = reset_scratch(ctx);
  if (rc != BK_OK) return done(rc);
  if (cudaMemcpyAsync(d_rb, rb.data(), rb.size() * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return done(BK_ERR_CUDA);
  if ((rpad < 0 || lpad > 0) && lpad < 0) {
    k_pad_zone<<<grid, 256, 0, ctx->stream>>>(p);
    if ((rc = read_scratch(ctx)) != BK_OK) return done(rc);
    if (ctx->h_scratch[SC_COUNT_A]) p.zone_end = ~ctx->h_scratch[SC_COUNT_A];
    if ((rc = reset_scratch(ctx)) != BK_OK) return done(rc);
  }
  prof_begin(ctx, "k_pad_flags");
  k_pad_flags<<<grid, 256, 0, ctx->stream>>>(p);
  prof_end(ctx);
  ctx->launches++;
  uint64_t nk = 0, nc = 0;
  if ((rc = flag_ranks(ctx, p.flags, n, 1, keep_rank, &nk)) != BK_OK) return done(rc);
  if (ctx->h_scratch[SC_ERR_CODE])
    return done(fail(ctx, BK_ERR_COORD_RANGE, "--range moves a coordinate of row %llu outside the 32-bit device layout (the reference wraps an end below |R| there)",
                     (unsigned long long)ctx->h_scratch[SC_ERR_ROW] + 1));
  if ((rc = flag_ranks(ctx, p.flags, n, 2, clamp_rank, &nc)) != BK_OK) return done(rc);
  uint32_t* perm = dalloc<uint32_t>(ctx, nk);
  tmp.push_back(perm);
  if (!perm) return done(BK_ERR_NOMEM);
  PermParams q{};
  q.flags = p.flags; q.keep_rank = keep_rank; q.clamp_rank = clamp_rank; q.ne = p.ne; q.run_begin = d_rb; q.nruns = p.nruns; q.n = n; q.perm = perm;
  uint64_t *ckey = nullptr, *ckey2 = nullptr;
  uint32_t *crow = nullptr, *crow2 = nullptr, *cslot = nullptr;
  if (nc) {
    ckey = dalloc<uint64_t>(ctx, nc); ckey2 = dalloc<uint64_t>(ctx, nc);
    crow = dalloc<uint32_t>(ctx, nc); crow2 = dalloc<uint32_t>(ctx, nc); cslot = dalloc<uint32_t>(ctx, nc);
    tmp.insert(tmp.end(), {ckey, ckey2, crow, crow2, cslot});
    if (!ckey || !ckey2 || !crow || !crow2 || !cslot) return done(BK_ERR_NOMEM);
  }
  q.ckey = ckey; q.crow = crow; q.cslot = cslot;
  k_pad_perm<<<grid, 256, 0, ctx->stream>>>(q);
  ctx->launches++;
  if (nc) {
    int run_bits = 1;
    while ((rb.size() - 1) >> run_bits) run_bits++;
    if ((rc = radix_sort_pairs(ctx, &ckey, &crow, &ckey2, &crow2, nc, 32 + run_bits)) != BK_OK) return done(rc);
    k_pad_deal<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(crow, cslot, nc, perm);
    ctx->launches++;
  }
  b->nrows = nk;
  b->start = dalloc<uint32_t>(ctx, nk);
  b->end = dalloc<uint32_t>(ctx, nk);
  if (src->line_off) b->line_off = dalloc<uint64_t>(ctx, nk + 1);
  if (src->score) b->score = dalloc<double>(ctx, nk);
  if (src->idspan) b->idspan = dalloc<uint32_t>(ctx, nk);
  if (!b->start || !b->end || (src->line_off && !b->line_off) || (src->score && !b->score) || (src->idspan && !b->idspan)) return done(BK_ERR_NOMEM);
  GatherParams g{};
  g.perm = perm; g.nk = nk; g.ns = p.ns; g.ne = p.ne; g.line = src->line_off; g.score = src->score; g.idspan = src->idspan;
  g.os = b->start; g.oe = b->end; g.oline = b->line_off; g.oscore = b->score; g.oidspan = b->idspan; g.end_sentinel = src->nbytes;
  k_pad_gather<<<(unsigned)((nk + 256) / 256), 256, 0, ctx->stream>>>(g);
  ctx->launches++;
  // rows per chromosome after the padding: the keep rank at every run border
  std::vector<uint32_t> at(src->runs.size() + 1, (uint32_t)nk);
  for (size_t k = 0; k < src->runs.size(); k++)
    if (src->runs[k].row_begin < n &&
        cudaMemcpyAsync(&at[k], keep_rank + src->runs[k].row_begin, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      return done(BK_ERR_CUDA);
  if ((rc = reset_scratch(ctx)) != BK_OK) return done(rc);
  if (b->line_off && nk > 1 && nc > 1) {
    k_pad_tie_order<<<(unsigned)((nk + 255) / 256), 256, 0, ctx->stream>>>(b->start, b->end, b->line_off, b->d_text, nk, ctx->d_scratch);
    ctx->launches++;
  }
  if ((rc = read_scratch(ctx)) != BK_OK) return done(rc);  // syncs: `at`, rb are filled / consumed
  b->pad_tie_disorder = ctx->h_scratch[SC_COUNT_D] != 0;
  for (size_t k = 0; k < src->runs.size(); k++) b->runs.push_back({src->runs[k].name, at[k], k + 1 < src->runs.size() ? at[k + 1] : nk});
  return done(BK_OK);
}
