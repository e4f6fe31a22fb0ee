// closest.cu -- closest-features: nearest left / right query element per reference row.
//
// Replaces the streaming two-pointer state machine findDistances (applications/bed/closestfeats/src/
// ClosestFeature.cpp:260-413) by its declarative rule, evaluated per reference row by one warp over the
// candidate window of the sorted query columns (the same [lo,hi) window the bedmap kernel scans):
//   non-overlapping left  = the query row with the largest end <= ref.start, the LATER row on ties (":300  >=")
//                           = max over { prefix [0,lo): running-max end with its last index;  window rows with end <= ref.start }
//   non-overlapping right = the first query row with start >= ref.end (= row hi)
//   overlapping rows take precedence, in file order (:326-388):  start <= ref.start -> left (last one wins);
//       else end >= ref.end -> right; else (contained) by the centroid proportion: >= 0.5 -> left if no overlapping left
//       yet, < 0.5 -> right; the last row assigned to the right wins.
// The reference's streaming push-back list loses candidates when reference rows are nested (SURVEY 8c hazard 3);
// this implementation always answers the declarative rule (parity notes).
#include "common.cuh"
#include "emit.cuh"
#include "fmt.cuh"
#include "parse.cuh"

namespace bk {

constexpr uint32_t kNoRow = 0xFFFFFFFFu;

struct ClosestParams {
  const uint32_t* rs;
  const uint32_t* re;
  uint64_t        row0, n;
  const uint64_t* run_ref_begin;  // [nruns+1]
  const uint64_t* run_q_begin;    // [nruns]
  const uint64_t* run_q_end;
  int             nruns;
  const uint32_t* qs;
  const uint32_t* qe;
  const uint32_t* pm;   // running max of end within the run
  const uint32_t* pmi;  // 1 + (index within the run) of the last row attaining pm, same scan
  int             allow_overlaps;
  uint64_t*       left;   // global query row or UINT64_MAX
  uint64_t*       right;
};

// pmi source: v[k] = (end[k] == pm[k]) ? k - run_begin + 1 : 0
__global__ void k_argmark(const uint32_t* __restrict__ end, const uint32_t* __restrict__ pm, uint32_t* __restrict__ v,
                          uint64_t n, const uint64_t* __restrict__ run_begin, int nruns) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    int lo = 0, hi = nruns;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (run_begin[mid] <= k) lo = mid; else hi = mid;
    }
    v[k] = end[k] == pm[k] ? (uint32_t)(k - run_begin[lo] + 1) : 0u;
  }
}

constexpr int CF_THREADS = 256;

__global__ void __launch_bounds__(CF_THREADS) k_closest(ClosestParams p) {
  const int      lane = threadIdx.x & 31;
  const uint64_t warp0 = ((uint64_t)blockIdx.x * CF_THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * CF_THREADS) >> 5;
  const uint64_t nbatch = (p.n + 31) >> 5;
  for (uint64_t batch = warp0; batch < nbatch; batch += nwarps) {
    const uint64_t i = (batch << 5) + lane;
    const bool     valid = i < p.n;
    const uint64_t row = p.row0 + (valid ? i : p.n - 1);
    const uint32_t my_rs = p.rs[row], my_re = p.re[row];
    int my_run = 0;
    {
      int hi_r = p.nruns;
      while (hi_r - my_run > 1) {
        int mid = (my_run + hi_r) >> 1;
        if (p.run_ref_begin[mid] <= row) my_run = mid; else hi_r = mid;
      }
    }
    // Candidate windows of the whole batch at once (same scheme as k_map_stats): lo = first query row whose running-max
    // end exceeds ref.start, hi = first query row with start >= ref.end.  Two warp-cooperative searches bracket the
    // batch's lo values (reference rows are sorted by start), one bounds every hi, then each lane bisects its own.
    const int nj = (int)((p.n - (batch << 5)) < 32 ? (p.n - (batch << 5)) : 32);
    uint32_t  my_lo, my_hi;
    {
      const int run_a = __shfl_sync(0xffffffffu, my_run, 0);
      if (__all_sync(0xffffffffu, my_run == run_a)) {
        const uint64_t  qb0 = p.run_q_begin[run_a];
        const uint32_t  nr0 = (uint32_t)(p.run_q_end[run_a] - qb0);
        const uint32_t* pm0 = p.pm + qb0;
        const uint32_t* qs0 = p.qs + qb0;
        const uint32_t  key_a = __shfl_sync(0xffffffffu, my_rs, 0) + 1u, key_b = __shfl_sync(0xffffffffu, my_rs, nj - 1) + 1u;
        uint32_t        a = warp_search32(pm0, nr0, key_a, lane);
        uint32_t        b = warp_gallop(pm0, a, nr0, key_b, lane, true);
        while (a < b) {
          const uint32_t mid = a + ((b - a) >> 1);
          if (__ldg(&pm0[mid]) < my_rs + 1u) a = mid + 1; else b = mid;
        }
        my_lo = a;
        const uint32_t hmax = warp_search32(qs0, nr0, __reduce_max_sync(0xffffffffu, my_re), lane);
        uint32_t       h0 = my_lo, h1 = hmax > my_lo ? hmax : my_lo;
        while (h0 < h1) {
          const uint32_t mid = h0 + ((h1 - h0) >> 1);
          if (__ldg(&qs0[mid]) < my_re) h0 = mid + 1; else h1 = mid;
        }
        my_hi = h0;
      } else {  // the batch straddles a chromosome boundary: every lane searches its own chromosome
        const uint64_t qbl = p.run_q_begin[my_run];
        const uint64_t nrl = p.run_q_end[my_run] - qbl;
        my_lo = (uint32_t)lower_bound_u32(p.pm + qbl, 0, nrl, (uint64_t)my_rs + 1);
        my_hi = (uint32_t)lower_bound_u32(p.qs + qbl, my_lo, nrl, my_re);
      }
    }
    uint64_t out_left = ~0ull, out_right = ~0ull;
    int      hint_run = -1;
    uint32_t nr = 0;
    uint64_t qb = 0;
    const uint32_t *qs = nullptr, *qe = nullptr, *pm = nullptr, *pmi = nullptr;
#pragma unroll 1
    for (int j = 0; j < nj; j++) {
      const uint32_t rs = __shfl_sync(0xffffffffu, my_rs, j), re = __shfl_sync(0xffffffffu, my_re, j);
      const int      run = __shfl_sync(0xffffffffu, my_run, j);
      const uint32_t lo = __shfl_sync(0xffffffffu, my_lo, j), hi = __shfl_sync(0xffffffffu, my_hi, j);
      if (run != hint_run) {
        qb = p.run_q_begin[run];
        nr = (uint32_t)(p.run_q_end[run] - qb);
        qs = p.qs + qb; qe = p.qe + qb; pm = p.pm + qb; pmi = p.pmi + qb;
        hint_run = run;
      }
      // non-overlapping left candidate from the prefix [0,lo): (max end, last row attaining it)
      unsigned long long bestL = 0;  // (end + 1) << 32 | (row-in-run + 1); 0 = none
      if (lo > 0) bestL = ((unsigned long long)(__ldg(&pm[lo - 1]) + 1ull) << 32) | __ldg(&pmi[lo - 1]);
      uint32_t lastL = 0, firstC = 0xFFFFFFFFu, lastR = 0;  // row-in-run + 1 (0 / ~0 = none)
      const double centroid = ((double)re - 1.0 + (double)rs) / 2.0;  // getCentroid, :236-239
#pragma unroll 1
      for (uint32_t k = lo + lane; k < hi; k += 32) {
        const uint32_t s = __ldg(&qs[k]), e = __ldg(&qe[k]);
        if (e <= rs) {  // left of the reference (start < end <= ref.start)
          const unsigned long long key = ((unsigned long long)(e + 1ull) << 32) | (k + 1u);
          bestL = key > bestL ? key : bestL;
        } else if (p.allow_overlaps) {  // overlaps: start < ref.end (k < hi) and end > ref.start
          if (s <= rs) {
            lastL = k + 1u;
          } else if (e >= re) {
            lastR = k + 1u;
          } else {
            // proportionOverlapLeft(c, centroid), :226-231
            const double prop = centroid < (double)s ? 0.0 : (centroid + 1.0 - (double)s) / ((double)e - (double)s);
            if (prop >= 0.5) firstC = firstC < k + 1u ? firstC : k + 1u;
            else lastR = k + 1u;
          }
        }
      }
      // warp reductions
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, bestL, d);
        bestL = o > bestL ? o : bestL;
      }
      lastL = __reduce_max_sync(0xffffffffu, lastL);
      lastR = __reduce_max_sync(0xffffffffu, lastR);
      firstC = __reduce_min_sync(0xffffffffu, firstC);
      uint32_t l = kNoRow, r = kNoRow;  // row-in-run
      if (lastL) l = lastL - 1;
      else if (firstC != 0xFFFFFFFFu) l = firstC - 1;
      else if (bestL) l = (uint32_t)(bestL & 0xFFFFFFFFull) - 1;
      // a contained row with proportion < 0.5 that precedes the first ">= 0.5" row goes right either way; rows after
      // it with proportion >= 0.5 are ignored once an overlapping left exists (:356-368) -- both covered by lastR
      if (lastR) r = lastR - 1;
      else if (hi < nr) r = hi;
      if (lane == j) {
        out_left = l == kNoRow ? ~0ull : qb + l;
        out_right = r == kNoRow ? ~0ull : qb + r;
      }
    }
    if (valid) {
      p.left[i] = out_left;
      p.right[i] = out_right;
    }
  }
}

struct ClosestRow {
  const char*     rtext;
  const uint64_t* rline;
  const uint32_t* rs;
  const uint32_t* re;
  uint64_t        row0;
  const char*     qtext;
  const uint64_t* qline;
  const uint32_t* qs;
  const uint32_t* qe;
  const uint64_t* left;
  const uint64_t* right;
  int             dist, closest, no_ref;
  char            delim[24];
  int             delim_len;

  // signed distance of query element (s,e) from reference (a,b): getDistance(c, b), ClosestFeature.cpp:244-255
  __device__ __forceinline__ long long distance(uint32_t s, uint32_t e, uint32_t a, uint32_t b) const {
    if (e <= a) return -((long long)a - (long long)e + 1);
    if (b <= s) return (long long)s - (long long)b + 1;
    return 0;
  }
  template <class Sink>
  __device__ __forceinline__ void put_elem(Sink& s, uint64_t q, uint32_t a, uint32_t b, bool last) const {
    if (q == ~0ull) {
      s.puts_("NA", 2);
      if (dist) {
        s.puts_(delim, delim_len);
        s.puts_("NA", 2);
      }
    } else {
      echo_b3rest(s, qtext, qline[q], qs[q], qe[q]);
      if (dist) {
        s.puts_(delim, delim_len);
        long long d = distance(qs[q], qe[q], a, b);
        if (d < 0) {
          s.put('-');
          d = -d;
        }
        s.put_u64((uint64_t)d);
      }
    }
    if (last) s.put('\n'); else s.puts_(delim, delim_len);
  }
  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& s) const {
    const uint64_t row = row0 + i;
    const uint32_t a = rs[row], b = re[row];
    if (!no_ref) {
      echo_b3rest(s, rtext, rline[row], a, b);
      s.puts_(delim, delim_len);
    }
    const uint64_t l = left[i], r = right[i];
    if (!closest) {  // PrintAll, Printers.hpp:46-98
      put_elem(s, l, a, b, false);
      put_elem(s, r, a, b, true);
      return;
    }
    // PrintShortest, Printers.hpp:104-203
    uint64_t pick;
    if (l == ~0ull && r == ~0ull) pick = ~0ull;
    else if (l == ~0ull) pick = r;
    else if (qe[l] > a) pick = l;            // overlapping left wins outright
    else if (r == ~0ull) pick = l;
    else if (b > qs[r]) pick = r;            // overlapping right
    else pick = ((unsigned long long)a - qe[l]) <= ((unsigned long long)qs[r] - b) ? l : r;  // ties go left
    put_elem(s, pick, a, b, true);
  }
};

int finish_text(bk_ctx* ctx, char* d_out, uint64_t bytes, uint64_t rows, int on_device, bk_text* out);

}  // namespace bk

using namespace bk;

extern "C" void bk_cfspec_default(bk_cfspec* spec) {
  memset(spec, 0, sizeof(*spec));
  spec->delim = "|";
}

extern "C" int bk_closest(bk_ctx* ctx, const bk_bed* ref, const bk_bed* query, const bk_cfspec* spec, bk_text* out) {
  if (!ctx || !ref || !query || !spec || !out) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  if ((!ref->line_off && ref->nrows) || (!query->line_off && query->nrows))
    return fail(ctx, BK_ERR_ARG, "closest-features needs both files loaded with BK_COL_LINE");
  if (spec->center || spec->no_query) return fail(ctx, BK_ERR_UNSUPPORTED, "closest-features option outside the device hot path");
  if (query->nrows >= 0xFFFFFFFEull) return fail(ctx, BK_ERR_UNSUPPORTED, "query file has more than 2^32-2 rows");
  const char* delim = spec->delim ? spec->delim : "|";
  if (strlen(delim) > 23) return fail(ctx, BK_ERR_UNSUPPORTED, "delimiter longer than 23 bytes");

  const bool all = !spec->chrom || !*spec->chrom || strcmp(spec->chrom, "all") == 0;
  std::vector<uint64_t> rrb, qb, qe;
  uint64_t row0 = 0, row1 = 0;
  bool     first = true;
  for (auto& r : ref->runs) {
    if (!all && r.name != spec->chrom) continue;
    if (r.row_end == r.row_begin) continue;
    if (first) { row0 = r.row_begin; first = false; }
    row1 = r.row_end;
    rrb.push_back(r.row_begin);
    uint64_t b = 0, e = 0;
    for (auto& q : query->runs)
      if (q.name == r.name) { b = q.row_begin; e = q.row_end; }
    qb.push_back(b);
    qe.push_back(e);
  }
  const uint64_t n = row1 - row0;
  if (n == 0) return finish_text(ctx, nullptr, 0, 0, spec->out_on_device, out);
  rrb.push_back(row1);
  const int nruns = (int)qb.size();

  BK_TRY(ensure_pmax(ctx, query));
  // last row attaining the running max (ties: the later row wins, ClosestFeature.cpp:300)
  uint32_t* d_pmi = dalloc<uint32_t>(ctx, query->nrows);
  if (!d_pmi) return BK_ERR_NOMEM;
  if (query->nrows) {
    std::vector<uint64_t> qrb;
    for (auto& q : query->runs) qrb.push_back(q.row_begin);
    uint64_t* d_qrb = dalloc<uint64_t>(ctx, qrb.size());
    uint32_t* d_mark = dalloc<uint32_t>(ctx, query->nrows);
    if (!d_qrb || !d_mark) return BK_ERR_NOMEM;
    BK_CUDA(ctx, cudaMemcpyAsync(d_qrb, qrb.data(), qrb.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t blocks = (query->nrows + 255) / 256, cap = (uint64_t)ctx->sms * 16;
    prof_begin(ctx, "k_argmark");
    k_argmark<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(query->end, query->pmax_end, d_mark, query->nrows,
                                                                                d_qrb, (int)qrb.size());
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = seg_prefix_max(ctx, d_mark, d_pmi, query->nrows, query->runs);
    dfree(ctx, d_qrb);
    dfree(ctx, d_mark);
    if (rc != BK_OK) return rc;
  }

  std::vector<uint64_t> tab;
  tab.insert(tab.end(), rrb.begin(), rrb.end());
  tab.insert(tab.end(), qb.begin(), qb.end());
  tab.insert(tab.end(), qe.begin(), qe.end());
  uint64_t* d_tab = dalloc<uint64_t>(ctx, tab.size());
  if (!d_tab) return BK_ERR_NOMEM;
  BK_CUDA(ctx, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  ClosestParams p{};
  p.rs = ref->start; p.re = ref->end; p.row0 = row0; p.n = n;
  p.run_ref_begin = d_tab; p.run_q_begin = d_tab + nruns + 1; p.run_q_end = d_tab + 2 * nruns + 1; p.nruns = nruns;
  p.qs = query->start; p.qe = query->end; p.pm = query->pmax_end; p.pmi = d_pmi;
  p.allow_overlaps = !spec->no_overlaps;
  p.left = dalloc<uint64_t>(ctx, n);
  p.right = dalloc<uint64_t>(ctx, n);
  if (!p.left || !p.right) return BK_ERR_NOMEM;
  {
    const uint64_t batches = (n + 31) / 32, per_block = CF_THREADS / 32;
    uint64_t       blocks = (batches + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t)ctx->sms * 8 * 8;
    if (blocks > cap) blocks = cap;
    prof_begin(ctx, "k_closest");
    k_closest<<<(unsigned)blocks, CF_THREADS, 0, ctx->stream>>>(p);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
  }
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_tab);
  dfree(ctx, d_pmi);

  ClosestRow fn{};
  fn.rtext = ref->d_text; fn.rline = ref->line_off; fn.rs = ref->start; fn.re = ref->end; fn.row0 = row0;
  fn.qtext = query->d_text; fn.qline = query->line_off; fn.qs = query->start; fn.qe = query->end;
  fn.left = p.left; fn.right = p.right;
  fn.dist = spec->dist; fn.closest = spec->closest; fn.no_ref = spec->no_ref;
  fn.delim_len = (int)strlen(delim);
  memcpy(fn.delim, delim, fn.delim_len);
  // bound: the reference line + two query lines (each at most the longest query line) + distances
  // the longest query line is not known cheaply; bound it by the whole query text when the file is tiny, otherwise
  // by a generous per-row constant and let the emitter report an overflow
  uint64_t per_row = 3 * (uint64_t)fn.delim_len + 2 * 24 + 4 + 2 * 22;
  uint64_t cap = ref->nbytes + n * per_row + 64;
  uint64_t qavg = query->nrows ? (query->nbytes / query->nrows + 1) : 0;
  cap += n * 2 * (qavg * 4 + 64);
  if (query->nbytes < (1u << 20)) cap += n * 2 * query->nbytes;
  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  int rc = run_emit(ctx, fn, n, cap, &d_out, &bytes, &rows);
  dfree(ctx, p.left);
  dfree(ctx, p.right);
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return rc;
  }
  return finish_text(ctx, d_out, bytes, rows, spec->out_on_device, out);
}
