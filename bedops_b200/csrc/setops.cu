// setops.cu -- the bedops set operators as scans and binary searches: --merge, --intersect,
// --element-of / --not-element-of, --complement, --difference, --symmdiff, --everything, --chop.
// Replaces the streaming k-way state machines nextMergeAllLines (Bedops.cpp:1186-1243),
// getNextFileMergedCoords/mergeOverlap (:791-814, :864-886), nextIntersectLine (:1105-1181), nextElementOfLine
// (:1023-1100), nextComplementLine (:891-943), nextDifferenceLine (:948-1018), nextSymmetricDiffLine (:1341-1463),
// nextUnionAllLine (:1468-1516) and doChop (:438-467).
//
//   union-merge of k sorted files:  every file is merged within itself, the segment lists are rank-merged (each row's
//       output slot = its own rank + lower/upper_bound ranks in the other lists, per chromosome) -> segmented
//       prefix-max of end -> a row opens a segment iff start > running max end (touching intervals coalesce:
//       "bt->start() <= toRtn->end()", Bedops.cpp:1233) -> compaction in two passes (heads per warp range, scan, write).
//   intersect / difference: every file is self-merged, then folded pairwise; for a segment a of A the overlapping
//       segments of the disjoint sorted list B are the index range [lower_bound(B.end, a.start+1),
//       lower_bound(B.start, a.end)); the pieces are the overlaps (-i) or what lies between them (-d).
//   symmdiff: union of all files minus the union of the pairwise intersections.  complement: gaps of the union.
//   element-of: overlap bases of each reference row with the union-merge of the other files, compared with the
//       threshold exactly as Bedops.cpp:1094-1099 does (double arithmetic), then the kept rows are echoed.
//   everything: rows ranked across files by (start, end, rest of line, file).  chop: closed-form piece counts.
// No kernel here waits on another CTA: every compaction is count -> scan -> write over warp-owned ranges.
#include <algorithm>
#include <map>
#include "common.cuh"
#include "emit.cuh"
#include "sort.cuh"
#include "fmt.cuh"
#include "parse.cuh"

namespace bk {

constexpr int kMaxFiles = 64;

// a sorted interval list on the device with its chromosome runs (runs may be empty)
struct IvList {
  uint32_t* s = nullptr;
  uint32_t* e = nullptr;
  uint64_t  n = 0;
  bool      owned = false;
  std::vector<ChromRun> runs;
};

static void free_list(bk_ctx* ctx, IvList& l) {
  if (l.owned) {
    dfree(ctx, l.s);
    dfree(ctx, l.e);
  }
  l.s = l.e = nullptr;
  l.n = 0;
  l.owned = false;
}

static IvList view_of(const bk_bed* b, const char* chrom) {
  IvList l;
  l.s = b->start;
  l.e = b->end;
  l.n = b->nrows;
  const bool all = !chrom || !*chrom || strcmp(chrom, "all") == 0;
  for (auto& r : b->runs)
    if (all || r.name == chrom) l.runs.push_back(r);
  return l;
}

// ---- rank-merge -------------------------------------------------------------------------------------------
struct RankParams {
  const uint32_t* s;
  const uint32_t* e;
  const uint64_t* run_begin;  // [nruns+1] rows of THIS file that take part (selected runs, ascending)
  const uint32_t* run_g;      // [nruns] global chromosome index of each run
  int             nruns;
  int             k, f;
  const uint32_t* fs[kMaxFiles];  // start columns of every file
  const uint64_t* g_begin;        // [G*k] row range of chromosome g in file j (begin == end if absent)
  const uint64_t* g_end;
  const uint64_t* g_out;          // [G] first output slot of chromosome g
  uint32_t*       outS;
  uint32_t*       outE;
  uint64_t        nsel;       // number of participating rows of this file
  const uint64_t* sel_prefix; // [nruns+1] prefix of participating rows per run
};

__global__ void __launch_bounds__(256) k_rank_merge(RankParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.nsel; t += stride) {
    int lo = 0, hi = p.nruns;  // last run with sel_prefix <= t
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.sel_prefix[mid] <= t) lo = mid; else hi = mid;
    }
    const uint64_t row = p.run_begin[lo] + (t - p.sel_prefix[lo]);
    const uint32_t g = p.run_g[lo];
    const uint32_t key = p.s[row];
    uint64_t       pos = p.g_out[g] + (row - p.g_begin[(uint64_t)g * p.k + p.f]);
    for (int j = 0; j < p.k; j++) {
      if (j == p.f) continue;
      const uint64_t b = p.g_begin[(uint64_t)g * p.k + j], e = p.g_end[(uint64_t)g * p.k + j];
      if (b == e) continue;
      // ties: rows of lower-numbered files first (stable)
      pos += (j < p.f ? upper_bound_u32(p.fs[j], b, e, key) : lower_bound_u32(p.fs[j], b, e, key)) - b;
    }
    p.outS[pos] = key;
    p.outE[pos] = p.e[row];
  }
}

// ---- segments from a sorted list + its inclusive prefix-max of end ----------------------------------------
// Merged segments of a start-sorted list with its running-max-end index: row i opens a segment iff it is the first
// row of its chromosome or start[i] > pmax[i-1] (touching intervals coalesce: Bedops.cpp:872, :1233); the segment ends
// at the pmax of its last row.  Two passes over warp ranges that never cross a chromosome (same cut as the prefix-max
// index): count the heads of every range, scan, then write -- no carry chain between tiles.
struct SegParams {
  const uint32_t* s;
  const uint32_t* pm;
  const PmRun*    runs;   // [nruns] non-empty chromosome runs
  int             nruns;
  uint64_t        nranges;
  uint64_t*       range_heads;       // [nranges] pass 1 out
  const uint64_t* range_base;        // [nranges+1] exclusive scan of range_heads
  uint32_t*       outS;
  uint32_t*       outE;
  uint64_t*       run_seg_begin;     // [nruns] first output segment of each run
};

constexpr int SEG_THREADS = 256;
__global__ void __launch_bounds__(SEG_THREADS) k_segment_heads(SegParams p) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * SEG_THREADS + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * SEG_THREADS) >> 5;
  for (uint64_t r = w0; r < p.nranges; r += nw) {
    uint64_t a, b, first;
    pm_locate(p.runs, p.nruns, r, a, b, first);
    const bool run_start = r == first;  // row a is the first row of its chromosome
    uint32_t   heads = 0;
    for (uint64_t k = a + lane; k < b; k += 32) {
      const bool head = (k == a && run_start) || (!(k == a && run_start) && __ldg(&p.s[k]) > __ldg(&p.pm[k - 1]));
      heads += head ? 1u : 0u;
    }
    heads = __reduce_add_sync(0xffffffffu, heads);
    if (lane == 0) p.range_heads[r] = heads;
  }
}

__global__ void __launch_bounds__(SEG_THREADS) k_segments(SegParams p) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * SEG_THREADS + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * SEG_THREADS) >> 5;
  for (uint64_t r = w0; r < p.nranges; r += nw) {
    uint64_t  a, b, first;
    const int run = pm_locate(p.runs, p.nruns, r, a, b, first);
    const bool     run_start = r == first;
    const uint64_t run_end = p.runs[run].row_end;
    uint64_t       seg = p.range_base[r];  // segments opened before this range
    if (run_start && lane == 0) p.run_seg_begin[run] = seg;
    for (uint64_t k0 = a; k0 < b; k0 += 32) {
      const uint64_t k = k0 + lane;
      bool           head = false, tail = false;
      uint32_t       sv = 0, pv = 0;
      if (k < b) {
        sv = __ldg(&p.s[k]);
        pv = __ldg(&p.pm[k]);
        head = (k == a && run_start) || (!(k == a && run_start) && sv > __ldg(&p.pm[k - 1]));
        tail = k + 1 == run_end || __ldg(&p.s[k + 1]) > pv;  // the next row opens a segment (k+1 < run_end: same run)
      }
      const unsigned hm = __ballot_sync(0xffffffffu, head);
      const uint64_t mine = seg + __popc(hm & ((2u << lane) - 1u));  // heads up to and including this row
      if (head) p.outS[mine - 1] = sv;
      if (tail) p.outE[mine - 1] = pv;
      seg += __popc(hm);
    }
  }
}

// ---- pairwise intersection of two disjoint sorted lists ---------------------------------------------------
struct IsectParams {
  const uint32_t* as;
  const uint32_t* ae;
  uint64_t        row0, n;        // A rows [row0,row0+n) take part
  const uint64_t* run_a_begin;    // [nruns+1] A rows per common chromosome (ascending)
  const uint64_t* run_b_begin;    // [nruns]
  const uint64_t* run_b_end;
  int             nruns;
  const uint32_t* bs;
  const uint32_t* be;
  uint32_t*       outS;
  uint32_t*       outE;
  uint64_t*       run_out_begin;  // [nruns]
  uint32_t*       local;          // [n] pieces of the earlier rows of the same warp range (pass 1)
  uint64_t*       range_total;    // [nranges]
  const uint64_t* range_base;     // [nranges+1]
  uint64_t        nranges;
};
constexpr int IS_RANGE = 1024;  // A rows per warp range

// B's segments [lo,hi) that overlap A row `row`, and the number of pieces the row produces.
// DIFF = false: pieces of A covered by B (intersection).  DIFF = true: pieces of A that B does NOT cover (difference,
// nextDifferenceLine, Bedops.cpp:948-1018); B's segments are disjoint and sorted.
template <bool DIFF>
__device__ __forceinline__ uint32_t isect_row(const IsectParams& p, uint64_t row, int& run, uint32_t& a0, uint32_t& a1,
                                              uint64_t& lo, uint64_t& hi) {
  int l = 0, h = p.nruns;
  while (h - l > 1) {
    const int mid = (l + h) >> 1;
    if (p.run_a_begin[mid] <= row) l = mid; else h = mid;
  }
  run = l;
  a0 = p.as[row];
  a1 = p.ae[row];
  const uint64_t bb = p.run_b_begin[l], bend = p.run_b_end[l];
  lo = lower_bound_u32(p.be, bb, bend, (uint64_t)a0 + 1);  // first b.end > a.start
  hi = lower_bound_u32(p.bs, bb, bend, (uint64_t)a1);      // first b.start >= a.end
  if (hi < lo) hi = lo;
  if (!DIFF) return (uint32_t)(hi - lo);
  uint32_t np = (uint32_t)(hi - lo) + 1;
  if (hi > lo) {
    if (p.bs[lo] <= a0) np--;      // no piece in front of the first covering segment
    if (p.be[hi - 1] >= a1) np--;  // none behind the last
  }
  return np;
}

// pass 1: pieces per row, warp-local prefixes over contiguous ranges of rows (no carry chain between tiles)
template <bool DIFF>
__global__ void __launch_bounds__(SEG_THREADS) k_intersect_count(IsectParams p) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * SEG_THREADS + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * SEG_THREADS) >> 5;
  for (uint64_t r = w0; r < p.nranges; r += nw) {
    const uint64_t a = r * IS_RANGE, b = a + IS_RANGE < p.n ? a + IS_RANGE : p.n;
    uint64_t       acc = 0;
    for (uint64_t k0 = a; k0 < b; k0 += 32) {
      const uint64_t i = k0 + lane;
      uint32_t       np = 0;
      if (i < b) {
        int      run;
        uint32_t a0, a1;
        uint64_t lo, hi;
        np = isect_row<DIFF>(p, p.row0 + i, run, a0, a1, lo, hi);
      }
      const uint32_t incl = warp_incl_scan(np);
      if (i < b) p.local[i] = (uint32_t)acc + (incl - np);  // a range of 1024 rows of 32-bit counts: < 2^42, host-checked
      acc += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) p.range_total[r] = acc;
  }
}

// pass 2: every row writes its pieces at range_base + local
template <bool DIFF>
__global__ void __launch_bounds__(SEG_THREADS) k_intersect(IsectParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const uint64_t row = p.row0 + i;
    int            run;
    uint32_t       a0, a1;
    uint64_t       lo, hi;
    isect_row<DIFF>(p, row, run, a0, a1, lo, hi);
    uint64_t o = p.range_base[i / IS_RANGE] + p.local[i];
    if (p.run_a_begin[run] == row) p.run_out_begin[run] = o;
    if (DIFF) {
      uint32_t cur = a0;
      for (uint64_t k = lo; k < hi; k++) {
        const uint32_t b0 = p.bs[k], b1 = p.be[k];
        if (b0 > cur) {
          p.outS[o] = cur;
          p.outE[o] = b0;
          o++;
        }
        cur = b1 > cur ? b1 : cur;
      }
      if (cur < a1) {
        p.outS[o] = cur;
        p.outE[o] = a1;
      }
    } else {
      for (uint64_t k = lo; k < hi; k++, o++) {
        const uint32_t b0 = p.bs[k], b1 = p.be[k];
        p.outS[o] = a0 > b0 ? a0 : b0;
        p.outE[o] = a1 < b1 ? a1 : b1;
      }
    }
  }
}

// ---- element-of: keep flag per reference row ---------------------------------------------------------------
struct ElemParams {
  const uint32_t* rs;
  const uint32_t* re;
  uint64_t        row0, n;
  const uint64_t* run_ref_begin;  // [nruns+1]
  const uint64_t* run_u_begin;    // [nruns]
  const uint64_t* run_u_end;
  int             nruns;
  const uint32_t* us;
  const uint32_t* ue;
  double          thr;
  int             use_pct, invert;
  const uint8_t*  run_has_later;  // [nruns] the union holds an element on a later chromosome
  uint8_t*        keep;
};

__global__ void __launch_bounds__(256) k_element_of(ElemParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const uint64_t row = p.row0 + i;
    int l = 0, h = p.nruns;
    while (h - l > 1) {
      int mid = (l + h) >> 1;
      if (p.run_ref_begin[mid] <= row) l = mid; else h = mid;
    }
    const uint32_t a0 = p.rs[row], a1 = p.re[row];
    const uint64_t ub = p.run_u_begin[l], uend = p.run_u_end[l];
    uint64_t       lo = lower_bound_u32(p.ue, ub, uend, (uint64_t)a0 + 1);
    const uint64_t hi = lower_bound_u32(p.us, ub, uend, (uint64_t)a1);
    double         range_overlap = 0;  // "double rangeOverlap", Bedops.cpp:1062
    for (; lo < hi; lo++) {
      const uint32_t b0 = p.us[lo], b1 = p.ue[lo];
      const uint32_t mn = a0 > b0 ? a0 : b0, mx = a1 < b1 ? a1 : b1;
      range_overlap += (double)(mx - mn);
    }
    const double range = (double)(a1 - a0);
    bool         is_elem = p.use_pct ? (range_overlap / range >= p.thr) : (range_overlap >= p.thr);
    // "-e 0": 0 >= 0 holds for every row the streaming loop reaches, but a row with no merged element at or after
    // it is reported as a non-element before the test (Bedops.cpp:1044-1049, :1053-1062)
    if (!p.use_pct && p.thr <= 0) is_elem = p.run_has_later[l] || (uend > ub && p.ue[uend - 1] > a0);
    p.keep[i] = (is_elem != (p.invert != 0)) ? 1 : 0;
  }
}

// ---- emitters -------------------------------------------------------------------------------------------------
struct NameTable {
  const char*     names;  // [nruns][128]
  const uint32_t* len;
  const uint64_t* run_begin;
  int             nruns;
};

struct Bed3Row {  // chrom \t start \t end \n   (B3NoRest::println, Bed.hpp:228-232)
  const uint32_t* s;
  const uint32_t* e;
  NameTable       nt;
  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& sk) const {
    int l = 0, h = nt.nruns;
    while (h - l > 1) {
      int mid = (l + h) >> 1;
      if (nt.run_begin[mid] <= i) l = mid; else h = mid;
    }
    sk.puts_(nt.names + (size_t)l * 128, (int)nt.len[l]);
    sk.put('\t');
    sk.put_u32(s[i]);
    sk.put('\t');
    sk.put_u32(e[i]);
    sk.put('\n');
  }
};

struct EchoKeptRow {  // record(nextRef) for -e/-n: B3Rest::println (Bedops.cpp:148-152, :555-556)
  const char*     text;
  const uint64_t* line;
  const uint32_t* s;
  const uint32_t* e;
  const uint8_t*  keep;
  uint64_t        row0;
  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& sk) const {
    if (!keep[i]) return;
    const uint64_t row = row0 + i;
    echo_b3rest(sk, text, line[row], s[row], e[row]);
    sk.put('\n');
  }
};

int finish_text(bk_ctx* ctx, char* d_out, uint64_t bytes, uint64_t rows, int on_device, bk_text* out);

// upload a host vector; the caller must sync before the vector dies (all callers below sync via read_scratch)
template <class T>
static T* upload(bk_ctx* ctx, const std::vector<T>& v) {
  T* d = dalloc<T>(ctx, v.size());
  if (d && !v.empty()) cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
  return d;
}

// union-merge of k lists -> disjoint sorted segments (touching coalesced)
// start-sorted rows (S, E) with their chromosome runs -> the disjoint merged segments
static int merge_list(bk_ctx* ctx, const uint32_t* S, const uint32_t* E, uint64_t N, const std::vector<ChromRun>& runs,
                      IvList* out) {
  out->runs.clear();
  out->n = 0;
  out->owned = true;
  out->s = out->e = nullptr;
  std::vector<PmRun>       pr;
  std::vector<std::string> names;
  uint64_t                 nranges = 0;
  for (const ChromRun& r : runs) {
    if (r.row_end == r.row_begin) continue;
    pr.push_back({r.row_begin, r.row_end, nranges});
    names.push_back(r.name);
    nranges += (r.row_end - r.row_begin + PM_RANGE - 1) / PM_RANGE;
  }
  if (N == 0 || nranges == 0) return BK_OK;
  const int G = (int)pr.size();
  uint32_t* pm = dalloc<uint32_t>(ctx, N);
  if (!pm) return BK_ERR_NOMEM;
  BK_TRY(seg_prefix_max(ctx, E, pm, N, runs));
  SegParams sp{};
  sp.s = S; sp.pm = pm; sp.nruns = G; sp.nranges = nranges;
  PmRun*    d_runs = dalloc<PmRun>(ctx, pr.size());
  uint64_t* d_heads = dalloc<uint64_t>(ctx, nranges);
  uint64_t* d_base = dalloc<uint64_t>(ctx, nranges + 1);
  sp.run_seg_begin = dalloc<uint64_t>(ctx, G);
  if (!d_runs || !d_heads || !d_base || !sp.run_seg_begin) return BK_ERR_NOMEM;
  if (nranges >> 32) return fail(ctx, BK_ERR_UNSUPPORTED, "more than 2^32 row ranges");
  BK_CUDA(ctx, cudaMemcpyAsync(d_runs, pr.data(), pr.size() * sizeof(PmRun), cudaMemcpyHostToDevice, ctx->stream));
  sp.runs = d_runs; sp.range_heads = d_heads; sp.range_base = d_base;
  const uint64_t want = (nranges + SEG_THREADS / 32 - 1) / (SEG_THREADS / 32);
  BK_TRY(reset_scratch(ctx));
  prof_begin(ctx, "k_segment_heads");
  k_segment_heads<<<grid_for_kernel((const void*)k_segment_heads, SEG_THREADS, want), SEG_THREADS, 0, ctx->stream>>>(sp);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  k_scan_totals<SC_OUT_ROWS><<<1, 1024, 0, ctx->stream>>>(d_heads, d_base, (uint32_t)nranges, ctx->d_scratch);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));  // also: pr (host vector) has been copied
  const uint64_t nseg = ctx->h_scratch[SC_OUT_ROWS];
  sp.outS = dalloc<uint32_t>(ctx, nseg);
  sp.outE = dalloc<uint32_t>(ctx, nseg);
  if (!sp.outS || !sp.outE) return BK_ERR_NOMEM;
  prof_begin(ctx, "k_segments");
  k_segments<<<grid_for_kernel((const void*)k_segments, SEG_THREADS, want), SEG_THREADS, 0, ctx->stream>>>(sp);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  std::vector<uint64_t> rsb(G);
  BK_CUDA(ctx, cudaMemcpyAsync(rsb.data(), sp.run_seg_begin, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_runs); dfree(ctx, d_heads); dfree(ctx, d_base); dfree(ctx, sp.run_seg_begin); dfree(ctx, pm);
  out->s = sp.outS; out->e = sp.outE; out->n = nseg;
  for (int g = 0; g < G; g++) out->runs.push_back({names[g], rsb[g], g + 1 < G ? rsb[g + 1] : nseg});
  return BK_OK;
}

static int union_merge(bk_ctx* ctx, const std::vector<IvList>& in_files, IvList* out) {
  // Several files: merge every file within itself first (what getNextFileMergedCoords does for --intersect,
  // Bedops.cpp:791-814) and interleave the -- usually far fewer -- segments; the union is the same.
  std::vector<IvList> merged_files;
  if (in_files.size() > 1) {
    merged_files.resize(in_files.size());
    for (size_t f = 0; f < in_files.size(); f++) {
      std::vector<IvList> one{in_files[f]};
      int rc = union_merge(ctx, one, &merged_files[f]);
      if (rc != BK_OK) {
        for (auto& m : merged_files) free_list(ctx, m);
        return rc;
      }
    }
  }
  struct Cleanup {
    bk_ctx* ctx;
    std::vector<IvList>& v;
    ~Cleanup() { for (auto& m : v) free_list(ctx, m); }
  } cleanup{ctx, merged_files};
  const std::vector<IvList>& in = in_files.size() > 1 ? merged_files : in_files;
  const int k = (int)in.size();
  // global chromosome table
  std::vector<std::string> names;
  for (auto& l : in)
    for (auto& r : l.runs)
      if (r.row_end > r.row_begin) names.push_back(r.name);
  std::sort(names.begin(), names.end(), [](const std::string& a, const std::string& b) { return strcmp(a.c_str(), b.c_str()) < 0; });
  names.erase(std::unique(names.begin(), names.end()), names.end());
  const int G = (int)names.size();
  out->runs.clear();
  out->n = 0;
  out->owned = true;
  if (G == 0) return BK_OK;
  std::map<std::string, int> gidx;
  for (int g = 0; g < G; g++) gidx[names[g]] = g;
  std::vector<uint64_t> g_begin((size_t)G * k, 0), g_end((size_t)G * k, 0), g_out(G + 1, 0);
  for (int f = 0; f < k; f++)
    for (auto& r : in[f].runs) {
      if (r.row_end == r.row_begin) continue;
      int g = gidx[r.name];
      g_begin[(size_t)g * k + f] = r.row_begin;
      g_end[(size_t)g * k + f] = r.row_end;
    }
  for (int g = 0; g < G; g++) {
    uint64_t c = 0;
    for (int f = 0; f < k; f++) c += g_end[(size_t)g * k + f] - g_begin[(size_t)g * k + f];
    g_out[g + 1] = g_out[g] + c;
  }
  const uint64_t N = g_out[G];
  // merged-order columns
  uint32_t *mS = nullptr, *mE = nullptr;
  bool      merged_owned = false;
  if (k == 1) {
    // a single file is already in order, but only its selected runs take part: they are contiguous iff all
    // runs were selected or exactly one; otherwise compact through the same rank kernel
    bool contiguous = true;
    for (size_t i = 1; i < in[0].runs.size(); i++)
      if (in[0].runs[i].row_begin != in[0].runs[i - 1].row_end) contiguous = false;
    if (contiguous) {
      uint64_t first = UINT64_MAX;
      for (auto& r : in[0].runs)
        if (r.row_end > r.row_begin && r.row_begin < first) first = r.row_begin;
      mS = in[0].s + first;
      mE = in[0].e + first;
    }
  }
  if (!mS) {
    if (k > kMaxFiles) return fail(ctx, BK_ERR_UNSUPPORTED, "more than %d input files", kMaxFiles);
    mS = dalloc<uint32_t>(ctx, N);
    mE = dalloc<uint32_t>(ctx, N);
    if (!mS || !mE) return BK_ERR_NOMEM;
    merged_owned = true;
    uint64_t* d_gb = upload(ctx, g_begin);
    uint64_t* d_ge = upload(ctx, g_end);
    uint64_t* d_go = upload(ctx, g_out);
    if (!d_gb || !d_ge || !d_go) return BK_ERR_NOMEM;
    std::vector<void*> tofree{d_gb, d_ge, d_go};
    std::vector<std::vector<uint64_t>> keep_rb(k), keep_sp(k);
    std::vector<std::vector<uint32_t>> keep_rg(k);
    for (int f = 0; f < k; f++) {
      std::vector<uint64_t>& rb = keep_rb[f];
      std::vector<uint64_t>& sp = keep_sp[f];
      std::vector<uint32_t>& rg = keep_rg[f];
      uint64_t acc = 0;
      for (auto& r : in[f].runs) {
        if (r.row_end == r.row_begin) continue;
        rb.push_back(r.row_begin);
        rg.push_back((uint32_t)gidx[r.name]);
        sp.push_back(acc);
        acc += r.row_end - r.row_begin;
      }
      if (acc == 0) continue;
      rb.push_back(0);
      sp.push_back(acc);
      RankParams p{};
      p.s = in[f].s; p.e = in[f].e;
      p.run_begin = upload(ctx, rb); p.run_g = upload(ctx, rg); p.sel_prefix = upload(ctx, sp);
      if (!p.run_begin || !p.run_g || !p.sel_prefix) return BK_ERR_NOMEM;
      tofree.push_back((void*)p.run_begin); tofree.push_back((void*)p.run_g); tofree.push_back((void*)p.sel_prefix);
      p.nruns = (int)rg.size(); p.k = k; p.f = f;
      for (int j = 0; j < k; j++) p.fs[j] = in[j].s;
      p.g_begin = d_gb; p.g_end = d_ge; p.g_out = d_go;
      p.outS = mS; p.outE = mE; p.nsel = acc;
      uint64_t blocks = (acc + 255) / 256, cap = (uint64_t)kSMs * 32;
      prof_begin(ctx, "k_rank_merge");
      k_rank_merge<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(p);
      prof_end(ctx);
      BK_LAUNCHED(ctx);
    }
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (void* q : tofree) dfree(ctx, q);
  }
  std::vector<ChromRun> mruns;
  for (int g = 0; g < G; g++) mruns.push_back({names[g], g_out[g], g_out[g + 1]});
  int rc = merge_list(ctx, mS, mE, N, mruns, out);
  if (merged_owned) { dfree(ctx, mS); dfree(ctx, mE); }
  return rc;
}

// A ∩ B for disjoint sorted lists
static int intersect_pair(bk_ctx* ctx, const IvList& A, const IvList& B, IvList* out, bool diff = false) {
  out->runs.clear();
  out->n = 0;
  out->owned = true;
  std::vector<uint64_t> ab, bb, be;
  std::vector<std::string> names;
  size_t j = 0;
  for (auto& r : A.runs) {
    if (r.row_end == r.row_begin) continue;
    while (j < B.runs.size() && strcmp(B.runs[j].name.c_str(), r.name.c_str()) < 0) j++;
    if (j < B.runs.size() && B.runs[j].name == r.name && B.runs[j].row_end > B.runs[j].row_begin) {
      // only chromosomes present on both sides can intersect; A rows of other chromosomes are skipped by
      // giving them an empty B range
    }
    ab.push_back(r.row_begin);
    names.push_back(r.name);
    if (j < B.runs.size() && B.runs[j].name == r.name) {
      bb.push_back(B.runs[j].row_begin);
      be.push_back(B.runs[j].row_end);
    } else {
      bb.push_back(0);
      be.push_back(0);
    }
  }
  const int nruns = (int)names.size();
  if (nruns == 0) return BK_OK;
  // participating A rows: from the first selected run to the end of the last one (runs in between that were not
  // selected cannot exist: `runs` only ever holds all runs or one run)
  const uint64_t row0 = ab.front();
  uint64_t       row1 = 0;
  for (auto& r : A.runs)
    if (r.row_end > r.row_begin) row1 = r.row_end;
  ab.push_back(row1);
  const uint64_t n = row1 - row0;
  IsectParams p{};
  p.as = A.s; p.ae = A.e; p.row0 = row0; p.n = n;
  p.run_a_begin = upload(ctx, ab); p.run_b_begin = upload(ctx, bb); p.run_b_end = upload(ctx, be); p.nruns = nruns;
  p.bs = B.s; p.be = B.e;
  p.nranges = (n + IS_RANGE - 1) / IS_RANGE;
  if (p.nranges >> 32) return fail(ctx, BK_ERR_UNSUPPORTED, "too many rows for one set operation");
  p.run_out_begin = dalloc<uint64_t>(ctx, nruns);
  p.local = dalloc<uint32_t>(ctx, n);
  p.range_total = dalloc<uint64_t>(ctx, p.nranges);
  uint64_t* d_base = dalloc<uint64_t>(ctx, p.nranges + 1);
  if (!p.run_a_begin || !p.run_b_begin || !p.run_b_end || !p.run_out_begin || !p.local || !p.range_total || !d_base)
    return BK_ERR_NOMEM;
  p.range_base = d_base;
  BK_CUDA(ctx, cudaMemsetAsync(p.run_out_begin, 0, (size_t)nruns * 8, ctx->stream));
  BK_TRY(reset_scratch(ctx));
  const uint64_t want = (p.nranges + SEG_THREADS / 32 - 1) / (SEG_THREADS / 32);
  prof_begin(ctx, diff ? "k_difference_count" : "k_intersect_count");
  if (diff) k_intersect_count<true><<<grid_for_kernel((const void*)k_intersect_count<true>, SEG_THREADS, want), SEG_THREADS, 0, ctx->stream>>>(p);
  else k_intersect_count<false><<<grid_for_kernel((const void*)k_intersect_count<false>, SEG_THREADS, want), SEG_THREADS, 0, ctx->stream>>>(p);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  k_scan_totals<SC_OUT_ROWS><<<1, 1024, 0, ctx->stream>>>(p.range_total, d_base, (uint32_t)p.nranges, ctx->d_scratch);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));  // exact output size (and: the uploaded host vectors have been copied)
  const uint64_t nout = ctx->h_scratch[SC_OUT_ROWS];
  p.outS = dalloc<uint32_t>(ctx, nout);
  p.outE = dalloc<uint32_t>(ctx, nout);
  if (!p.outS || !p.outE) return BK_ERR_NOMEM;
  const uint64_t blocks = (n + SEG_THREADS - 1) / SEG_THREADS, cap = (uint64_t)kSMs * 32;
  prof_begin(ctx, diff ? "k_difference" : "k_intersect");
  if (diff) k_intersect<true><<<(unsigned)(blocks < cap ? blocks : cap), SEG_THREADS, 0, ctx->stream>>>(p);
  else k_intersect<false><<<(unsigned)(blocks < cap ? blocks : cap), SEG_THREADS, 0, ctx->stream>>>(p);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  std::vector<uint64_t> rob(nruns);
  BK_CUDA(ctx, cudaMemcpyAsync(rob.data(), p.run_out_begin, (size_t)nruns * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dfree(ctx, (void*)p.run_a_begin); dfree(ctx, (void*)p.run_b_begin); dfree(ctx, (void*)p.run_b_end);
  dfree(ctx, p.run_out_begin); dfree(ctx, p.local); dfree(ctx, p.range_total); dfree(ctx, d_base);
  out->s = p.outS; out->e = p.outE; out->n = nout;
  for (int r = 0; r < nruns; r++) out->runs.push_back({names[r], rob[r], r + 1 < nruns ? rob[r + 1] : nout});
  return BK_OK;
}

// ---- complement: the gaps between consecutive segments of a merged list (nextComplementLine, Bedops.cpp:891-943) ----
struct ComplParams {
  const uint32_t* us;
  const uint32_t* ue;
  const uint64_t* seg_begin;  // [nruns] first segment of the run
  const uint64_t* out_begin;  // [nruns+1] first output row of the run
  const uint8_t*  has_left;   // [nruns] -L and the run's first segment does not start at 0: one more row, [0, first start)
  int             nruns;
  uint64_t        nout;
  uint32_t*       outS;
  uint32_t*       outE;
};
__global__ void k_complement(ComplParams p) {
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < p.nout; o += (uint64_t)gridDim.x * blockDim.x) {
    int l = 0, h = p.nruns;
    while (h - l > 1) {
      const int mid = (l + h) >> 1;
      if (p.out_begin[mid] <= o) l = mid; else h = mid;
    }
    uint64_t       j = o - p.out_begin[l];
    const uint64_t b = p.seg_begin[l];
    if (p.has_left[l]) {
      if (j == 0) {
        p.outS[o] = 0;
        p.outE[o] = p.us[b];
        continue;
      }
      j--;
    }
    p.outS[o] = p.ue[b + j];
    p.outE[o] = p.us[b + j + 1];
  }
}

static int complement_of(bk_ctx* ctx, const IvList& U, bool full_left, IvList* out) {
  out->runs.clear();
  out->n = 0;
  out->owned = true;
  out->s = out->e = nullptr;
  std::vector<uint64_t>    sb, ob;
  std::vector<uint8_t>     left;
  std::vector<std::string> names;
  std::vector<uint32_t>    first_start;
  for (auto& r : U.runs)
    if (r.row_end > r.row_begin) {
      sb.push_back(r.row_begin);
      names.push_back(r.name);
    }
  const int nruns = (int)sb.size();
  if (nruns == 0) return BK_OK;
  first_start.assign(nruns, 1);
  if (full_left) {
    for (int g = 0; g < nruns; g++)
      BK_CUDA(ctx, cudaMemcpyAsync(&first_start[g], U.s + sb[g], 4, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  uint64_t nout = 0;
  int      g = 0;
  for (auto& r : U.runs) {
    if (r.row_end == r.row_begin) continue;
    const bool l = full_left && first_start[g] != 0;
    left.push_back(l ? 1 : 0);
    ob.push_back(nout);
    nout += (r.row_end - r.row_begin - 1) + (l ? 1 : 0);
    g++;
  }
  ob.push_back(nout);
  if (nout == 0) return BK_OK;
  ComplParams p{};
  p.us = U.s; p.ue = U.e; p.nruns = nruns; p.nout = nout;
  p.seg_begin = upload(ctx, sb); p.out_begin = upload(ctx, ob); p.has_left = upload(ctx, left);
  p.outS = dalloc<uint32_t>(ctx, nout); p.outE = dalloc<uint32_t>(ctx, nout);
  if (!p.seg_begin || !p.out_begin || !p.has_left || !p.outS || !p.outE) return BK_ERR_NOMEM;
  const uint64_t blocks = (nout + 255) / 256, cap = (uint64_t)kSMs * 16;
  prof_begin(ctx, "k_complement");
  k_complement<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(p);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the uploaded host vectors go out of scope
  dfree(ctx, (void*)p.seg_begin); dfree(ctx, (void*)p.out_begin); dfree(ctx, (void*)p.has_left);
  out->s = p.outS; out->e = p.outE; out->n = nout;
  for (int r = 0; r < nruns; r++) out->runs.push_back({names[r], ob[r], ob[r + 1]});
  return BK_OK;
}

// ---- --chop: merged segments cut into fixed-size pieces (doChop, Bedops.cpp:438-467) ------------------------------
// Pieces per segment are a closed form; an exclusive scan over the segments (warp ranges, one-CTA scan of the range
// totals, no carry chain) gives every segment its first output row, and every OUTPUT row then finds its segment by
// bisection -- a chromosome-long segment chopped into single bases is still spread over the whole grid.
struct ChopParams {
  const uint32_t* us;
  const uint32_t* ue;
  uint64_t        nseg;
  uint32_t        chunk, step;
  int             exclude_short;
  uint32_t*       local;       // [nseg] pieces of the earlier segments of the same warp range
  uint64_t*       range_total; // [nranges]
  const uint64_t* range_base;  // [nranges+1]
  uint64_t        nranges;
  uint64_t        nout;
  uint32_t*       outS;
  uint32_t*       outE;
};
constexpr int CHOP_RANGE = 1024;  // segments per warp range

__device__ __forceinline__ uint64_t chop_pieces(uint32_t s, uint32_t e, uint32_t chunk, uint32_t step, int exclude_short) {
  const uint64_t len = (uint64_t)e - s;
  if (!exclude_short) return (len + step - 1) / step;     // i = s, s+step, ... while i < e
  return len >= chunk ? (len - chunk) / step + 1 : 0;    // ... while i + chunk <= e
}

__global__ void __launch_bounds__(256) k_chop_count(ChopParams p) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * 256) >> 5;
  for (uint64_t r = w0; r < p.nranges; r += nw) {
    const uint64_t a = r * CHOP_RANGE, b = a + CHOP_RANGE < p.nseg ? a + CHOP_RANGE : p.nseg;
    uint64_t       run = 0;
    for (uint64_t k0 = a; k0 < b; k0 += 32) {
      const uint64_t k = k0 + lane;
      uint64_t       c = k < b ? chop_pieces(p.us[k], p.ue[k], p.chunk, p.step, p.exclude_short) : 0;
      uint64_t       incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
      }
      // a range holds at most 2^32-1 pieces (checked on the host from the totals): 32-bit local prefixes
      if (k < b) p.local[k] = (uint32_t)(run + incl - c);
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) p.range_total[r] = run;
  }
}

__global__ void __launch_bounds__(256) k_chop_write(ChopParams p) {
  for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < p.nout; o += (uint64_t)gridDim.x * blockDim.x) {
    // last range with base <= o, then last segment of it with first-row <= o
    uint64_t lo = 0, hi = p.nranges;
    while (hi - lo > 1) {
      const uint64_t mid = (lo + hi) >> 1;
      if (p.range_base[mid] <= o) lo = mid; else hi = mid;
    }
    const uint64_t base = p.range_base[lo];
    uint64_t       a = lo * CHOP_RANGE, b = a + CHOP_RANGE < p.nseg ? a + CHOP_RANGE : p.nseg;
    while (b - a > 1) {
      const uint64_t mid = (a + b) >> 1;
      if (base + p.local[mid] <= o) a = mid; else b = mid;
    }
    const uint64_t j = o - (base + p.local[a]);
    const uint32_t s = p.us[a], e = p.ue[a];
    const uint64_t ps = (uint64_t)s + j * p.step, pe = ps + p.chunk;
    p.outS[o] = (uint32_t)ps;
    p.outE[o] = pe > e ? e : (uint32_t)pe;
  }
}

static int chop_list(bk_ctx* ctx, const IvList& U, uint32_t chunk, uint32_t stagger, bool exclude_short, IvList* out) {
  out->runs.clear();
  out->n = 0;
  out->owned = true;
  out->s = out->e = nullptr;
  if (U.n == 0) return BK_OK;
  ChopParams p{};
  p.us = U.s; p.ue = U.e; p.nseg = U.n; p.chunk = chunk; p.step = stagger ? stagger : chunk; p.exclude_short = exclude_short;
  p.nranges = (U.n + CHOP_RANGE - 1) / CHOP_RANGE;
  if (p.nranges >> 32) return fail(ctx, BK_ERR_UNSUPPORTED, "too many segments to chop");
  p.local = dalloc<uint32_t>(ctx, U.n);
  p.range_total = dalloc<uint64_t>(ctx, p.nranges);
  uint64_t* base = dalloc<uint64_t>(ctx, p.nranges + 1);
  if (!p.local || !p.range_total || !base) return BK_ERR_NOMEM;
  p.range_base = base;
  BK_TRY(reset_scratch(ctx));
  const uint64_t want = (p.nranges + 7) / 8;
  prof_begin(ctx, "k_chop_count");
  k_chop_count<<<grid_for_kernel((const void*)k_chop_count, 256, want), 256, 0, ctx->stream>>>(p);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  k_scan_totals<SC_OUT_ROWS><<<1, 1024, 0, ctx->stream>>>(p.range_total, base, (uint32_t)p.nranges, ctx->d_scratch);
  BK_LAUNCHED(ctx);
  // first output row of every chromosome run = first row of its first segment
  std::vector<uint64_t> h_base(p.nranges + 1);
  BK_CUDA(ctx, cudaMemcpyAsync(h_base.data(), base, (p.nranges + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BK_TRY(read_scratch(ctx));
  p.nout = ctx->h_scratch[SC_OUT_ROWS];
  for (uint64_t r = 0; r < p.nranges; r++)
    if (h_base[r + 1] - h_base[r] > 0xFFFFFFFFull) {
      dfree(ctx, p.local); dfree(ctx, p.range_total); dfree(ctx, base);
      return fail(ctx, BK_ERR_UNSUPPORTED, "--chop output too fine for the device writer (more than 2^32 pieces per 1024 segments)");
    }
  std::vector<uint64_t> run_first;
  std::vector<std::string> names;
  for (auto& r : U.runs) {
    if (r.row_end == r.row_begin) continue;
    uint32_t l = 0;
    BK_CUDA(ctx, cudaMemcpyAsync(&l, p.local + r.row_begin, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    run_first.push_back(h_base[r.row_begin / CHOP_RANGE] + l);
    names.push_back(r.name);
  }
  if (p.nout) {
    p.outS = dalloc<uint32_t>(ctx, p.nout);
    p.outE = dalloc<uint32_t>(ctx, p.nout);
    if (!p.outS || !p.outE) return BK_ERR_NOMEM;
    const uint64_t blocks = (p.nout + 255) / 256, cap = (uint64_t)kSMs * 32;
    prof_begin(ctx, "k_chop_write");
    k_chop_write<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(p);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  dfree(ctx, p.local); dfree(ctx, p.range_total); dfree(ctx, base);
  out->s = p.outS; out->e = p.outE; out->n = p.nout;
  for (size_t g = 0; g < names.size(); g++)
    out->runs.push_back({names[g], run_first[g], g + 1 < names.size() ? run_first[g + 1] : p.nout});
  return BK_OK;
}

// ---- --everything: multiset union of whole rows (doUnionAll / nextUnionAllLine, Bedops.cpp:761-786, :1468-1516) ----
// Order: chromosome, start, end, strcmp of the rest of the line, file number.  Every row finds its output slot by
// ranking itself in every other file (binary searches with that comparator: the inputs are sorted per sort-bed, which
// orders by the rest of the line too); the emitter then echoes row src[slot].
constexpr int kMaxEverything = 32;
struct EverythingFiles {
  const char*     text[kMaxEverything];
  const uint64_t* line[kMaxEverything];
  const uint32_t* s[kMaxEverything];
  const uint32_t* e[kMaxEverything];
};

// the rest of a line after its third field (what %[^\n] of B3Rest::readline keeps; strtoul skips blanks and a '+')
__device__ __forceinline__ const char* rest_of_line(const char* p) {
  while (is_tok((unsigned char)*p)) p++;
  for (int f = 0; f < 2; f++) {
    while (is_ws((unsigned char)*p)) p++;
    if (*p == '+') p++;
    while (is_digit((unsigned char)*p)) p++;
  }
  return p;
}
__device__ __forceinline__ int cmp_rest(const char* a, const char* b) {  // strcmp of two '\n'-terminated strings
  while (true) {
    const unsigned char x = *a == '\n' ? 0 : (unsigned char)*a, y = *b == '\n' ? 0 : (unsigned char)*b;
    if (x != y) return x < y ? -1 : 1;
    if (x == 0) return 0;
    a++;
    b++;
  }
}

struct RankRowsParams {
  EverythingFiles F;
  int             k, f;
  const uint64_t* run_begin;   // [nruns] first participating row of each selected run of file f
  const uint32_t* run_g;       // [nruns] global chromosome index
  const uint64_t* sel_prefix;  // [nruns+1]
  int             nruns;
  uint64_t        nsel;
  const uint64_t* g_begin;     // [G*k]
  const uint64_t* g_end;
  const uint64_t* g_out;       // [G]
  uint64_t*       src;         // [N] (file << 56) | row
};

__global__ void __launch_bounds__(256) k_rank_rows(RankRowsParams p) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.nsel; t += stride) {
    int lo = 0, hi = p.nruns;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.sel_prefix[mid] <= t) lo = mid; else hi = mid;
    }
    const uint64_t row = p.run_begin[lo] + (t - p.sel_prefix[lo]);
    const uint32_t g = p.run_g[lo];
    const uint32_t ks = p.F.s[p.f][row], ke = p.F.e[p.f][row];
    const char*    krest = nullptr;  // found lazily: only ties on (start, end) need it
    uint64_t       pos = p.g_out[g] + (row - p.g_begin[(uint64_t)g * p.k + p.f]);
    for (int j = 0; j < p.k; j++) {
      if (j == p.f) continue;
      uint64_t b = p.g_begin[(uint64_t)g * p.k + j], e = p.g_end[(uint64_t)g * p.k + j];
      if (b == e) continue;
      const uint64_t b0 = b;
      // rows of file j that precede this row: key(j,r) < key for a later file, key(j,r) <= key for an earlier one
      while (b < e) {
        const uint64_t mid = b + ((e - b) >> 1);
        const uint32_t ms = p.F.s[j][mid], me = p.F.e[j][mid];
        int            c;  // key(j,mid) vs key
        if (ms != ks) c = ms < ks ? -1 : 1;
        else if (me != ke) c = me < ke ? -1 : 1;
        else {
          if (!krest) krest = rest_of_line(p.F.text[p.f] + (p.F.line[p.f][row] & kLineOffMask));
          c = cmp_rest(rest_of_line(p.F.text[j] + (p.F.line[j][mid] & kLineOffMask)), krest);
        }
        const bool before = j < p.f ? c <= 0 : c < 0;
        if (before) b = mid + 1; else e = mid;
      }
      pos += b - b0;
    }
    p.src[pos] = ((uint64_t)p.f << 56) | row;
  }
}

struct EverythingRow {
  EverythingFiles F;
  const uint64_t* src;
  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& sk) const {
    const uint64_t v = src[i], row = v & ((1ull << 56) - 1);
    const int      f = (int)(v >> 56);
    echo_b3rest(sk, F.text[f], F.line[f][row], F.s[f][row], F.e[f][row]);
    sk.put('\n');
  }
};

static int everything(bk_ctx* ctx, const bk_bed* const* files, int k, const char* chrom, int on_device, bk_text* out) {
  if (k > kMaxEverything) return fail(ctx, BK_ERR_UNSUPPORTED, "--everything over more than %d files", kMaxEverything);
  EverythingFiles F{};
  std::vector<IvList> in;
  for (int f = 0; f < k; f++) {
    if (!files[f]->line_off && files[f]->nrows) return fail(ctx, BK_ERR_ARG, "--everything needs every file loaded with BK_COL_LINE");
    if (k > 1 && files[f]->pad_tie_disorder)
      return fail(ctx, BK_ERR_UNSUPPORTED,
                  "--range clamped rows of different starts onto equal coordinates in file %d; the order in which the reference merges "
                  "such rows across files depends on its reader state and is not reproduced", f + 1);
    F.text[f] = files[f]->d_text; F.line[f] = files[f]->line_off; F.s[f] = files[f]->start; F.e[f] = files[f]->end;
    in.push_back(view_of(files[f], chrom));
  }
  std::vector<std::string> names;
  for (auto& l : in)
    for (auto& r : l.runs)
      if (r.row_end > r.row_begin) names.push_back(r.name);
  std::sort(names.begin(), names.end(), [](const std::string& a, const std::string& b) { return strcmp(a.c_str(), b.c_str()) < 0; });
  names.erase(std::unique(names.begin(), names.end()), names.end());
  const int G = (int)names.size();
  if (G == 0) return finish_text(ctx, nullptr, 0, 0, on_device, out);
  std::map<std::string, int> gidx;
  for (int g = 0; g < G; g++) gidx[names[g]] = g;
  std::vector<uint64_t> g_begin((size_t)G * k, 0), g_end((size_t)G * k, 0), g_out(G + 1, 0);
  for (int f = 0; f < k; f++)
    for (auto& r : in[f].runs) {
      if (r.row_end == r.row_begin) continue;
      const int g = gidx[r.name];
      g_begin[(size_t)g * k + f] = r.row_begin;
      g_end[(size_t)g * k + f] = r.row_end;
    }
  for (int g = 0; g < G; g++) {
    uint64_t c = 0;
    for (int f = 0; f < k; f++) c += g_end[(size_t)g * k + f] - g_begin[(size_t)g * k + f];
    g_out[g + 1] = g_out[g] + c;
  }
  const uint64_t N = g_out[G];
  uint64_t*      src = dalloc<uint64_t>(ctx, N);
  uint64_t *d_gb = upload(ctx, g_begin), *d_ge = upload(ctx, g_end), *d_go = upload(ctx, g_out);
  if (!src || !d_gb || !d_ge || !d_go) return BK_ERR_NOMEM;
  std::vector<void*> tofree{d_gb, d_ge, d_go};
  std::vector<std::vector<uint64_t>> keep_rb(k), keep_sp(k);
  std::vector<std::vector<uint32_t>> keep_rg(k);
  for (int f = 0; f < k; f++) {
    auto&    rb = keep_rb[f];
    auto&    sp = keep_sp[f];
    auto&    rg = keep_rg[f];
    uint64_t acc = 0;
    for (auto& r : in[f].runs) {
      if (r.row_end == r.row_begin) continue;
      rb.push_back(r.row_begin);
      rg.push_back((uint32_t)gidx[r.name]);
      sp.push_back(acc);
      acc += r.row_end - r.row_begin;
    }
    if (acc == 0) continue;
    sp.push_back(acc);
    RankRowsParams p{};
    p.F = F; p.k = k; p.f = f; p.nruns = (int)rg.size(); p.nsel = acc;
    p.run_begin = upload(ctx, rb); p.run_g = upload(ctx, rg); p.sel_prefix = upload(ctx, sp);
    if (!p.run_begin || !p.run_g || !p.sel_prefix) return BK_ERR_NOMEM;
    tofree.push_back((void*)p.run_begin); tofree.push_back((void*)p.run_g); tofree.push_back((void*)p.sel_prefix);
    p.g_begin = d_gb; p.g_end = d_ge; p.g_out = d_go; p.src = src;
    const uint64_t blocks = (acc + 255) / 256, cap = (uint64_t)kSMs * 32;
    prof_begin(ctx, "k_rank_rows");
    k_rank_rows<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(p);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
  }
  EverythingRow fn{};
  fn.F = F; fn.src = src;
  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  int      rc = run_emit(ctx, fn, N, 0, &d_out, &bytes, &rows);  // run_emit syncs: the uploads are done
  for (void* q : tofree) dfree(ctx, q);
  dfree(ctx, src);
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return rc;
  }
  return finish_text(ctx, d_out, bytes, rows, on_device, out);
}

static int emit_bed3(bk_ctx* ctx, const IvList& l, int on_device, bk_text* out) {
  std::vector<char>     names;
  std::vector<uint32_t> lens;
  std::vector<uint64_t> rb;
  uint64_t              maxname = 0;
  for (auto& r : l.runs) {
    if (r.row_end == r.row_begin) continue;
    size_t o = names.size();
    names.resize(o + 128, 0);
    memcpy(&names[o], r.name.c_str(), r.name.size());
    lens.push_back((uint32_t)r.name.size());
    rb.push_back(r.row_begin);
    maxname = std::max<uint64_t>(maxname, r.name.size());
  }
  if (l.n == 0 || rb.empty()) return finish_text(ctx, nullptr, 0, 0, on_device, out);
  Bed3Row fn{};
  fn.s = l.s; fn.e = l.e;
  fn.nt.names = upload(ctx, names); fn.nt.len = upload(ctx, lens); fn.nt.run_begin = upload(ctx, rb); fn.nt.nruns = (int)rb.size();
  if (!fn.nt.names || !fn.nt.len || !fn.nt.run_begin) return BK_ERR_NOMEM;
  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  // rows before the first non-empty run cannot exist (lists start at their first run)
  int rc = run_emit(ctx, fn, l.n, l.n * (maxname + 24) + 64, &d_out, &bytes, &rows);
  dfree(ctx, (void*)fn.nt.names); dfree(ctx, (void*)fn.nt.len); dfree(ctx, (void*)fn.nt.run_begin);
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return rc;
  }
  return finish_text(ctx, d_out, bytes, rows, on_device, out);
}

// ---- --partition ------------------------------------------------------------------------------------------
// doPartitions / nextPartitionGroup (Bedops.cpp:615-653, :1249-1335): every start and every end of every input row is a
// break point; the output is the pieces between consecutive distinct break points of a chromosome that at least one
// input row covers (touching rows stay apart: their common coordinate is a break point; duplicates give one piece).
// Device form: the ends are not sorted in the input (nesting), so all 2N coordinates are radix-sorted as
// (chromosome rank << 32 | coordinate); a piece [key j, key j+1) is printed when the two keys differ, share the
// chromosome, and key j lies inside a segment of the merged union of all files (whose borders are break points too).
struct PartKeyParams {
  const uint32_t* s;
  const uint32_t* e;
  const uint64_t* run_begin;  // [nruns+1] rows of the participating runs of this file (contiguous, ascending)
  const uint32_t* run_rank;   // [nruns] rank of the run's chromosome among all names
  int             nruns;
  uint64_t        row0, n;    // participating rows [row0, row0+n)
};
__global__ void __launch_bounds__(256) k_part_keys(PartKeyParams p, uint64_t* __restrict__ keys) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const uint64_t row = p.row0 + i;
  int            l = 0, h = p.nruns;
  while (h - l > 1) {
    const int mid = (l + h) >> 1;
    if (p.run_begin[mid] <= row) l = mid; else h = mid;
  }
  const uint64_t r = (uint64_t)p.run_rank[l] << 32;
  keys[2 * i] = r | p.s[row];
  keys[2 * i + 1] = r | p.e[row];
}

struct PartRow {
  const uint64_t* keys;
  uint64_t        n;
  const uint32_t* us;        // merged union of all files
  const uint32_t* ue;
  const uint64_t* u_begin;   // [nranks] rows of the union per chromosome rank
  const uint64_t* u_end;
  const char*     names;     // [nranks][128]
  const uint32_t* name_len;
  template <class Sink>
  __device__ void operator()(uint64_t j, Sink& sk) const {
    if (j + 1 >= n) return;
    const uint64_t a = keys[j], b = keys[j + 1];
    if (a == b || (a >> 32) != (b >> 32)) return;
    const uint32_t r = (uint32_t)(a >> 32), x = (uint32_t)a;
    const uint64_t ub = u_begin[r], uend = u_end[r];
    const uint64_t hi = upper_bound_u32(us, ub, uend, x);  // first segment that starts after x
    if (hi == ub || x >= ue[hi - 1]) return;              // x lies in no segment: a gap between rows
    sk.puts_(names + (size_t)r * 128, (int)name_len[r]);
    sk.put('\t');
    sk.put_u32(x);
    sk.put('\t');
    sk.put_u32((uint32_t)b);
    sk.put('\n');
  }
};

static int partition(bk_ctx* ctx, const bk_bed* const* files, int k, const char* chrom, int on_device, bk_text* out) {
  std::vector<IvList> in;
  for (int i = 0; i < k; i++) in.push_back(view_of(files[i], chrom));
  std::vector<std::string> names;
  uint64_t                 total = 0;
  for (auto& l : in)
    for (auto& r : l.runs)
      if (r.row_end > r.row_begin) {
        names.push_back(r.name);
        total += r.row_end - r.row_begin;
      }
  if (total == 0) return finish_text(ctx, nullptr, 0, 0, on_device, out);
  auto less = [](const std::string& a, const std::string& b) { return strcmp(a.c_str(), b.c_str()) < 0; };
  std::sort(names.begin(), names.end(), less);
  names.erase(std::unique(names.begin(), names.end()), names.end());
  auto rank_of = [&](const std::string& nm) { return (uint32_t)(std::lower_bound(names.begin(), names.end(), nm, less) - names.begin()); };
  const uint64_t n2 = 2 * total;
  uint64_t*      keys = dalloc<uint64_t>(ctx, n2);
  uint64_t*      keys2 = dalloc<uint64_t>(ctx, n2);
  if (!keys || !keys2) return BK_ERR_NOMEM;
  std::vector<void*> tmp;
  uint64_t           at = 0;
  for (int i = 0; i < k; i++) {
    std::vector<uint64_t> rb;
    std::vector<uint32_t> rr;
    for (auto& r : in[i].runs)
      if (r.row_end > r.row_begin) {
        rb.push_back(r.row_begin);
        rr.push_back(rank_of(r.name));
      }
    if (rb.empty()) continue;
    uint64_t row1 = 0;
    for (auto& r : in[i].runs)
      if (r.row_end > r.row_begin) row1 = r.row_end;
    rb.push_back(row1);
    PartKeyParams p{};
    p.s = in[i].s; p.e = in[i].e; p.nruns = (int)rr.size(); p.row0 = rb[0]; p.n = row1 - rb[0];
    p.run_begin = upload(ctx, rb); p.run_rank = upload(ctx, rr);
    if (!p.run_begin || !p.run_rank) return BK_ERR_NOMEM;
    tmp.push_back((void*)p.run_begin); tmp.push_back((void*)p.run_rank);
    prof_begin(ctx, "k_part_keys");
    k_part_keys<<<(unsigned)((p.n + 255) / 256), 256, 0, ctx->stream>>>(p, keys + at);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // rb / rr (host vectors) must outlive their copies
    at += 2 * p.n;
  }
  int rank_bits = 1;
  while ((names.size() - 1) >> rank_bits) rank_bits++;
  int rc = radix_sort_pairs(ctx, &keys, nullptr, &keys2, nullptr, n2, 32 + rank_bits);
  IvList u;
  if (rc == BK_OK) rc = union_merge(ctx, in, &u);
  if (rc == BK_OK) {
    std::vector<uint64_t> ub(names.size(), 0), ue(names.size(), 0);
    for (auto& r : u.runs) {
      if (r.row_end == r.row_begin) continue;
      const uint32_t q = rank_of(r.name);
      ub[q] = r.row_begin;
      ue[q] = r.row_end;
    }
    std::vector<char>     nm(names.size() * 128, 0);
    std::vector<uint32_t> nl(names.size());
    for (size_t q = 0; q < names.size(); q++) {
      memcpy(&nm[q * 128], names[q].c_str(), names[q].size());
      nl[q] = (uint32_t)names[q].size();
    }
    PartRow fn{};
    fn.keys = keys; fn.n = n2; fn.us = u.s; fn.ue = u.e;
    fn.u_begin = upload(ctx, ub); fn.u_end = upload(ctx, ue); fn.names = upload(ctx, nm); fn.name_len = upload(ctx, nl);
    if (!fn.u_begin || !fn.u_end || !fn.names || !fn.name_len) rc = BK_ERR_NOMEM;
    char*    d_out = nullptr;
    uint64_t bytes = 0, rows = 0;
    if (rc == BK_OK) rc = run_emit(ctx, fn, n2, 0, &d_out, &bytes, &rows);  // run_emit syncs: the uploads are done
    dfree(ctx, (void*)fn.u_begin); dfree(ctx, (void*)fn.u_end); dfree(ctx, (void*)fn.names); dfree(ctx, (void*)fn.name_len);
    if (rc == BK_OK) rc = finish_text(ctx, d_out, bytes, rows, on_device, out);
    else dfree(ctx, d_out);
  }
  free_list(ctx, u);
  for (void* q : tmp) dfree(ctx, q);
  dfree(ctx, keys);
  dfree(ctx, keys2);
  return rc;
}

}  // namespace bk

using namespace bk;

extern "C" int bk_setop(bk_ctx* ctx, int op, const bk_bed* const* files, int n_files, double thr, int thr_is_pct,
                        const char* chrom, int out_on_device, bk_text* out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !files || !out || n_files < 1) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < n_files; i++)
    if (!files[i]) return BK_ERR_ARG;
  if (n_files > kMaxFiles) return fail(ctx, BK_ERR_UNSUPPORTED, "more than %d input files", kMaxFiles);

  if (op == BK_SETOP_MERGE) {
    std::vector<IvList> in;
    for (int i = 0; i < n_files; i++) in.push_back(view_of(files[i], chrom));
    IvList u;
    BK_TRY(union_merge(ctx, in, &u));
    int rc = emit_bed3(ctx, u, out_on_device, out);
    free_list(ctx, u);
    return rc;
  }
  if (op == BK_SETOP_INTERSECT) {
    if (n_files < 2) return fail(ctx, BK_ERR_ARG, "Not enough files");
    IvList acc;
    for (int i = 0; i < n_files; i++) {
      std::vector<IvList> one{view_of(files[i], chrom)};
      IvList m;
      BK_TRY(union_merge(ctx, one, &m));  // getNextFileMergedCoords: every file is merged within itself first
      if (i == 0) {
        acc = m;
      } else {
        IvList r;
        int    rc = intersect_pair(ctx, acc, m, &r);
        free_list(ctx, acc);
        free_list(ctx, m);
        if (rc != BK_OK) return rc;
        acc = r;
      }
    }
    int rc = emit_bed3(ctx, acc, out_on_device, out);
    free_list(ctx, acc);
    return rc;
  }
  if (op == BK_SETOP_EVERYTHING) return everything(ctx, files, n_files, chrom, out_on_device, out);
  if (op == BK_SETOP_PARTITION) return partition(ctx, files, n_files, chrom, out_on_device, out);
  if (op == BK_SETOP_COMPLEMENT) {
    std::vector<IvList> in;
    for (int i = 0; i < n_files; i++) in.push_back(view_of(files[i], chrom));
    IvList u, c;
    BK_TRY(union_merge(ctx, in, &u));
    int rc = complement_of(ctx, u, thr != 0.0, &c);
    free_list(ctx, u);
    if (rc == BK_OK) rc = emit_bed3(ctx, c, out_on_device, out);
    free_list(ctx, c);
    return rc;
  }
  if (op == BK_SETOP_DIFFERENCE) {
    if (n_files < 2) return fail(ctx, BK_ERR_ARG, "Not enough files");
    std::vector<IvList> first{view_of(files[0], chrom)}, rest;
    for (int i = 1; i < n_files; i++) rest.push_back(view_of(files[i], chrom));
    IvList a, b, d;
    BK_TRY(union_merge(ctx, first, &a));
    int rc = union_merge(ctx, rest, &b);
    if (rc == BK_OK) rc = intersect_pair(ctx, a, b, &d, true);
    free_list(ctx, a);
    free_list(ctx, b);
    if (rc == BK_OK) rc = emit_bed3(ctx, d, out_on_device, out);
    free_list(ctx, d);
    return rc;
  }
  if (op == BK_SETOP_SYMMDIFF) {
    if (n_files < 2) return fail(ctx, BK_ERR_ARG, "Not enough files");
    // bases covered by exactly one file = union of all files minus the union of the pairwise intersections
    std::vector<IvList> per(n_files), pairs;
    int                 rc = BK_OK;
    for (int i = 0; i < n_files && rc == BK_OK; i++) {
      std::vector<IvList> one{view_of(files[i], chrom)};
      rc = union_merge(ctx, one, &per[i]);
    }
    for (int i = 0; i < n_files && rc == BK_OK; i++)
      for (int j = i + 1; j < n_files && rc == BK_OK; j++) {
        pairs.emplace_back();
        rc = intersect_pair(ctx, per[i], per[j], &pairs.back());
      }
    IvList u, twice, d;
    if (rc == BK_OK) rc = union_merge(ctx, per, &u);
    if (rc == BK_OK) rc = union_merge(ctx, pairs, &twice);
    if (rc == BK_OK) rc = intersect_pair(ctx, u, twice, &d, true);
    for (auto& l : per) free_list(ctx, l);
    for (auto& l : pairs) free_list(ctx, l);
    free_list(ctx, u);
    free_list(ctx, twice);
    if (rc == BK_OK) rc = emit_bed3(ctx, d, out_on_device, out);
    free_list(ctx, d);
    return rc;
  }
  if (op == BK_SETOP_ELEMENT_OF || op == BK_SETOP_NOT_ELEMENT_OF) {
    if (n_files < 2) return fail(ctx, BK_ERR_ARG, "Not enough files");
    const bk_bed* ref = files[0];
    if (!ref->line_off && ref->nrows) return fail(ctx, BK_ERR_ARG, "reference file was loaded without BK_COL_LINE");
    std::vector<IvList> in;
    for (int i = 1; i < n_files; i++) in.push_back(view_of(files[i], chrom));
    IvList u;
    BK_TRY(union_merge(ctx, in, &u));
    IvList rv = view_of(ref, chrom);
    std::vector<uint64_t> rrb, ub, ue;
    std::vector<uint8_t>  later;
    std::string           last_u;
    for (auto& r : u.runs)
      if (r.row_end > r.row_begin) last_u = r.name;
    uint64_t row0 = 0, row1 = 0;
    bool     first = true;
    size_t   j = 0;
    for (auto& r : rv.runs) {
      if (r.row_end == r.row_begin) continue;
      if (first) { row0 = r.row_begin; first = false; }
      row1 = r.row_end;
      rrb.push_back(r.row_begin);
      later.push_back(!last_u.empty() && strcmp(r.name.c_str(), last_u.c_str()) < 0);
      while (j < u.runs.size() && strcmp(u.runs[j].name.c_str(), r.name.c_str()) < 0) j++;
      if (j < u.runs.size() && u.runs[j].name == r.name) {
        ub.push_back(u.runs[j].row_begin);
        ue.push_back(u.runs[j].row_end);
      } else {
        ub.push_back(0);
        ue.push_back(0);
      }
    }
    const uint64_t n = row1 - row0;
    if (n == 0) {
      free_list(ctx, u);
      return finish_text(ctx, nullptr, 0, 0, out_on_device, out);
    }
    rrb.push_back(row1);
    ElemParams p{};
    p.rs = ref->start; p.re = ref->end; p.row0 = row0; p.n = n;
    p.run_ref_begin = upload(ctx, rrb); p.run_u_begin = upload(ctx, ub); p.run_u_end = upload(ctx, ue); p.nruns = (int)ub.size();
    p.us = u.s; p.ue = u.e; p.thr = thr; p.use_pct = thr_is_pct; p.invert = op == BK_SETOP_NOT_ELEMENT_OF;
    p.keep = dalloc<uint8_t>(ctx, n);
    p.run_has_later = upload(ctx, later);
    if (!p.run_ref_begin || !p.run_u_begin || !p.run_u_end || !p.keep || !p.run_has_later) return BK_ERR_NOMEM;
    uint64_t blocks = (n + 255) / 256, cap = (uint64_t)kSMs * 32;
    prof_begin(ctx, "k_element_of");
    k_element_of<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(p);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
    EchoKeptRow fn{};
    fn.text = ref->d_text; fn.line = ref->line_off; fn.s = ref->start; fn.e = ref->end; fn.keep = p.keep; fn.row0 = row0;
    char*    d_out = nullptr;
    uint64_t bytes = 0, rows = 0;
    int rc = run_emit(ctx, fn, n, ref->nbytes + n * 24 + 64, &d_out, &bytes, &rows);  // run_emit syncs: uploads done
    dfree(ctx, (void*)p.run_ref_begin); dfree(ctx, (void*)p.run_u_begin); dfree(ctx, (void*)p.run_u_end); dfree(ctx, p.keep);
    dfree(ctx, (void*)p.run_has_later);
    free_list(ctx, u);
    if (rc != BK_OK) {
      dfree(ctx, d_out);
      return rc;
    }
    return finish_text(ctx, d_out, bytes, rows, out_on_device, out);
  }
  return fail(ctx, BK_ERR_UNSUPPORTED, "bedops operation %d is outside the device hot path", op);
}

extern "C" int bk_chop(bk_ctx* ctx, const bk_bed* const* files, int n_files, uint64_t chunk, uint64_t stagger,
                       int exclude_short, const char* chrom, int out_on_device, bk_text* out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !files || !out || n_files < 1) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < n_files; i++)
    if (!files[i]) return BK_ERR_ARG;
  if (chunk == 0 || chunk > 0xFFFFFFFFull || stagger > 0xFFFFFFFFull) return fail(ctx, BK_ERR_ARG, "bp setting for chop must be > 0");
  std::vector<IvList> in;
  for (int i = 0; i < n_files; i++) in.push_back(view_of(files[i], chrom));
  IvList u, c;
  BK_TRY(union_merge(ctx, in, &u));
  int rc = chop_list(ctx, u, (uint32_t)chunk, (uint32_t)stagger, exclude_short != 0, &c);
  free_list(ctx, u);
  if (rc == BK_OK) rc = emit_bed3(ctx, c, out_on_device, out);
  free_list(ctx, c);
  return rc;
}
