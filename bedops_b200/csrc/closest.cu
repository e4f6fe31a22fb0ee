// closest.cu -- closest-features: nearest left / right query element per reference row.
//
// Replaces findDistances (applications/bed/closestfeats/src/ClosestFeature.cpp:260-413).  The reference's answer is
// a function of its streaming state, not only of the two files: query elements it drops are gone for every later
// reference row, and elements held as left/right re-enter the stream behind the ones pushed back while they were held
// (BedReader::PushBack, BedReader.hpp:64-66), so with nested or overlapping reference rows the stream is neither
// complete nor in file order.  Byte parity therefore needs the state machine itself.  It is sequential per chromosome,
// so it is run speculatively in parallel:
//   k_cf_sim    one thread per chunk of CF_CHUNK consecutive reference rows of one chromosome.  A chunk that does not
//               begin its chromosome warms up over the CF_WARM rows in front of it from a guessed state (the best
//               non-overlapping left of the warm-up row + the unread file from the row's candidate window), records
//               the stream it ASSUMED at its first row, simulates its rows, records the stream it LEFT behind.
//   k_cf_check  chunk k is consistent iff the stream it assumed equals, element by element, the stream chunk k-1 left.
//   k_cf_sim    (rerun mode) every inconsistent chunk starts again from the stream chunk k-1 left; check; repeat.
// The first chunk of a chromosome starts from the exact state (empty push-back, file at the chromosome's first row), so
// by induction round r makes the first r chunks of every chromosome exact, and the loop ends when every chunk ran from
// the true stream: no approximation anywhere.  In practice streams converge within a chunk or two (measured on the
// 50 M x 200 M configuration: a third of the chunks rerun once, then a handful for 3-6 more rounds).
#include <algorithm>
#include "common.cuh"
#include "emit.cuh"
#include "fmt.cuh"
#include "parse.cuh"

namespace bk {

constexpr uint32_t kNoRow = 0xFFFFFFFFu;
constexpr int      CF_THREADS = 128;
constexpr int      CF_CHUNK = 256;      // reference rows per chunk
constexpr int      CF_WARM = 96;        // warm-up rows in front of a speculative chunk
constexpr uint64_t kExactState = ~0ull;  // "assumed" marker of a chunk that begins its chromosome

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

struct CfRun {
  uint64_t ref_begin, ref_end;  // reference rows of the chromosome
  uint64_t q_begin, q_end;      // query rows of the same chromosome (empty if absent)
  uint64_t first_chunk;
};

struct CfParams {
  const uint32_t* rs;
  const uint32_t* re;
  const CfRun*    runs;
  int             nruns;
  uint64_t        nchunks;
  uint64_t        row0;  // first reference row of the call (left/right are indexed by row - row0)
  const uint32_t* qs;
  const uint32_t* qe;
  const uint32_t* pm;   // running max of end within the run
  const uint32_t* pmi;  // 1 + (index within the run) of the last row attaining pm
  int             allow_overlaps;
  uint64_t*       left;  // global query row or UINT64_MAX
  uint64_t*       right;
  // per-thread stream buffers: two of `cap` entries each
  uint4*          bufs;
  uint32_t        cap;
  // recorded streams: arena[off] = file position, arena[off+1] = n, then n query rows (relative to the run)
  uint32_t*       arena;
  uint64_t        arena_cap;
  uint64_t*       arena_top;
  uint64_t*       assumed;   // [nchunks] arena offset of the stream the chunk started from (kExactState: exact)
  uint64_t*       fin_cur;   // [nchunks] arena offset of the stream the chunk left (read by rerun mode: chunk k-1)
  uint64_t*       fin_next;  // [nchunks] written by this launch
  const uint32_t* todo;      // chunks to run (null in the first speculative pass: all of them)
  uint32_t        ntodo;
  uint32_t*       ovf;       // chunks whose stream outgrew the buffers of this launch: run again with larger ones
  uint64_t*       ovf_count;
  uint64_t*       scratch;
};

__device__ __forceinline__ void cf_locate(const CfRun* __restrict__ runs, int nruns, uint64_t c, int& run, uint64_t& a, uint64_t& b) {
  int lo = 0, hi = nruns;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (runs[mid].first_chunk <= c) lo = mid; else hi = mid;
  }
  run = lo;
  a = runs[lo].ref_begin + (c - runs[lo].first_chunk) * CF_CHUNK;
  b = a + CF_CHUNK < runs[lo].ref_end ? a + CF_CHUNK : runs[lo].ref_end;
}

// The stream of one simulation: the push-back stack S[0..top) (BedReader::cache_, BedReader.hpp:76-84: the back is
// read next), then the file from row `pos`.  Entries carry the row's coordinates so that a step is one 16-byte load.
struct CfStream {
  uint4*   S;   // x = query row (relative to the run), y = start, z = end
  uint4*   RD;  // the push-back list of the current reference row (`read`, ClosestFeature.cpp:275), filled downwards
  uint32_t top, pos, cap;
  bool     overflow;
};

// Rows at the bottom of the stack that are exactly the file rows in front of `pos`, in file order, are the same stream
// as not having read them: how many can be handed back (the canonical form used to compare and to record states, and
// to keep the stack short behind a long reference row)
__device__ __forceinline__ uint32_t cf_strippable(const CfStream& st) {
  uint32_t b = 0;
  while (b < st.top && st.S[b].x + 1u + b == st.pos) b++;
  return b;
}

// one reference row of findDistances (:277-411); returns left/right as query rows relative to the run
__device__ __forceinline__ void cf_row(CfStream& st, const uint32_t* __restrict__ qs, const uint32_t* __restrict__ qe,
                                       uint32_t nq, uint32_t rs, uint32_t re, bool allow, uint32_t& out_left,
                                       uint32_t& out_right) {
  uint4    left = make_uint4(kNoRow, 0, 0, 0), right = make_uint4(kNoRow, 0, 0, 0);
  bool     left_ov = false, left_cached = false, drained = true;
  uint4* const   RD = st.RD;
  const uint32_t cap = st.cap;
  uint32_t       m = 0;  // RD[cap-1], RD[cap-2], ... RD[cap-m]: the list in push order
  const long long c2 = (long long)re - 1 + (long long)rs;  // 2 * getCentroid(b), :236-239
  auto push = [&](const uint4& v) { RD[cap - 1 - m++] = v; };
  while (true) {
    uint4 c;
    if (st.top) c = st.S[--st.top];
    else if (st.pos < nq) {
      c = make_uint4(st.pos, __ldg(&qs[st.pos]), __ldg(&qe[st.pos]), 0);
      st.pos++;
    } else break;
    if (m + 3 > cap) {  // every branch appends at most three entries
      st.overflow = true;
      return;
    }
    const uint32_t s = c.y, e = c.z;
    if (e <= rs) {  // left of the reference row (dist < 0)
      if (left.x == kNoRow || (!left_ov && e >= left.z)) {  // :300-309 new best left: the list so far is dropped
        m = 0;
        left = c; left_ov = false; left_cached = false;
      } else {  // :310-314 (c is dropped)
        if (!left_cached) push(left);
        left_cached = true;
      }
    } else if (s >= re) {  // right of the reference row (dist > 0): :315-331
      if (left.x != kNoRow && !left_cached) push(left);
      left_cached = left.x != kNoRow;
      if (right.x == kNoRow) right = c; else push(right);
      push(c);
      drained = false;
      break;
    } else if (allow) {  // overlap, :332-388
      if (s <= rs) {  // hangs over the left edge
        if (left.x != kNoRow && !left_cached && left.z > e) push(left);  // else dropped (:336-339)
        left = c; left_ov = true; left_cached = false;
      } else {
        // proportionOverlapLeft(c, centroid) >= 0.5 (:226-231) in exact integers: centroid >= s and
        // 2 * (centroid + 1 - s) >= e - s  (the correctly rounded double quotient compares the same way: 2x and the
        // length are integers below 2^53)
        const bool contained = e < re;
        const bool more_left = contained && c2 >= 2ll * s && c2 + 2 >= (long long)e + (long long)s;
        if (more_left && left_ov) {  // :362-368 left stays; c may be needed later
          if (!left_cached) push(left);
          left_cached = true;
          push(c);
        } else if (more_left) {  // :369-378 new left: the list so far is dropped
          m = 0;
          left = c; left_ov = true; left_cached = false;
        } else {  // hangs over the right edge (:343-350) or contained nearer the right edge (:354-361, :379-387)
          if (left.x != kNoRow && !left_cached) push(left);
          left_cached = left.x != kNoRow;
          if (right.x != kNoRow) push(right);
          right = c;
        }
      }
    } else {  // :389-397 --no-overlaps: kept for later reference rows
      if (left.x != kNoRow && !left_cached) {
        push(left);
        left_cached = true;
      }
      push(c);
    }
  }
  if (drained) {  // :403-406 the stream ran dry
    if (left.x != kNoRow && !left_cached) push(left);
    if (right.x != kNoRow) push(right);
  }
  // PushBack(read) (:407): the front of the list is read first = ends on top of the stack.  RD[cap-m .. cap) in address
  // order is the list reversed: a straight copy.
  if (st.top + m > cap) {  // make room: hand the file-order tail back to the file
    const uint32_t b = cf_strippable(st);
    if (b == 0 || st.top - b + m > cap) {
      st.overflow = true;
      return;
    }
    for (uint32_t k = b; k < st.top; k++) st.S[k - b] = st.S[k];
    st.top -= b;
    st.pos -= b;
  }
  for (uint32_t k = 0; k < m; k++) st.S[st.top + k] = RD[cap - m + k];
  st.top += m;
  out_left = left.x;
  out_right = right.x;
}

// recorded state (canonical form): arena[off] = file position, arena[off+1] = n, then the n stack rows, bottom first
__device__ __forceinline__ uint64_t cf_record(const CfParams& p, const CfStream& st) {
  const uint32_t b = cf_strippable(st), len = st.top - b;
  const uint64_t off = atomicAdd(reinterpret_cast<unsigned long long*>(p.arena_top), (unsigned long long)len + 2ull);
  if (off + len + 2 > p.arena_cap) {
    dev_set_error(p.scratch, BK_ERR_NOMEM, 1);
    return 0;
  }
  p.arena[off] = st.pos - b;
  p.arena[off + 1] = len;
  for (uint32_t k = 0; k < len; k++) p.arena[off + 2 + k] = st.S[b + k].x;
  return off;
}

// MODE 0: speculative run (warm-up from a guessed state).  MODE 1: run from the stream the predecessor left.
// Either over every chunk (todo == null) or over the chunks listed in todo[].
template <int MODE>
__global__ void __launch_bounds__(CF_THREADS) k_cf_sim(CfParams p) {
  const uint64_t t = (uint64_t)blockIdx.x * CF_THREADS + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * CF_THREADS;
  const uint64_t nwork = p.todo ? p.ntodo : p.nchunks;
  CfStream       st;
  st.cap = p.cap;
  st.S = p.bufs + t * 2ull * p.cap;
  st.RD = st.S + p.cap;
  for (uint64_t wk = t; wk < nwork; wk += nthreads) {
    const uint64_t c = p.todo ? p.todo[wk] : wk;
    int      run;
    uint64_t a, b;
    cf_locate(p.runs, p.nruns, c, run, a, b);
    const CfRun     R = p.runs[run];
    const uint32_t  nq = (uint32_t)(R.q_end - R.q_begin);
    const uint32_t* qs = p.qs + R.q_begin;
    const uint32_t* qe = p.qe + R.q_begin;
    st.top = 0; st.pos = 0; st.overflow = false;
    uint32_t l = kNoRow, r = kNoRow;
    if (MODE == 0) {
      uint64_t assumed = kExactState;
      if (a != R.ref_begin) {
        const uint64_t w = a - R.ref_begin > (uint64_t)CF_WARM ? a - CF_WARM : R.ref_begin;
        if (w != R.ref_begin && nq) {
          // guessed state at the warm-up row: its best non-overlapping left, then the file from its candidate window
          const uint32_t lo = (uint32_t)lower_bound_u32(p.pm + R.q_begin, 0, nq, (uint64_t)p.rs[w] + 1);
          if (lo > 0) {
            const uint32_t k = __ldg(&p.pmi[R.q_begin + lo - 1]) - 1u;
            st.S[st.top++] = make_uint4(k, qs[k], qe[k], 0);
          }
          st.pos = lo;
        }
        for (uint64_t row = w; row < a && !st.overflow; row++) cf_row(st, qs, qe, nq, p.rs[row], p.re[row], p.allow_overlaps, l, r);
        if (w != R.ref_begin && !st.overflow) assumed = cf_record(p, st);
      }
      p.assumed[c] = assumed;
    } else {
      const uint64_t off = p.fin_cur[c - 1];  // chunk c-1 is in the same chromosome (first chunks are never inconsistent)
      const uint32_t len = p.arena[off + 1];
      st.pos = p.arena[off];
      if (len > p.cap) st.overflow = true;
      else
        for (uint32_t k = 0; k < len; k++) {
          const uint32_t q = p.arena[off + 2 + k];
          st.S[k] = make_uint4(q, qs[q], qe[q], 0);
        }
      st.top = len;
      p.assumed[c] = off;
    }
    for (uint64_t row = a; row < b && !st.overflow; row++) {
      cf_row(st, qs, qe, nq, p.rs[row], p.re[row], p.allow_overlaps, l, r);
      p.left[row - p.row0] = l == kNoRow ? ~0ull : R.q_begin + l;
      p.right[row - p.row0] = r == kNoRow ? ~0ull : R.q_begin + r;
    }
    if (st.overflow) {
      p.ovf[atomicAdd(reinterpret_cast<unsigned long long*>(p.ovf_count), 1ull)] = (uint32_t)c;
      continue;
    }
    p.fin_next[c] = cf_record(p, st);
  }
}

// consistent(k) = the stream chunk k assumed == the stream chunk k-1 left (canonical forms, element by element).
// Inconsistent chunks are appended to todo[].
__global__ void k_cf_check(CfParams p, uint32_t* __restrict__ todo) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < p.nchunks; c += stride) {
    const uint64_t as = p.assumed[c];
    if (as == kExactState) continue;
    const uint64_t fo = p.fin_cur[c - 1];
    bool same = as == fo;
    if (!same) {
      const uint32_t n = p.arena[as + 1];
      same = p.arena[as] == p.arena[fo] && n == p.arena[fo + 1];
      for (uint32_t k = 0; same && k < n; k++) same = p.arena[as + 2 + k] == p.arena[fo + 2 + k];
    }
    if (same) continue;
    const uint64_t slot = atomicAdd(reinterpret_cast<unsigned long long*>(&p.scratch[SC_COUNT_A]), 1ull);
    todo[slot] = (uint32_t)c;
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
  bk::DeviceGuard device_guard(ctx);
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
  std::vector<CfRun> runs;
  uint64_t row0 = 0, row1 = 0, nchunks = 0;
  for (auto& r : ref->runs) {
    if (!all && r.name != spec->chrom) continue;
    if (r.row_end == r.row_begin) continue;
    if (runs.empty()) row0 = r.row_begin;
    row1 = r.row_end;
    CfRun c{r.row_begin, r.row_end, 0, 0, nchunks};
    for (auto& q : query->runs)
      if (q.name == r.name) { c.q_begin = q.row_begin; c.q_end = q.row_end; }
    nchunks += (r.row_end - r.row_begin + CF_CHUNK - 1) / CF_CHUNK;
    runs.push_back(c);
  }
  const uint64_t n = row1 - row0;
  if (n == 0) return finish_text(ctx, nullptr, 0, 0, spec->out_on_device, out);
  if (nchunks >= 0xFFFFFFFEull) return fail(ctx, BK_ERR_UNSUPPORTED, "reference file too large for the chunk index");
  const int nruns = (int)runs.size();

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
    if (rc != BK_OK) { dfree(ctx, d_pmi); return rc; }
  }

  CfParams p{};
  p.rs = ref->start; p.re = ref->end; p.nruns = nruns; p.nchunks = nchunks; p.row0 = row0;
  p.qs = query->start; p.qe = query->end; p.pm = query->pmax_end; p.pmi = d_pmi;
  p.allow_overlaps = !spec->no_overlaps;
  p.scratch = ctx->d_scratch;
  CfRun*    d_runs = dalloc<CfRun>(ctx, runs.size());
  uint64_t* d_chunk = dalloc<uint64_t>(ctx, 3 * nchunks + 1);  // assumed | fin A | fin B | arena top
  uint32_t* d_todo = dalloc<uint32_t>(ctx, nchunks + 1);
  p.left = dalloc<uint64_t>(ctx, n);
  p.right = dalloc<uint64_t>(ctx, n);
  std::vector<void*> temps{d_pmi, d_runs, d_chunk, d_todo};
  auto drop = [&](bool results) {
    for (void* t : temps) dfree(ctx, t);
    temps.clear();
    if (results) { dfree(ctx, p.left); dfree(ctx, p.right); }
  };
  if (!d_runs || !d_chunk || !d_todo || !p.left || !p.right) { drop(true); return BK_ERR_NOMEM; }
  BK_CUDA(ctx, cudaMemcpyAsync(d_runs, runs.data(), runs.size() * sizeof(CfRun), cudaMemcpyHostToDevice, ctx->stream));
  p.runs = d_runs;
  p.assumed = d_chunk;
  p.arena_top = d_chunk + 3 * nchunks;

  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k_cf_sim<0>, CF_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const uint64_t resident = (uint64_t)ctx->sms * per_sm * CF_THREADS;
  const uint64_t buf_budget = 24ull << 30;  // bytes of stream buffers per launch
  uint64_t       arena_cap = 24 * n + (4ull << 20);  // u32 entries
  int            rc = BK_OK;
  uint64_t       rounds = 0, reruns = 0, big_runs = 0;
  uint32_t       cap_used = 0;
  uint32_t*      d_ovf = dalloc<uint32_t>(ctx, nchunks + 4);  // the list, then its counter (u64, 8-byte aligned)
  if (!d_ovf) { drop(true); return BK_ERR_NOMEM; }
  temps.push_back(d_ovf);
  std::vector<uint32_t> h_list;
  // One launch over `count` chunks (list == null: all) with stacks of `cap` entries; chunks whose stream outgrows them
  // come back in d_ovf and run again, alone, with stacks four times as deep -- as often as it takes (loud failure at
  // the memory budget).  The results of a chunk do not depend on the buffers it ran in.
  auto launch = [&](int mode, const uint32_t* d_list, uint64_t count) -> int {
    uint32_t cap = 4096;
    while (count) {
      uint64_t threads = std::min<uint64_t>(resident, (count + CF_THREADS - 1) / CF_THREADS * CF_THREADS);
      while (threads > CF_THREADS && threads * 2ull * cap * sizeof(uint4) > buf_budget) threads /= 2;
      threads = std::max<uint64_t>(CF_THREADS, threads / CF_THREADS * CF_THREADS);
      if (threads * 2ull * cap * sizeof(uint4) > buf_budget)
        return fail(ctx, BK_ERR_NOMEM, "closest-features: a push-back stream exceeds %u entries", cap / 4);
      p.cap = cap;
      cap_used = std::max(cap_used, cap);
      p.bufs = dalloc<uint4>(ctx, threads * 2ull * cap);
      if (!p.bufs) return BK_ERR_NOMEM;
      p.todo = d_list; p.ntodo = (uint32_t)count;
      p.ovf = d_ovf; p.ovf_count = reinterpret_cast<uint64_t*>(d_ovf + ((nchunks + 1) & ~1ull));
      BK_CUDA(ctx, cudaMemsetAsync(p.ovf_count, 0, 8, ctx->stream));
      prof_begin(ctx, "k_cf_sim");
      if (mode == 0) k_cf_sim<0><<<(unsigned)(threads / CF_THREADS), CF_THREADS, 0, ctx->stream>>>(p);
      else k_cf_sim<1><<<(unsigned)(threads / CF_THREADS), CF_THREADS, 0, ctx->stream>>>(p);
      prof_end(ctx);
      BK_LAUNCHED(ctx);
      uint64_t novf = 0;
      BK_CUDA(ctx, cudaMemcpyAsync(&novf, p.ovf_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
      BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      dfree(ctx, p.bufs);
      p.bufs = nullptr;
      if (novf == 0) break;
      // the overflowed chunks become the work list of the next, deeper launch (d_todo is free to reuse: its content was
      // consumed by the launch that just ended)
      BK_CUDA(ctx, cudaMemcpyAsync(d_todo, d_ovf, novf * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      d_list = d_todo;
      count = novf;
      cap *= 4;
      big_runs++;
    }
    return BK_OK;
  };
  uint64_t* fin[2] = {d_chunk + nchunks, d_chunk + 2 * nchunks};
  int       cur = 0;
  auto run = [&]() -> int {
    BK_TRY(reset_scratch(ctx));
    BK_CUDA(ctx, cudaMemsetAsync(p.arena_top, 0, 8, ctx->stream));
    p.fin_cur = fin[cur]; p.fin_next = fin[cur ^ 1];
    BK_TRY(launch(0, nullptr, nchunks));
    cur ^= 1;
    while (true) {  // at most (chunks of the longest chromosome) rounds: round r makes chunk r of every chromosome exact
      p.fin_cur = fin[cur]; p.fin_next = fin[cur ^ 1];
      const uint64_t cb = std::min<uint64_t>((nchunks + 255) / 256, (uint64_t)ctx->sms * 8);
      k_cf_check<<<(unsigned)cb, 256, 0, ctx->stream>>>(p, d_todo);
      BK_LAUNCHED(ctx);
      BK_TRY(read_scratch(ctx));
      if (ctx->h_scratch[SC_ERR_CODE]) return BK_ERR_NOMEM;
      const uint64_t bad = ctx->h_scratch[SC_COUNT_A];
      if (bad == 0) return BK_OK;
      rounds++;
      reruns += bad;
      BK_CUDA(ctx, cudaMemsetAsync(&ctx->d_scratch[SC_COUNT_A], 0, 8, ctx->stream));
      BK_CUDA(ctx, cudaMemcpyAsync(p.fin_next, p.fin_cur, nchunks * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      BK_TRY(launch(1, d_todo, bad));
      cur ^= 1;
    }
  };
  while (true) {  // the arena of recorded states grows on demand (a full rerun: rare, sized generously)
    p.arena_cap = arena_cap;
    p.arena = dalloc<uint32_t>(ctx, arena_cap);
    if (!p.arena) { rc = BK_ERR_NOMEM; break; }
    rc = run();
    const bool grow = rc == BK_ERR_NOMEM && ctx->h_scratch[SC_ERR_CODE] == BK_ERR_NOMEM && ctx->h_scratch[SC_ERR_ROW] == 1;
    dfree(ctx, p.arena);
    dfree(ctx, p.bufs);
    p.bufs = nullptr;
    if (!grow) break;
    arena_cap *= 4;
    if (arena_cap > (1ull << 36)) {
      rc = fail(ctx, BK_ERR_NOMEM, "closest-features: recorded push-back states exceed the device arena");
      break;
    }
  }
  if (getenv("BEDKIT_TRACE"))
    fprintf(stderr, "[bedkit] closest: %llu chunks, %llu repair rounds, %llu chunk reruns, %llu deep launches, deepest stack %u\n",
            (unsigned long long)nchunks, (unsigned long long)rounds, (unsigned long long)reruns, (unsigned long long)big_runs, cap_used);
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  drop(false);
  if (rc != BK_OK) { dfree(ctx, p.left); dfree(ctx, p.right); return rc; }

  ClosestRow fn{};
  fn.rtext = ref->d_text; fn.rline = ref->line_off; fn.rs = ref->start; fn.re = ref->end; fn.row0 = row0;
  fn.qtext = query->d_text; fn.qline = query->line_off; fn.qs = query->start; fn.qe = query->end;
  fn.left = p.left; fn.right = p.right;
  fn.dist = spec->dist; fn.closest = spec->closest; fn.no_ref = spec->no_ref;
  fn.delim_len = (int)strlen(delim);
  memcpy(fn.delim, delim, fn.delim_len);
  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  rc = run_emit(ctx, fn, n, 0, &d_out, &bytes, &rows);
  dfree(ctx, p.left);
  dfree(ctx, p.right);
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return rc;
  }
  return finish_text(ctx, d_out, bytes, rows, spec->out_on_device, out);
}
