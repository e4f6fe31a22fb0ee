// bedmap.cu -- the bedmap overlap-mapping sweep as data-parallel kernels.
//
// Replaces WindowSweep::sweep + BedBaseVisitor::fixWindow + the visitor zoo
// (interfaces/src/algorithm/sweep/WindowSweepImpl.cpp:174-256, visitors/bed/BedBaseVisitor.hpp:118-225).
// The streaming window is replaced by a per-reference-row candidate range over the sorted map columns:
//     hi = lower_bound(map.start, ref.end + pad)            first map row that starts at/after the reference end
//     lo = lower_bound(map.pmax_end, ref.start - pad + 1)   first map row whose running max end reaches the reference
// Every row in [lo,hi) is tested with the exact overlap predicate (Bed::Overlapping / RangedDist / PercentOverlap* /
// Exact, BedDistances.hpp:41-317) and reduced; the prefix-max index makes the range tight under nesting.
#include <algorithm>
#include <type_traits>
#include "common.cuh"
#include "emit.cuh"
#include "fmt.cuh"
#include "parse.cuh"

namespace bk {

struct OverlapSpec {
  int      kind;
  uint32_t bp;    // BK_OVR_BP: required overlap; BK_OVR_RANGE: padding
  double   frac;  // perc_ exactly as PercentOverlapMapping's constructor leaves it (BedDistances.hpp:120-131)
};

// PercentOverlapMapping::Ref2Map(ref=a, map=b) == 0 given that a and b overlap by ov > 0 bases
__device__ __forceinline__ bool frac_of(uint32_t ov, uint32_t blen, double perc) {
  return (double)ov / (double)blen >= perc;
}

// dist_.Map2Ref(m, r) == 0
__device__ __forceinline__ bool qualifies(const OverlapSpec& o, uint32_t rs, uint32_t re, uint32_t ms, uint32_t me,
                                          uint32_t& ov) {
  const uint32_t mn = rs > ms ? rs : ms, mx = re < me ? re : me;
  ov = mx > mn ? mx - mn : 0;
  switch (o.kind) {
    case BK_OVR_BP: return ov >= o.bp && ov > 0;
    case BK_OVR_RANGE:
      if (ms < re) return (uint64_t)me + o.bp > rs;
      return (uint64_t)re + o.bp > ms;
    case BK_OVR_FRAC_MAP: return ov > 0 && frac_of(ov, me - ms, o.frac);
    case BK_OVR_FRAC_REF: return ov > 0 && frac_of(ov, re - rs, o.frac);
    case BK_OVR_FRAC_EITHER: return ov > 0 && (frac_of(ov, me - ms, o.frac) || frac_of(ov, re - rs, o.frac));
    case BK_OVR_FRAC_BOTH: return ov > 0 && frac_of(ov, me - ms, o.frac) && frac_of(ov, re - rs, o.frac);
    case BK_OVR_EXACT: return rs == ms && re == me;
  }
  return false;
}

enum { NEED_BASES = 1, NEED_SUM = 2, NEED_MAX = 4, NEED_MIN = 8, NEED_IDS = 16 };

struct MapStatsParams {
  const uint32_t* rs;
  const uint32_t* re;
  uint64_t        row0, n;
  const uint64_t* run_ref_begin;  // [nruns+1] ascending reference row numbers
  const uint64_t* run_map_begin;  // [nruns] map rows of the same chromosome (empty range if absent)
  const uint64_t* run_map_end;
  int             nruns;
  const uint32_t* ms;
  const uint32_t* me;
  const uint32_t* pm;
  const uint32_t* bmax;  // max end per 32-row block of the map file (global row index / 32)
  const double*   score;
  const uint32_t* idspan;
  uint64_t        map_rows;  // rows of the whole map file (k_map_group clamps its column loads to it)
  OverlapSpec     ov;
  unsigned        need;
  uint32_t        mdelim_len;
  uint32_t*       count;
  uint64_t*       bases;
  double*         sum;
  double*         vmax;
  double*         vmin;
  uint64_t*       win_lo;
  uint32_t*       win_n;
  uint32_t*       idbytes;
  uint64_t*       scratch;
};

constexpr int MS_THREADS = 256;

// One warp per batch of 32 consecutive reference rows.  For each row of the batch (broadcast by shuffle) the 32
// lanes scan the candidate range [lo,hi) of the map rows of the same chromosome, 32 rows at a time with coalesced
// loads:  lo = first map row whose running-max end exceeds ref.start (galloped from the previous row's lo, a lower
// bound because reference rows are sorted by start),  hi = first map row with start >= ref.end (galloped from lo).
// KIND < 0 selects the generic predicate (runtime overlap kind); FLAGS = NEED_* known at compile time.
template <int KIND, unsigned FLAGS, bool SKIP>
__global__ void __launch_bounds__(MS_THREADS) k_map_stats(MapStatsParams p) {
  const int      lane = threadIdx.x & 31;
  const uint64_t warp0 = ((uint64_t)blockIdx.x * MS_THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * MS_THREADS) >> 5;
  const uint64_t nbatch = (p.n + 31) >> 5;
  OverlapSpec    ov = p.ov;
  if (KIND >= 0) ov.kind = KIND;
  const uint32_t pad = ov.kind == BK_OVR_RANGE ? ov.bp : 0;
  constexpr bool kScore = (FLAGS & (NEED_SUM | NEED_MAX | NEED_MIN)) != 0;
  constexpr bool kMinMax = (FLAGS & (NEED_MAX | NEED_MIN)) != 0;
  for (uint64_t batch = warp0; batch < nbatch; batch += nwarps) {
    const uint64_t i = (batch << 5) + lane;
    const bool     valid = i < p.n;
    const uint64_t row = p.row0 + (valid ? i : p.n - 1);
    const uint32_t my_rs = p.rs[row], my_re = p.re[row];
    int my_run = 0;
    {  // chromosome run of this reference row: last run_ref_begin <= row
      int hi_r = p.nruns;
      while (hi_r - my_run > 1) {
        int mid = (my_run + hi_r) >> 1;
        if (p.run_ref_begin[mid] <= row) my_run = mid; else hi_r = mid;
      }
    }
    // Window starts of the whole batch at once: lo = first map row whose running-max end reaches the (padded) reference
    // start.  Reference rows are sorted by start, so lo is non-decreasing within a chromosome: two warp-cooperative
    // searches bracket the batch (first and last row), then every lane bisects its own key inside the bracket.
    const int      nj = (int)((p.n - (batch << 5)) < 32 ? (p.n - (batch << 5)) : 32);
    const uint32_t my_key = my_rs >= pad ? my_rs - pad + 1 : 0;
    uint32_t       my_lo;
    {
      const int run_a = __shfl_sync(0xffffffffu, my_run, 0);
      if (__all_sync(0xffffffffu, my_run == run_a)) {
        const uint64_t  mb0 = p.run_map_begin[run_a];
        const uint32_t  nr0 = (uint32_t)(p.run_map_end[run_a] - mb0);
        const uint32_t* pm0 = p.pm + mb0;
        const uint32_t  key_a = __shfl_sync(0xffffffffu, my_key, 0), key_b = __shfl_sync(0xffffffffu, my_key, nj - 1);
        uint32_t        a = warp_search32(pm0, nr0, key_a, lane);
        uint32_t        b = warp_gallop(pm0, a, nr0, key_b, lane, true);
        while (a < b) {  // same trip count on every lane
          const uint32_t mid = a + ((b - a) >> 1);
          if (__ldg(&pm0[mid]) < my_key) a = mid + 1; else b = mid;
        }
        my_lo = a;
      } else {  // the batch straddles a chromosome boundary: every lane searches its own chromosome
        const uint64_t mbl = p.run_map_begin[my_run];
        my_lo = (uint32_t)lower_bound_u32(p.pm + mbl, 0, p.run_map_end[my_run] - mbl, my_key);
      }
    }
    uint32_t out_cnt = 0, out_idb = 0, out_n = 0, out_lo = 0;
    uint64_t out_bases = 0, out_mb = 0;
    double   out_sum = 0.0, out_max = 0.0, out_min = 0.0;
    int      hint_run = -1;
    uint32_t nr = 0;
    uint64_t mb = 0;
    const uint32_t *ms = nullptr, *me = nullptr, *ids = nullptr;
    const double*   sc = nullptr;
#pragma unroll 1
    for (int j = 0; j < nj; j++) {
      const uint32_t rs = __shfl_sync(0xffffffffu, my_rs, j), re = __shfl_sync(0xffffffffu, my_re, j);
      const int      run = __shfl_sync(0xffffffffu, my_run, j);
      const uint32_t lo = __shfl_sync(0xffffffffu, my_lo, j);
      if (run != hint_run) {
        mb = p.run_map_begin[run];
        nr = (uint32_t)(p.run_map_end[run] - mb);
        ms = p.ms + mb;
        me = p.me + mb;
        if (kScore) sc = p.score + mb;
        if ((FLAGS & NEED_IDS) && p.idspan) ids = p.idspan + mb;
        hint_run = run;
      }
      const uint64_t re_pad64 = (uint64_t)re + pad;
      const uint32_t re_pad = re_pad64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)re_pad64;
      uint32_t cnt = 0, idb = 0, nwin = 0;
      uint64_t bases = 0;
      double   sum = 0.0, vmax = 0.0, vmin = 0.0;
      bool     have = false;
      if (SKIP) {
        // Dense map files (windows of hundreds of rows, most of them short rows that end in front of the reference
        // row): steps are aligned to the 32-row blocks of the block-max index, and a step whose two blocks hold no end
        // that reaches the reference row (block max end < key) is skipped without touching its rows.  Rows in front
        // of lo (alignment) never qualify; a skipped step cannot contain the row that ends the window (start >=
        // ref.end implies that its end reaches).
        const uint32_t key = rs >= pad ? rs - pad + 1 : 0;
        const uint64_t gend = mb + nr;
#pragma unroll 1
        for (uint64_t g0 = (mb + lo) & ~31ull; g0 < gend; g0 += 64) {
          const uint32_t m0 = __ldg(&p.bmax[g0 >> 5]), m1 = g0 + 32 < gend ? __ldg(&p.bmax[(g0 >> 5) + 1]) : 0u;
          if (m0 < key && m1 < key) continue;
          const uint64_t ga = g0 + lane, gb = ga + 32;
          const bool     va = ga >= mb && ga < gend, vb = gb < gend;
          const uint32_t sa = va ? __ldg(&p.ms[ga]) : 0xFFFFFFFFu, sb = vb ? __ldg(&p.ms[gb]) : 0xFFFFFFFFu;
          const uint32_t ea = va ? __ldg(&p.me[ga]) : 0u, eb = vb ? __ldg(&p.me[gb]) : 0u;
          const bool     ina = va && sa < re_pad, inb = vb && sb < re_pad;
          uint32_t       ova = 0, ovb = 0;
          const bool     qa = ina && qualifies(ov, rs, re, sa, ea, ova);
          const bool     qb = inb && qualifies(ov, rs, re, sb, eb, ovb);
          if (qa) {
            cnt++;
            if (FLAGS & NEED_BASES) bases += ova;
            if (kScore) {
              const double v = __ldg(&p.score[ga]);
              sum += v;
              if (kMinMax) {
                vmax = have ? (v > vmax ? v : vmax) : v;
                vmin = have ? (v < vmin ? v : vmin) : v;
                have = true;
              }
            }
          }
          if (qb) {
            cnt++;
            if (FLAGS & NEED_BASES) bases += ovb;
            if (kScore) {
              const double v = __ldg(&p.score[gb]);
              sum += v;
              if (kMinMax) {
                vmax = have ? (v > vmax ? v : vmax) : v;
                vmin = have ? (v < vmin ? v : vmin) : v;
                have = true;
              }
            }
          }
          if (__ballot_sync(0xffffffffu, inb) != 0xffffffffu) break;
        }
      } else {
      // Scan 64 map rows per step (two coalesced 32-row chunks whose loads are issued together).  Map starts are
      // sorted, so the window ends in the first step in which some lane of the second chunk sees start >= ref.end.
#pragma unroll 1
      for (uint32_t k0 = lo; k0 < nr; k0 += 64) {
        const uint32_t ka = k0 + lane, kb = ka + 32;
        uint32_t       sa, sb, ea, eb;
        if (k0 + 64 <= nr) {  // interior step (warp-uniform): no bounds checks on the 64 rows
          sa = __ldg(&ms[ka]); sb = __ldg(&ms[kb]);
          ea = __ldg(&me[ka]); eb = __ldg(&me[kb]);
        } else {              // last step of the chromosome: rows past its end read as "starts at infinity"
          const bool va = ka < nr, vb = kb < nr;
          sa = va ? __ldg(&ms[ka]) : 0xFFFFFFFFu; sb = vb ? __ldg(&ms[kb]) : 0xFFFFFFFFu;
          ea = va ? __ldg(&me[ka]) : 0u; eb = vb ? __ldg(&me[kb]) : 0u;
        }
        const bool ina = sa < re_pad, inb = sb < re_pad;  // re_pad <= 0xFFFFFFFF: the sentinel is never in range
        uint32_t       ova = 0, ovb = 0;
        const bool     qa = ina && qualifies(ov, rs, re, sa, ea, ova);
        const bool     qb = inb && qualifies(ov, rs, re, sb, eb, ovb);
        if (qa) {
          cnt++;
          if (FLAGS & NEED_BASES) bases += ova;
          if (kScore) {
            const double v = __ldg(&sc[ka]);
            sum += v;
            if (kMinMax) {
              vmax = have ? (v > vmax ? v : vmax) : v;
              vmin = have ? (v < vmin ? v : vmin) : v;
              have = true;
            }
          }
          if ((FLAGS & NEED_IDS) && ids) idb += __ldg(&ids[ka]) & 0xFFFFu;
        }
        if (qb) {
          cnt++;
          if (FLAGS & NEED_BASES) bases += ovb;
          if (kScore) {
            const double v = __ldg(&sc[kb]);
            sum += v;
            if (kMinMax) {
              vmax = have ? (v > vmax ? v : vmax) : v;
              vmin = have ? (v < vmin ? v : vmin) : v;
              have = true;
            }
          }
          if ((FLAGS & NEED_IDS) && ids) idb += __ldg(&ids[kb]) & 0xFFFFu;
        }
        const unsigned mb_ = __ballot_sync(0xffffffffu, inb);
        if (FLAGS & NEED_IDS) nwin += __popc(__ballot_sync(0xffffffffu, ina)) + __popc(mb_);
        if (mb_ != 0xffffffffu) break;
      }
      }
      const uint32_t hi = lo + nwin;
      // warp reductions (fixed order: deterministic)
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (FLAGS & NEED_BASES) bases = warp_sum_u64(bases);
      if (kScore) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, d);
          if (kMinMax) {
            const double omx = __shfl_xor_sync(0xffffffffu, vmax, d);
            const double omn = __shfl_xor_sync(0xffffffffu, vmin, d);
            const bool   oh = __shfl_xor_sync(0xffffffffu, (int)have, d);
            if (oh) {
              vmax = have ? (omx > vmax ? omx : vmax) : omx;
              vmin = have ? (omn < vmin ? omn : vmin) : omn;
              have = true;
            }
          }
        }
      }
      if (FLAGS & NEED_IDS) {
        idb = __reduce_add_sync(0xffffffffu, idb);
        if (cnt) idb += (cnt - 1) * p.mdelim_len;
      }
      if (lane == j) {
        out_cnt = cnt; out_bases = bases; out_sum = sum; out_max = vmax; out_min = vmin;
        out_lo = lo; out_mb = mb; out_n = hi - lo; out_idb = idb;
      }
    }
    if (valid) {
      p.count[i] = out_cnt;
      if (FLAGS & NEED_BASES) p.bases[i] = out_bases;
      if (FLAGS & NEED_SUM) p.sum[i] = out_sum;
      if (FLAGS & NEED_MAX) p.vmax[i] = out_max;
      if (FLAGS & NEED_MIN) p.vmin[i] = out_min;
      if (FLAGS & NEED_IDS) {
        p.win_lo[i] = out_mb + out_lo;
        p.win_n[i] = out_n;
        p.idbytes[i] = out_idb;
      }
    }
  }
}

// ---- lane-per-row form ---------------------------------------------------------------------------------------------
// The same windows reduced without any cross-lane reduction: a lane OWNS one reference row, and the eight lanes of a group
// (eight consecutive reference rows, whose windows overlap almost completely: rows are sorted) walk the union of their
// windows together.  The group stages 64 map rows at a time in its own 1 KB of shared memory (coalesced column loads, start
// and end interleaved), then every lane tests the same staged row against its own reference row: the shared-memory reads
// are broadcasts (four distinct 16-byte addresses per warp instruction, the groups' buffers are skewed by 16 banks), a hit
// costs a predicated add, and each lane sums its hits in file order.  Per map row and lane ~12 instructions and no
// shuffle; the warp-per-row kernel above spends ~270 warp instructions per reference row at configuration 2, this form
// ~110 (a lane also tests the rows that only its neighbours' windows hold: 196 instead of 126 rows).
// Rows in front of a lane's own lo never qualify (their end does not reach the reference row) but may belong to the
// previous chromosome, so a lane looks at rows [glo, gend) only: a 64-bit mask per chunk.
constexpr int MG_THREADS = 256;
constexpr int MG_G = 8;    // lanes = reference rows per group
constexpr int MG_CH = 64;  // map rows per staged chunk
struct alignas(16) MgBuf {
  uint2    se[MG_CH];  // (start, end)
  double   sc[MG_CH];
  uint32_t skew[16];   // group stride = 1088 bytes = 16 banks: the two groups of a half-warp touch disjoint banks (STS.64 and LDS.128)
};

__device__ __forceinline__ uint32_t group_min_u32(uint32_t v) {
#pragma unroll
  for (int d = 1; d < MG_G; d <<= 1) {
    const uint32_t o = __shfl_xor_sync(0xffffffffu, v, d);
    v = o < v ? o : v;
  }
  return v;
}

#ifndef BK_MG_COOP
#define BK_MG_COOP 6  // at most this many unfinished lanes: their rows go to the whole warp (break-even ~14, see below)
#endif

// The phase in which every lane walks its group's chunks pays for the longest union among the four groups of the warp,
// and window lengths are heavy-tailed (log-normal reference and map lengths: one 20 kb reference row keeps its group
// busy for ten chunks while 31 lanes idle).  A chunk step costs ~1000 warp instructions whatever the number of lanes
// that still need it, the warp-cooperative step of k_map_stats ~75 per reference row and 64 map rows: once no more than
// BK_MG_COOP lanes are unfinished, the rest of each one's window is scanned by all 32 lanes (coalesced column loads, one
// warp reduction per row) and added to the lane's sums.  Map row indices are 32-bit here (the caller checks the file).
template <int KIND, unsigned FLAGS>
__global__ void __launch_bounds__(MG_THREADS) k_map_group(MapStatsParams p) {
  __shared__ MgBuf bufs[MG_THREADS / MG_G];
  const int      lane = threadIdx.x & 31;
  const int      l8 = lane & (MG_G - 1);
  const unsigned gmask = ((1u << MG_G) - 1u) << (lane & ~(MG_G - 1));
  MgBuf&         B = bufs[threadIdx.x / MG_G];
  const uint64_t warp0 = ((uint64_t)blockIdx.x * MG_THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * MG_THREADS) >> 5;
  const uint64_t nbatch = (p.n + 31) >> 5;
  OverlapSpec    ov = p.ov;
  if (KIND >= 0) ov.kind = KIND;
  const uint32_t pad = ov.kind == BK_OVR_RANGE ? ov.bp : 0;
  const uint32_t bp1 = ov.bp > 1 ? ov.bp : 1;
  constexpr bool kScore = (FLAGS & (NEED_SUM | NEED_MAX | NEED_MIN)) != 0;
  constexpr bool kMinMax = (FLAGS & (NEED_MAX | NEED_MIN)) != 0;
  const uint32_t last_row = (uint32_t)p.map_rows - 1u;
  for (uint64_t batch = warp0; batch < nbatch; batch += nwarps) {
    const uint64_t i = (batch << 5) + lane;
    const bool     valid = i < p.n;
    const uint64_t row = p.row0 + (valid ? i : p.n - 1);
    const uint32_t rs = p.rs[row], re = p.re[row];
    int my_run = 0;
    {
      int hi_r = p.nruns;
      while (hi_r - my_run > 1) {
        int mid = (my_run + hi_r) >> 1;
        if (p.run_ref_begin[mid] <= row) my_run = mid; else hi_r = mid;
      }
    }
    // window starts of the batch: as in k_map_stats
    const int      nj = (int)((p.n - (batch << 5)) < 32 ? (p.n - (batch << 5)) : 32);
    const uint32_t my_key = rs >= pad ? rs - pad + 1 : 0;
    const uint32_t mbl = (uint32_t)p.run_map_begin[my_run];
    const uint32_t gend = (uint32_t)p.run_map_end[my_run];
    uint32_t       my_lo;
    {
      const int run_a = __shfl_sync(0xffffffffu, my_run, 0);
      if (__all_sync(0xffffffffu, my_run == run_a)) {
        const uint32_t  nr0 = gend - mbl;
        const uint32_t* pm0 = p.pm + mbl;
        const uint32_t  key_a = __shfl_sync(0xffffffffu, my_key, 0), key_b = __shfl_sync(0xffffffffu, my_key, nj - 1);
        uint32_t        a = warp_search32(pm0, nr0, key_a, lane);
        uint32_t        b = warp_gallop(pm0, a, nr0, key_b, lane, true);
        while (a < b) {
          const uint32_t mid = a + ((b - a) >> 1);
          if (__ldg(&pm0[mid]) < my_key) a = mid + 1; else b = mid;
        }
        my_lo = a;
      } else {
        my_lo = (uint32_t)lower_bound_u32(p.pm + mbl, 0, gend - mbl, my_key);
      }
    }
    const uint32_t glo = mbl + my_lo;
    const uint64_t re_pad64 = (uint64_t)re + pad;
    const uint32_t re_pad = re_pad64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)re_pad64;

    uint32_t idb = 0, nwin = 0;
    uint64_t bases = 0;
    double   cntd = 0.0, sum = 0.0, vmax = 0.0, vmin = 0.0;  // hits are counted on the FP64 pipe (exact; the ALU pipe is the busy one)
    bool     have = false;
    bool     fin = !valid || glo >= gend;
    // the reference row the tests use: a finished lane carries a null row that no map row overlaps
    uint32_t rs_e = fin ? 0xFFFFFFFFu : rs, re_e = fin ? 0u : re, rp_e = fin ? 0u : re_pad;
    // one staged map row against my reference row; `mine`: the row is one of [glo, gend) (constant true on the fast path).
    // NARROW: overlaps are summed in 32 bits and folded into the 64-bit sum once per chunk (one add per row instead of
    // two; the ALU pipe, two cycles per warp instruction, is the busy one); the caller guarantees 64 * (re - rs) < 2^32.
    uint32_t bases32 = 0;
    auto test_row = [&](auto narrow, uint32_t s, uint32_t e, double v, bool mine, uint32_t k) {
      uint32_t ovl;
      bool     q;
      if (KIND == BK_OVR_BP) {
        const uint32_t mn = rs_e > s ? rs_e : s, mx = re_e < e ? re_e : e;
        ovl = mx - (mn < mx ? mn : mx);
        q = mine && ovl >= bp1;
      } else {
        q = mine && s < rp_e && qualifies(ov, rs_e, re_e, s, e, ovl);
      }
      if (q) {  // the adds as volatile asm: a predicated DADD each, not an unconditional add followed by two selects
        asm volatile("add.f64 %0, %0, 0d3FF0000000000000;" : "+d"(cntd));
        if (FLAGS & NEED_BASES) {
          if (decltype(narrow)::value) bases32 += ovl;
          else bases += ovl;
        }
        if (kScore) {
          asm volatile("add.f64 %0, %0, %1;" : "+d"(sum) : "d"(v));
          if (kMinMax) {
            vmax = have ? (v > vmax ? v : vmax) : v;
            vmin = have ? (v < vmin ? v : vmin) : v;
            have = true;
          }
        }
        if ((FLAGS & NEED_IDS) && p.idspan) idb += __ldg(&p.idspan[k]) & 0xFFFFu;
      }
      if (FLAGS & NEED_IDS) nwin += (mine && s < rp_e) ? 1u : 0u;
    };
    const bool huge = re - rs >= (1u << 26);  // 64 overlaps of this reference row may not fit 32 bits: masked path
    // chunk starts are multiples of 8 rows (sector-aligned column loads); with the block-max index (dense map files) of 32
    const uint32_t calign = p.bmax ? ~31u : ~7u;
    uint32_t c0 = group_min_u32(fin ? 0xFFFFFFFFu : glo) & calign;  // next chunk of my group
    unsigned live;
    while (__popc(live = __ballot_sync(0xffffffffu, !fin)) > BK_MG_COOP) {
      if (p.bmax) {
        // Dense map files: most rows between lo and the first hit are short rows that end in front of the reference rows.
        // A chunk (two 32-row blocks of the block-max-end index) in which no end reaches the smallest key of the group's
        // unfinished lanes holds no hit for them and not the row that ends a window either (start >= ref.end implies
        // that its end reaches): when that is so for all four groups, the chunk is neither staged nor scanned.
        const uint32_t gk = group_min_u32(fin ? 0xFFFFFFFFu : my_key);
        bool           dead = true;
        if (live & gmask) {
          const uint32_t b0 = c0 >> 5;
          const uint32_t m0 = __ldg(&p.bmax[b0]), m1 = c0 + 32u <= last_row ? __ldg(&p.bmax[b0 + 1]) : 0u;
          dead = m0 < gk && m1 < gk;
        }
        if (__all_sync(0xffffffffu, dead)) {
          const uint32_t c1 = c0 + MG_CH;
          if (!fin && c1 >= gend) {
            fin = true;
            rs_e = 0xFFFFFFFFu; re_e = 0u; rp_e = 0u;
          }
          c0 = group_min_u32(fin ? 0xFFFFFFFFu : (glo > c1 ? glo & calign : c1));
          continue;
        }
      }
      if (live & gmask) {  // stage map rows [c0, c0+64) of the columns; rows past the file repeat its last row (masked below)
#pragma unroll
        for (int r = 0; r < MG_CH / MG_G; r++) {
          const uint32_t k = c0 + (uint32_t)(r * MG_G + l8);
          const uint32_t idx = k < last_row ? k : last_row;
          B.se[r * MG_G + l8] = make_uint2(__ldg(&p.ms[idx]), __ldg(&p.me[idx]));
          if (kScore) B.sc[r * MG_G + l8] = __ldg(&p.score[idx]);
        }
      }
      __syncwarp();
      // Fast path: the chunk lies inside the chromosome of every unfinished lane.  Rows in front of a lane's lo cannot
      // qualify (their end does not reach the reference row), so no per-row mask is needed: 64 rows, ~9 instructions each.
      // (Reference rows longer than 2^26 bases also send the warp to the masked path: see test_row.)
      const bool edge = !fin && (c0 < mbl || c0 + MG_CH > gend || huge);
      if (!(FLAGS & NEED_IDS) && !__any_sync(0xffffffffu, edge)) {
#pragma unroll 1
        for (int j = 0; j < MG_CH / 8; j++) {
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            const uint4 se2 = *reinterpret_cast<const uint4*>(&B.se[8 * j + u]);
            double2     v2 = make_double2(0.0, 0.0);
            if (kScore) v2 = *reinterpret_cast<const double2*>(&B.sc[8 * j + u]);
            test_row(std::true_type{}, se2.x, se2.y, v2.x, true, 0u);
            test_row(std::true_type{}, se2.z, se2.w, v2.y, true, 0u);
          }
        }
        if (FLAGS & NEED_BASES) {
          bases += bases32;
          bases32 = 0;
        }
      } else {  // chromosome edges and the list operations (which need the exact window): per-row mask of my rows [glo, gend)
        uint64_t wmask = 0;
        if (!fin) {
          const uint32_t a = glo > c0 ? (glo - c0 < 64u ? glo - c0 : 64u) : 0u;
          const uint32_t b = gend - c0 < 64u ? gend - c0 : 64u;  // gend > c0: the lane is not finished
          const uint64_t upto_b = b >= 64u ? ~0ull : ((1ull << b) - 1ull);
          const uint64_t upto_a = a >= 64u ? ~0ull : ((1ull << a) - 1ull);
          wmask = upto_b & ~upto_a;
        }
#pragma unroll 1
        for (int j = 0; j < MG_CH / 8; j++) {
          const uint32_t m8 = (uint32_t)(wmask >> (8 * j)) & 0xFFu;
          if (!__any_sync(0xffffffffu, m8 != 0u)) continue;
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            const uint4 se2 = *reinterpret_cast<const uint4*>(&B.se[8 * j + u]);
            double2     v2 = make_double2(0.0, 0.0);
            if (kScore) v2 = *reinterpret_cast<const double2*>(&B.sc[8 * j + u]);
            test_row(std::false_type{}, se2.x, se2.y, v2.x, (m8 >> u) & 1u, c0 + (uint32_t)(8 * j + u));
            test_row(std::false_type{}, se2.z, se2.w, v2.y, (m8 >> (u + 1)) & 1u, c0 + (uint32_t)(8 * j + u + 1));
          }
        }
      }
      // a lane is finished when its chromosome is used up or the last staged row (one of its rows) starts at/after its end
      const uint32_t s_last = B.se[MG_CH - 1].x;
      const uint32_t c1 = c0 + MG_CH;
      if (!fin && (c1 >= gend || (c1 > glo && s_last >= re_pad))) {
        fin = true;
        rs_e = 0xFFFFFFFFu; re_e = 0u; rp_e = 0u;
      }
      __syncwarp();  // everyone has read the chunk before the next one is staged
      c0 = group_min_u32(fin ? 0xFFFFFFFFu : (glo > c1 ? glo & calign : c1));  // skips gaps (chromosome change, sparse windows)
    }
    uint32_t cnt = (uint32_t)cntd;
    // the stragglers, one after the other: rows [max(c0, glo), gend) of lane j's window by the whole warp, 64 rows per step
    while (live) {
      const int j = __ffs(live) - 1;
      live &= live - 1;
      const uint32_t rs_j = __shfl_sync(0xffffffffu, rs, j), re_j = __shfl_sync(0xffffffffu, re, j);
      const uint32_t rp_j = __shfl_sync(0xffffffffu, re_pad, j), end_j = __shfl_sync(0xffffffffu, gend, j);
      const uint32_t from_j = __shfl_sync(0xffffffffu, c0 > glo ? c0 : glo, j);
      uint32_t       cnt2 = 0, idb2 = 0, nwin2 = 0;
      uint64_t       bases2 = 0;
      double         sum2 = 0.0, vmax2 = 0.0, vmin2 = 0.0;
      bool           have2 = false;
#pragma unroll 1
      for (uint32_t k0 = from_j; k0 < end_j; k0 += 64) {
        const uint32_t ka = k0 + lane, kb = ka + 32;
        const bool     va = ka < end_j, vb = kb < end_j;
        const uint32_t sa = va ? __ldg(&p.ms[ka]) : 0xFFFFFFFFu, sb = vb ? __ldg(&p.ms[kb]) : 0xFFFFFFFFu;
        const uint32_t ea = va ? __ldg(&p.me[ka]) : 0u, eb = vb ? __ldg(&p.me[kb]) : 0u;
        const bool     ina = sa < rp_j, inb = sb < rp_j;  // rp_j <= 0xFFFFFFFF: the sentinel is never in range
        uint32_t       ova = 0, ovb = 0;
        const bool     qa = ina && qualifies(ov, rs_j, re_j, sa, ea, ova);
        const bool     qb = inb && qualifies(ov, rs_j, re_j, sb, eb, ovb);
        if (qa) {
          cnt2++;
          if (FLAGS & NEED_BASES) bases2 += ova;
          if (kScore) {
            const double v = __ldg(&p.score[ka]);
            sum2 += v;
            if (kMinMax) {
              vmax2 = have2 ? (v > vmax2 ? v : vmax2) : v;
              vmin2 = have2 ? (v < vmin2 ? v : vmin2) : v;
              have2 = true;
            }
          }
          if ((FLAGS & NEED_IDS) && p.idspan) idb2 += __ldg(&p.idspan[ka]) & 0xFFFFu;
        }
        if (qb) {
          cnt2++;
          if (FLAGS & NEED_BASES) bases2 += ovb;
          if (kScore) {
            const double v = __ldg(&p.score[kb]);
            sum2 += v;
            if (kMinMax) {
              vmax2 = have2 ? (v > vmax2 ? v : vmax2) : v;
              vmin2 = have2 ? (v < vmin2 ? v : vmin2) : v;
              have2 = true;
            }
          }
          if ((FLAGS & NEED_IDS) && p.idspan) idb2 += __ldg(&p.idspan[kb]) & 0xFFFFu;
        }
        const unsigned mb_ = __ballot_sync(0xffffffffu, inb);
        if (FLAGS & NEED_IDS) nwin2 += __popc(__ballot_sync(0xffffffffu, ina)) + __popc(mb_);
        if (mb_ != 0xffffffffu) break;
      }
      cnt2 = __reduce_add_sync(0xffffffffu, cnt2);
      if (FLAGS & NEED_BASES) bases2 = warp_sum_u64(bases2);
      if (kScore) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          sum2 += __shfl_xor_sync(0xffffffffu, sum2, d);
          if (kMinMax) {
            const double omx = __shfl_xor_sync(0xffffffffu, vmax2, d);
            const double omn = __shfl_xor_sync(0xffffffffu, vmin2, d);
            const bool   oh = __shfl_xor_sync(0xffffffffu, (int)have2, d);
            if (oh) {
              vmax2 = have2 ? (omx > vmax2 ? omx : vmax2) : omx;
              vmin2 = have2 ? (omn < vmin2 ? omn : vmin2) : omn;
              have2 = true;
            }
          }
        }
      }
      if (FLAGS & NEED_IDS) idb2 = __reduce_add_sync(0xffffffffu, idb2);
      if (lane == j) {
        cnt += cnt2;
        bases += bases2;
        sum += sum2;
        if (kMinMax && have2) {
          vmax = have ? (vmax2 > vmax ? vmax2 : vmax) : vmax2;
          vmin = have ? (vmin2 < vmin ? vmin2 : vmin) : vmin2;
          have = true;
        }
        idb += idb2;
        nwin += nwin2;
      }
    }
    if (valid) {
      p.count[i] = cnt;
      if (FLAGS & NEED_BASES) p.bases[i] = bases;
      if ((FLAGS & NEED_SUM) && p.sum) p.sum[i] = sum;
      if ((FLAGS & NEED_MAX) && p.vmax) p.vmax[i] = vmax;
      if ((FLAGS & NEED_MIN) && p.vmin) p.vmin[i] = vmin;
      if (FLAGS & NEED_IDS) {
        if (cnt) idb += (cnt - 1) * p.mdelim_len;
        p.win_lo[i] = glo;
        p.win_n[i] = nwin;
        p.idbytes[i] = idb;
      }
    }
  }
}

// MAX and MIN share an instantiation (both are kept; a null output pointer drops the store), SUM rides along with them
template <int KIND>
static void launch_map_group(unsigned need, unsigned blocks, cudaStream_t st, const MapStatsParams& sp) {
  unsigned f = need & 31u;
  if (f & (NEED_MAX | NEED_MIN)) f |= NEED_MAX | NEED_MIN;
  switch (f) {
#define BK_F(F) case F: k_map_group<KIND, F><<<blocks, MG_THREADS, 0, st>>>(sp); break;
    BK_F(0) BK_F(1) BK_F(2) BK_F(3) BK_F(12) BK_F(13) BK_F(14) BK_F(15)
    BK_F(16) BK_F(17) BK_F(18) BK_F(19) BK_F(28) BK_F(29) BK_F(30) BK_F(31)
#undef BK_F
  }
}

template <int KIND, bool SKIP>
static void launch_map_stats(unsigned need, unsigned blocks, cudaStream_t st, const MapStatsParams& sp) {
  switch (need & (SKIP ? 15u : 31u)) {
#define BK_F(F) case F: k_map_stats<KIND, F, SKIP><<<blocks, MS_THREADS, 0, st>>>(sp); break;
    BK_F(0) BK_F(1) BK_F(2) BK_F(3) BK_F(4) BK_F(5) BK_F(6) BK_F(7) BK_F(8) BK_F(9) BK_F(10) BK_F(11) BK_F(12) BK_F(13)
    BK_F(14) BK_F(15)
#undef BK_F
#define BK_F(F) case F: if (!SKIP) k_map_stats<KIND, F, false><<<blocks, MS_THREADS, 0, st>>>(sp); break;
    BK_F(16) BK_F(17) BK_F(18) BK_F(19) BK_F(20) BK_F(21) BK_F(22) BK_F(23) BK_F(24) BK_F(25) BK_F(26)
    BK_F(27) BK_F(28) BK_F(29) BK_F(30) BK_F(31)
#undef BK_F
  }
}

// RARE bit 0: --sci output; bit 1: the reference file is a B4Rest/B5Rest (single-file mode); bit 2: per-hit list /
// unique-bases operations other than --echo-map-id -- compile-time so that the
// common instantiation carries none of that code (the emitter is register-bound).
template <int RARE>
struct BedmapRow {
  // reference rows
  const char*     rtext;
  const uint64_t* rline;
  const uint32_t* rs;
  const uint32_t* re;
  uint64_t        row0;
  int             ref_fields;  // record type of the reference file: 3 = B3Rest; 4/5 only in single-file mode
  const double*   rscore;
  // map side (echo-map-id)
  const char*     mtext;
  const uint64_t* mline;
  const uint32_t* midspan;
  const uint32_t* ms;
  const uint32_t* me;
  const double*   mscore;
  int             map_fields;  // record type of the map file (3|4|5): how --echo-map prints a row
  // per-row results (indexed by i = row - row0)
  const uint32_t* count;
  const uint64_t* bases;
  const double*   sum;
  const double*   vmax;
  const double*   vmin;
  const uint64_t* win_lo;
  const uint32_t* win_n;
  const uint32_t* idbytes;
  const uint64_t* rank;        // --echo-ref-row-id with --skip-unmapped: printed rows before row i (else null: i)
  uint64_t        rowid_base;  // ids issued by earlier calls of the same command
  int             rowid_ops;   // how many --echo-ref-row-id columns a row prints (each one bumps the counter)
  OverlapSpec     ov;
  int             n_ops;
  unsigned char   ops[BK_MAX_OPS];
  double          kth_arg[BK_MAX_OPS];  // BK_OP_KTH fraction, BK_OP_MAD multiplier, BK_OP_TMEAN <low>
  double          arg2[BK_MAX_OPS];     // BK_OP_TMEAN <hi>
  uint64_t        stop_row;             // --max/min-element: the first unmapped row (printed up to the element column), or ~0
  uint64_t*       idfill_off;           // --echo-map-id filled by k_fill_ids: where the column of row i lands in the result (else null)
  int             prec;
  int             sci;
  int             skip_unmapped;
  char            delim[24];
  int             delim_len;
  char            mdelim[24];
  int             mdelim_len;
  uint64_t*       scratch;

  template <class Sink>
  __device__ __forceinline__ void put_score(Sink& s, double v, uint32_t cnt, uint64_t i) const {
    if (cnt == 0) {
      s.puts_("NAN", 3);  // Signal::NaN::nan_ (interfaces/src/data/measurement/NaN.cpp:27)
      return;
    }
    if ((RARE & 1) && sci) {  // "%.<prec>e" (Formats.hpp:42-49)
      put_sci(s, v, i);
      return;
    }
    if (Sink::counting) {
      const int L = fixed_len_fast(v, prec);
      if (L >= 0) {
        s.copy(nullptr, (uint64_t)L);
        return;
      }
    }
    Fixed f;
    if (!to_fixed(v, prec, f)) {
      dev_set_error(scratch, BK_ERR_UNSUPPORTED, i);
      s.put('?');
      return;
    }
    put_fixed(s, f, prec);
  }

  template <class Sink>
  __device__ __noinline__ void put_sci(Sink& s, double v, uint64_t i) const {
    {
      Sci e;
      if (!to_sci(v, prec, e)) {
        dev_set_error(scratch, BK_ERR_UNSUPPORTED, i);
        s.put('?');
        return;
      }
      if (e.neg) s.put('-');
      if (e.special) {
        if (e.special == 1) s.puts_("nan", 3); else s.puts_("inf", 3);
        return;
      }
      const uint64_t pw = pow10_u64(prec);
      s.put((char)('0' + (int)(e.digits / pw)));
      if (prec > 0) {
        s.put('.');
        s.put_padded(e.digits % pw, prec);
      }
      s.put('e');
      s.put(e.exp10 < 0 ? '-' : '+');
      const uint32_t ae = (uint32_t)(e.exp10 < 0 ? -e.exp10 : e.exp10);
      if (ae < 10) s.put('0');
      s.put_u32(ae);
    }
  }

  // the operations that walk the qualifying map rows of reference row `row` in file order (= the order of the
  // reference's std::set<MapType*, GenomicAddressCompare>, EchoMapBedVisitor.hpp:62)
  template <class Sink>
  __device__ __noinline__ void window_op(Sink& s, int op, uint64_t i, uint64_t row) const {
    const uint64_t lo = win_lo[i], hi = lo + win_n[i];
    const uint32_t a = rs[row], b = re[row];
    bool           first = true;
    uint32_t       mn = 0xFFFFFFFFu, mx = 0;   // --echo-map-range
    uint32_t       cs = 0, ce = 0;             // --bases-uniq: the open union run
    uint64_t       uniq = 0;
    double         vsum = 0.0, vsq = 0.0;      // --variance / --stdev / --cv
    uint32_t       nhit = 0;
    for (uint64_t k = lo; k < hi; k++) {
      uint32_t       ovl;
      const uint32_t s0 = ms[k], e0 = me[k];
      if (!qualifies(ov, a, b, s0, e0, ovl)) continue;
      if (op == BK_OP_ECHO_MAP_RANGE) {
        mn = s0 < mn ? s0 : mn;
        mx = e0 > mx ? e0 : mx;
        first = false;
        continue;
      }
      if (op >= BK_OP_VARIANCE) {  // variance family: sum and sum of squares in file order
        const double v = mscore[k];
        vsum += v;
        vsq += v * v;
        nhit++;
        continue;
      }
      if (op == BK_OP_BASES_UNIQ || op == BK_OP_BASES_UNIQ_F) {
        if (first) {
          cs = s0; ce = e0;
        } else if ((ce < e0 ? ce : e0) > (cs > s0 ? cs : s0)) {  // Bed.hpp:172-191 overlap() > 0 -> eunion (:202-210)
          cs = cs < s0 ? cs : s0;
          ce = ce > e0 ? ce : e0;
        } else {
          const uint32_t x = ce < b ? ce : b, y = cs > a ? cs : a;
          uniq += x > y ? x - y : 0;
          cs = s0; ce = e0;
        }
        first = false;
        continue;
      }
      if (!first) s.puts_(mdelim, mdelim_len);
      first = false;
      switch (op) {
        case BK_OP_ECHO_MAP:
          if (map_fields <= 3) echo_b3rest(s, mtext, mline[k], s0, e0);
          else echo_b45rest(s, mtext, mline[k], s0, e0, map_fields, map_fields >= 5 ? mscore[k] : 0.0, scratch, i);
          break;
        case BK_OP_ECHO_MAP_SCORE: put_score(s, mscore[k], 1, i); break;
        case BK_OP_ECHO_MAP_SIZE: s.put_u32(e0 - s0); break;
        case BK_OP_ECHO_OVERLAP_SIZE: {
          const uint32_t x = e0 < b ? e0 : b, y = s0 > a ? s0 : a;
          s.put_u32(x > y ? x - y : 0);
          break;
        }
      }
    }
    if (op == BK_OP_ECHO_MAP_RANGE && !first) {
      const char* p = rtext + (rline[row] & kLineOffMask);  // the chromosome name: the reference row's token
      int         n = 0;
      while (is_tok((unsigned char)p[n])) n++;
      s.copy(p, n);
      s.put('\t');
      s.put_u32(mn);
      s.put('\t');
      s.put_u32(mx);
    }
    if (op >= BK_OP_VARIANCE) {
      if (nhit <= 1) {
        s.puts_("NAN", 3);
        return;
      }
      const double n = (double)nhit;
      const double var = ((n * vsq) - (vsum * vsum)) / (n * (n - 1.0));
      if (op == BK_OP_VARIANCE) {
        put_score(s, var, 1, i);
      } else if (op == BK_OP_STDEV) {
        put_score(s, sqrt(var), 1, i);
      } else {
        const double mean = vsum / n;
        if (mean == 0.0) s.puts_("NAN", 3);
        else put_score(s, sqrt(var) / mean, 1, i);
      }
      return;
    }
    if (op == BK_OP_BASES_UNIQ || op == BK_OP_BASES_UNIQ_F) {
      if (!first) {
        const uint32_t x = ce < b ? ce : b, y = cs > a ? cs : a;
        uniq += x > y ? x - y : 0;
      }
      if (op == BK_OP_BASES_UNIQ) s.put_u64(uniq);
      else put_score(s, (double)(uint32_t)uniq / (double)(b - a), 1, i);
    }
  }

  // --echo-map-id-uniq: the distinct ids of the qualifying map rows in strcmp order (std::set<std::string>,
  // ProcessBedVisitorRow.hpp:361-389).  Selection by repeated minimum: every round emits the smallest id greater than
  // the one emitted before -- quadratic in the hits, a rare operation on windows of tens of rows.
  __device__ __forceinline__ int cmp_id(const char* a, uint32_t la, const char* b, uint32_t lb) const {
    const uint32_t n = la < lb ? la : lb;
    for (uint32_t k = 0; k < n; k++) {
      const unsigned char x = (unsigned char)a[k], y = (unsigned char)b[k];
      if (x != y) return x < y ? -1 : 1;
    }
    return la == lb ? 0 : (la < lb ? -1 : 1);
  }
  template <class Sink>
  __device__ __noinline__ void uniq_ids(Sink& s, uint64_t i, uint64_t row) const {
    const uint64_t lo = win_lo[i], hi = lo + win_n[i];
    const uint32_t a = rs[row], b = re[row];
    const char*    last = nullptr;
    uint32_t       last_len = 0;
    while (true) {
      const char* best = nullptr;
      uint32_t    best_len = 0;
      for (uint64_t k = lo; k < hi; k++) {
        uint32_t ovl;
        if (!qualifies(ov, a, b, ms[k], me[k], ovl)) continue;
        const uint32_t sp = midspan[k];
        const char*    id = mtext + (mline[k] & kLineOffMask) + (sp >> 16);
        const uint32_t len = sp & 0xFFFFu;
        if (last && cmp_id(id, len, last, last_len) <= 0) continue;
        if (!best || cmp_id(id, len, best, best_len) < 0) {
          best = id;
          best_len = len;
        }
      }
      if (!best) break;
      if (last) s.puts_(mdelim, mdelim_len);
      s.copy(best, best_len);
      last = best;
      last_len = best_len;
    }
  }

  // --median / --kth: generalised median of the qualifying scores (RollingKthAverageVisitor.hpp:37-68).  The element
  // at a sorted position is found by rank counting (value, then file order): quadratic in the hits, a rare operation.
  __device__ __forceinline__ double score_at_rank(uint64_t lo, uint64_t hi, uint32_t a, uint32_t b, uint32_t want) const {
    for (uint64_t x = lo; x < hi; x++) {
      uint32_t ovl;
      if (!qualifies(ov, a, b, ms[x], me[x], ovl)) continue;
      const double vx = mscore[x];
      uint32_t     rank = 0;
      for (uint64_t y = lo; y < hi; y++) {
        if (!qualifies(ov, a, b, ms[y], me[y], ovl)) continue;
        const double vy = mscore[y];
        rank += (vy < vx || (vy == vx && y < x)) ? 1u : 0u;
      }
      if (rank == want) return vx;
    }
    return 0.0;
  }
  // |score - centre| at sorted position `want` among the qualifying rows (rank counting as above)
  __device__ __forceinline__ double dev_at_rank(uint64_t lo, uint64_t hi, uint32_t a, uint32_t b, uint32_t want, double centre) const {
    for (uint64_t x = lo; x < hi; x++) {
      uint32_t ovl;
      if (!qualifies(ov, a, b, ms[x], me[x], ovl)) continue;
      const double vx = fabs(mscore[x] - centre);
      uint32_t     rank = 0;
      for (uint64_t y = lo; y < hi; y++) {
        if (!qualifies(ov, a, b, ms[y], me[y], ovl)) continue;
        const double vy = fabs(mscore[y] - centre);
        rank += (vy < vx || (vy == vx && y < x)) ? 1u : 0u;
      }
      if (rank == want) return vx;
    }
    return 0.0;
  }
  // --mad [mult]: median of |x - median| times mult; NAN for fewer than two hits (MedianAbsoluteDeviationVisitor.hpp:57-93)
  template <class Sink>
  __device__ __noinline__ void mad_op(Sink& s, uint64_t i, uint64_t row, double mult) const {
    const uint32_t n = count[i];
    if (n <= 1) {
      s.puts_("NAN", 3);
      return;
    }
    const uint64_t lo = win_lo[i], hi = lo + win_n[i];
    const uint32_t a = rs[row], b = re[row];
    double         med, mad;
    if (n % 2 == 0) {
      med = (score_at_rank(lo, hi, a, b, n / 2 - 1) + score_at_rank(lo, hi, a, b, n / 2)) / 2.0;
      mad = (dev_at_rank(lo, hi, a, b, n / 2 - 1, med) + dev_at_rank(lo, hi, a, b, n / 2, med)) / 2.0;
    } else {
      med = score_at_rank(lo, hi, a, b, (n - 1) / 2);
      mad = dev_at_rank(lo, hi, a, b, n / 2, med);
    }
    put_score(s, mad * mult, 1, i);
  }

  template <class Sink>
  __device__ __noinline__ void kth_op(Sink& s, uint64_t i, uint64_t row, double k) const {
    const uint32_t n = count[i];
    if (n == 0) {
      s.puts_("NAN", 3);
      return;
    }
    const uint64_t lo = win_lo[i], hi = lo + win_n[i];
    const uint32_t a = rs[row], b = re[row];
    if (n == 1) {
      put_score(s, score_at_rank(lo, hi, a, b, 0), 1, i);
      return;
    }
    const double kn = k * (double)n;
    const double cl = ceil(kn), fl = floor(kn);
    uint32_t     up = (uint32_t)cl, down = (uint32_t)fl, pos = (uint32_t)((cl - kn > 0.5) ? fl : cl);  // iround
    if (up) up--;
    if (down) down--;
    if (pos) pos--;
    if (up == down) {  // "a true integer" (also whenever floor(k*n) is 0): the average of two neighbours
      const double one = score_at_rank(lo, hi, a, b, pos), two = score_at_rank(lo, hi, a, b, pos + 1);
      put_score(s, (one + two) / 2.0, 1, i);
    } else {
      put_score(s, score_at_rank(lo, hi, a, b, pos == up ? pos : pos + 1), 1, i);
    }
  }

  // --wmean: sum(w * score) / sum(w), w = overlap / reference length (WeightedAverageVisitor.hpp:55-70).  The reference
  // adds in heap-address order (std::set<MapType*>); here in file order -- equal up to the rounding of the sums.
  template <class Sink>
  __device__ __noinline__ void wmean_op(Sink& s, uint64_t i, uint64_t row) const {
    if (count[i] == 0) {
      s.puts_("NAN", 3);
      return;
    }
    const uint64_t lo = win_lo[i], hi = lo + win_n[i];
    const uint32_t a = rs[row], b = re[row];
    double         value = 0.0, wsum = 0.0;
    for (uint64_t k = lo; k < hi; k++) {
      uint32_t       ovl;
      const uint32_t s0 = ms[k], e0 = me[k];
      if (!qualifies(ov, a, b, s0, e0, ovl)) continue;
      const uint32_t x = e0 < b ? e0 : b, y = s0 > a ? s0 : a;
      const double   w = (double)(x > y ? x - y : 0u) / (double)(b - a);
      value += w * mscore[k];
      wsum += w;
    }
    if (wsum == 0.0 && value == 0.0) {  // --range hits that do not touch the row: 0/0, printed by printf as "-nan"
      s.puts_("-nan", 4);
      return;
    }
    put_score(s, value / wsum, 1, i);
  }

  // --tmean <low> <hi> (TrimmedMeanVisitor.hpp:93-141): the mean of the sorted scores after dropping the lowest
  // round(low*n) and the highest round(hi*n); sums in file order over the hits whose rank falls in the kept range.
  __device__ __forceinline__ static double iround(double d) {  // :231-238 (d >= 0 here)
    const double d1 = ceil(d);
    return (d1 - d > 0.5) ? floor(d) : d1;
  }
  template <class Sink>
  __device__ __noinline__ void tmean_op(Sink& s, uint64_t i, uint64_t row, double low, double high) const {
    const uint32_t n = count[i];
    if (n == 0) {
      s.puts_("NAN", 3);
      return;
    }
    const double eps = 2.220446049250313e-16;
    const bool   do_kth = fabs(1.0 - low - high) <= eps, symmetric = fabs(low - high) <= eps;
    uint64_t     pl = (uint64_t)iround(low * (double)n), ph = (uint64_t)n - (uint64_t)iround(high * (double)n);
    if (symmetric) {
      const uint64_t t = (uint64_t)n - ph;
      pl = pl > t ? pl : t;
      ph = (uint64_t)n - pl;
    }
    const bool do_low = pl > 0;
    if (do_low) --pl;
    if (ph > 0) --ph;
    const uint64_t lo = win_lo[i], hi = lo + win_n[i];
    const uint32_t a = rs[row], b = re[row];
    if (do_kth || (do_low && ph == pl)) {  // a single element
      put_score(s, score_at_rank(lo, hi, a, b, (uint32_t)ph), 1, i);
      return;
    }
    // ranks (value, then file order) of the kept range: (pl, ph] when something is trimmed below, [0, ph] otherwise
    double sum = 0.0;
    for (uint64_t x = lo; x < hi; x++) {
      uint32_t ovl;
      if (!qualifies(ov, a, b, ms[x], me[x], ovl)) continue;
      const double vx = mscore[x];
      uint64_t     rank = 0;
      for (uint64_t y = lo; y < hi; y++) {
        if (!qualifies(ov, a, b, ms[y], me[y], ovl)) continue;
        const double vy = mscore[y];
        rank += (vy < vx || (vy == vx && y < x)) ? 1u : 0u;
      }
      if (rank <= ph && (!do_low || rank > pl)) sum += vx;
    }
    put_score(s, do_low ? sum / (double)(ph - pl) : sum / (double)(ph + 1), 1, i);
  }

  // --max-element / --min-element (ExtremeVisitor.hpp:84-134 over ScoreThenGenomicCompare{Greater,Lesser},
  // BedCompare.hpp:263-288): the best score; among equal scores the genomically last (max) or first (min) row; among
  // equal rows the first in file order (the std::set keeps the first it was given).  Printed by PrintAllScorePrecision
  // (ProcessBedVisitorRow.hpp:181-222): chrom, start, end, id, score with --prec/--sci, rest of the line.
  template <class Sink>
  __device__ __noinline__ void element_op(Sink& s, bool want_max, uint64_t i, uint64_t row) const {
    const uint64_t lo = win_lo[i], hi = lo + win_n[i];
    const uint32_t a = rs[row], b = re[row];
    uint64_t       best = ~0ull;
    double         bv = 0.0;
    for (uint64_t k = lo; k < hi; k++) {
      uint32_t       ovl;
      const uint32_t s0 = ms[k], e0 = me[k];
      if (!qualifies(ov, a, b, s0, e0, ovl)) continue;
      const double v = mscore[k];
      bool         better = best == ~0ull;
      if (!better) {
        if (v != bv) better = want_max ? v > bv : v < bv;
        else if (s0 != ms[best]) better = want_max ? s0 > ms[best] : s0 < ms[best];
        else if (e0 != me[best]) better = want_max ? e0 > me[best] : e0 < me[best];
      }
      if (better) {
        best = k;
        bv = v;
      }
    }
    if (best == ~0ull) return;
    const char* p = mtext + (mline[best] & kLineOffMask);
    int         n = 0;
    while (is_tok((unsigned char)p[n])) n++;
    s.copy(p, n);
    s.put('\t');
    s.put_u32(ms[best]);
    s.put('\t');
    s.put_u32(me[best]);
    const char* q = p + n;
    for (int f = 0; f < 2; f++) {
      while (is_ws((unsigned char)*q)) q++;
      if (*q == '+') q++;
      while (is_digit((unsigned char)*q)) q++;
    }
    while (is_ws((unsigned char)*q)) q++;
    int idn = 0;
    while (is_tok((unsigned char)q[idn])) idn++;
    s.put('\t');
    s.copy(q, idn);
    q += idn;
    while (is_ws((unsigned char)*q)) q++;
    GlobalCursor gc{q};
    int64_t      adv = 0;
    double       dummy;
    parse_decimal(gc, adv, dummy);  // only to find where strtod stopped: the rest of the line starts there
    q += adv;
    s.put('\t');
    put_score(s, bv, 1, i);
    int m = 0;
    while (q[m] != '\n') m++;
    s.copy(q, m);
  }

  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& s) const {
    const uint32_t cnt = count[i];
    if (skip_unmapped && cnt == 0) return;  // MultiVisitor.hpp:84-85
    const uint64_t row = row0 + i;
    int            rowid_seen = 0;
    for (int c = 0; c < n_ops; c++) {
      if (c) s.puts_(delim, delim_len);
      switch (ops[c]) {
        case BK_OP_ECHO:
          if (!(RARE & 2) || ref_fields <= 3) echo_b3rest(s, rtext, rline[row], rs[row], re[row]);
          else echo_b45rest(s, rtext, rline[row], rs[row], re[row], ref_fields, ref_fields >= 5 ? rscore[row] : 0.0, scratch, i);
          break;
        case BK_OP_COUNT: s.put_u32(cnt); break;
        case BK_OP_INDICATOR: s.put(cnt ? '1' : '0'); break;
        case BK_OP_BASES: s.put_u64(bases[i]); break;
        case BK_OP_SUM: put_score(s, sum[i], cnt, i); break;
        case BK_OP_MEAN: put_score(s, cnt ? sum[i] / (double)(int)cnt : 0.0, cnt, i); break;
        case BK_OP_MAX: put_score(s, vmax[i], cnt, i); break;
        case BK_OP_MIN: put_score(s, vmin[i], cnt, i); break;
        case BK_OP_ECHO_MAP_ID: {
          if (Sink::counting) {
            s.copy(nullptr, idbytes[i]);
            break;
          }
          if (idfill_off) {  // the ids are written by a warp per row afterwards (k_fill_ids): leave the room, note the place
            idfill_off[i] = s.gpos();
            s.skip(idbytes[i]);
            break;
          }
          const uint64_t lo = win_lo[i], hi = lo + win_n[i];
          const uint32_t a = rs[row], b = re[row];
          bool           first = true;
          for (uint64_t k = lo; k < hi; k++) {
            uint32_t ovl;
            if (!qualifies(ov, a, b, ms[k], me[k], ovl)) continue;
            if (!first) s.puts_(mdelim, mdelim_len);
            first = false;
            const uint32_t sp = midspan[k];
            s.copy(mtext + (mline[k] & kLineOffMask) + (sp >> 16), sp & 0xFFFFu);
          }
          break;
        }
        case BK_OP_ECHO_MAP: case BK_OP_ECHO_MAP_SCORE: case BK_OP_ECHO_MAP_SIZE: case BK_OP_ECHO_OVERLAP_SIZE:
        case BK_OP_ECHO_MAP_RANGE: case BK_OP_BASES_UNIQ: case BK_OP_BASES_UNIQ_F:
        case BK_OP_VARIANCE: case BK_OP_STDEV: case BK_OP_CV:
          if (RARE & 4) window_op(s, ops[c], i, row);
          break;
        case BK_OP_ECHO_MAP_ID_UNIQ:
          if (RARE & 4) uniq_ids(s, i, row);
          break;
        case BK_OP_MEDIAN: case BK_OP_KTH:
          if (RARE & 4) kth_op(s, i, row, ops[c] == BK_OP_MEDIAN ? 0.5 : kth_arg[c]);
          break;
        case BK_OP_MAD:
          if (RARE & 4) mad_op(s, i, row, kth_arg[c] > 0.0 ? kth_arg[c] : 1.0);
          break;
        case BK_OP_WMEAN:
          if (RARE & 4) wmean_op(s, i, row);
          break;
        case BK_OP_TMEAN:
          if (RARE & 4) tmean_op(s, i, row, kth_arg[c], arg2[c]);
          break;
        case BK_OP_MAX_ELEMENT: case BK_OP_MIN_ELEMENT:
          if (RARE & 4) {
            if (i == stop_row) return;  // the reference threw here: the row ends after the delimiter, without NL
            element_op(s, ops[c] == BK_OP_MAX_ELEMENT, i, row);
          }
          break;
        case BK_OP_ECHO_REF_SIZE: s.put_u32(re[row] - rs[row]); break;
        case BK_OP_ECHO_REF_NAME: {
          const char* p = rtext + (rline[row] & kLineOffMask);
          int         n = 0;
          while (is_tok((unsigned char)p[n])) n++;
          s.copy(p, n);
          s.put(':');
          s.put_u32(rs[row]);
          s.put('-');
          s.put_u32(re[row]);
          break;
        }
        case BK_OP_ECHO_REF_ROW_ID:
          // PrintRowID: a static counter bumped once per printed id (ProcessBedVisitorRow.hpp:347-354) -- rows the
          // command does not print (--skip-unmapped, other chromosomes under --chrom) do not count
          s.puts_("id-", 3);
          s.put_u64(rowid_base + (rank ? rank[i] : i) * (uint64_t)rowid_ops + (uint64_t)(++rowid_seen));
          break;
      }
    }
    s.put('\n');
  }
};

// --echo-map-id, cooperative form (EchoMapBedVisitor.hpp:39-66): one warp per reference row.  The lanes test 32 window
// rows at a time, a warp scan over (id length + delimiter) gives every hit its place in the column, and each lane copies
// its own id: neighbouring lanes write neighbouring bytes, instead of one thread walking ~70 ids of ~10 bytes (the
// per-row emitter leaves the column's room empty and records where it starts).
struct FillIdsParams {
  const uint32_t* rs;
  const uint32_t* re;
  uint64_t        row0, n;
  const uint32_t* ms;
  const uint32_t* me;
  const uint32_t* midspan;
  const uint64_t* mline;
  const char*     mtext;
  const uint64_t* win_lo;
  const uint32_t* win_n;
  const uint32_t* idbytes;
  const uint64_t* col_off;
  OverlapSpec     ov;
  char            mdelim[24];
  uint32_t        mdelim_len;
  char*           out;
};
__global__ void __launch_bounds__(256) k_fill_ids(FillIdsParams p) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * 256) >> 5;
  for (uint64_t i = w0; i < p.n; i += nw) {
    if (p.idbytes[i] == 0) continue;
    const uint64_t row = p.row0 + i;
    const uint32_t a = p.rs[row], b = p.re[row];
    const uint64_t lo = p.win_lo[i], hi = lo + p.win_n[i];
    char*          dst = p.out + p.col_off[i];
    uint32_t       run = 0;
    bool           none_yet = true;
    for (uint64_t k0 = lo; k0 < hi; k0 += 32) {
      const uint64_t k = k0 + lane;
      uint32_t       ovl, len = 0;
      bool           q = false;
      uint32_t       sp = 0;
      if (k < hi && qualifies(p.ov, a, b, p.ms[k], p.me[k], ovl)) {
        q = true;
        sp = p.midspan[k];
        len = sp & 0xFFFFu;
      }
      const unsigned m = __ballot_sync(0xffffffffu, q);
      if (m == 0) continue;
      const bool     lead = none_yet && q && (m & ((1u << lane) - 1u)) == 0;  // the very first id has no delimiter in front
      const uint32_t contrib = q ? len + (lead ? 0u : p.mdelim_len) : 0u;
      const uint32_t incl = warp_incl_scan(contrib);
      if (q) {
        char* d = dst + run + (incl - contrib);
        if (!lead)
          for (uint32_t t = 0; t < p.mdelim_len; t++) *d++ = p.mdelim[t];
        const char* src = p.mtext + (p.mline[k] & kLineOffMask) + (sp >> 16);
        for (uint32_t t = 0; t < len; t++) d[t] = src[t];
      }
      run += __shfl_sync(0xffffffffu, incl, 31);
      none_yet = false;
    }
  }
}

// first row without a mapped element: scratch[SC_COUNT_A] = ~row (atomicMax; 0 = none)
__global__ void __launch_bounds__(256) k_first_unmapped(const uint32_t* __restrict__ count, uint64_t n, uint64_t* scratch) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n && count[i] == 0) atomicMax(reinterpret_cast<unsigned long long*>(&scratch[SC_COUNT_A]), ~(unsigned long long)i);
}

// rank[i] = number of rows j < i with count[j] > 0 (the rows --skip-unmapped prints): warp ranges of 2048 rows,
// range totals -> k_scan_totals -> ranks
constexpr int RK_RANGE = 2048;
__global__ void __launch_bounds__(256) k_rank_totals(const uint32_t* __restrict__ count, uint64_t n, uint64_t* __restrict__ tot) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * RK_RANGE, b = a + RK_RANGE < n ? a + RK_RANGE : n;
  if (a >= n) return;
  uint32_t c = 0;
  for (uint64_t k = a + lane; k < b; k += 32) c += count[k] ? 1u : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) tot[w] = c;
}
__global__ void __launch_bounds__(256) k_rank_write(const uint32_t* __restrict__ count, uint64_t n, const uint64_t* __restrict__ base,
                                                    uint64_t* __restrict__ rank) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * RK_RANGE, b = a + RK_RANGE < n ? a + RK_RANGE : n;
  if (a >= n) return;
  uint64_t run = base[w];
  for (uint64_t k0 = a; k0 < b; k0 += 32) {
    const uint64_t k = k0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, k < b && count[k] != 0);
    if (k < b) rank[k] = run + __popc(m & ((1u << lane) - 1u));
    run += __popc(m);
  }
}

// ---------------------------------------------------------------------------------------------------------
int finish_text(bk_ctx* ctx, char* d_out, uint64_t bytes, uint64_t rows, int on_device, bk_text* out);

static int find_run(const bk_bed* b, const std::string& name) {
  // runs are in strcmp order: binary search
  int lo = 0, hi = (int)b->runs.size();
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    int c = strcmp(b->runs[mid].name.c_str(), name.c_str());
    if (c == 0) return mid;
    if (c < 0) lo = mid + 1; else hi = mid;
  }
  return -1;
}

}  // namespace bk

using namespace bk;

extern "C" void bk_mapspec_default(bk_mapspec* spec) {
  memset(spec, 0, sizeof(*spec));
  spec->overlap_kind = BK_OVR_BP;
  spec->overlap_bp = 1;
  spec->precision = 6;
  spec->delim = "|";
  spec->multidelim = ";";
}

extern "C" int bk_bedmap(bk_ctx* ctx, const bk_bed* ref, const bk_bed* map, const bk_mapspec* spec, bk_text* out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !ref || !spec || !out) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  if (!map) map = ref;
  if (spec->n_ops <= 0 || spec->n_ops > BK_MAX_OPS) return fail(ctx, BK_ERR_ARG, "No processing option specified (ie; --max).");
  if (spec->precision < 0 || spec->precision > (spec->sci ? 17 : 18))
    return fail(ctx, BK_ERR_UNSUPPORTED, "--prec %d: the exact device formatter supports 0..%d", spec->precision, spec->sci ? 17 : 18);
  const char* delim = spec->delim ? spec->delim : "|";
  const char* mdelim = spec->multidelim ? spec->multidelim : ";";
  if (strlen(delim) > 23 || strlen(mdelim) > 23) return fail(ctx, BK_ERR_UNSUPPORTED, "delimiter longer than 23 bytes");

  unsigned need = 0;
  int      rowid_ops = 0;
  bool     need_echo = false, need_refline = false, need_ids = false, need_mapline = false, need_mapscore = false;
  bool     window_ops = false, element_ops = false;
  for (int c = 0; c < spec->n_ops; c++) {
    switch (spec->ops[c]) {
      case BK_OP_ECHO: need_echo = true; need_refline = true; break;
      case BK_OP_ECHO_REF_NAME: need_refline = true; break;
      case BK_OP_COUNT: case BK_OP_INDICATOR: case BK_OP_ECHO_REF_SIZE: break;
      case BK_OP_ECHO_REF_ROW_ID: rowid_ops++; break;
      case BK_OP_BASES: need |= NEED_BASES; break;
      case BK_OP_SUM: case BK_OP_MEAN: need |= NEED_SUM; break;
      case BK_OP_MAX: need |= NEED_MAX; break;
      case BK_OP_MIN: need |= NEED_MIN; break;
      case BK_OP_ECHO_MAP_ID: need |= NEED_IDS; need_ids = true; break;
      case BK_OP_ECHO_MAP: need |= NEED_IDS; need_mapline = true; window_ops = true; break;
      case BK_OP_ECHO_MAP_SCORE: need |= NEED_IDS; need_mapscore = true; window_ops = true; break;
      case BK_OP_ECHO_MAP_RANGE: need |= NEED_IDS; need_refline = true; window_ops = true; break;
      case BK_OP_ECHO_MAP_SIZE: case BK_OP_ECHO_OVERLAP_SIZE: case BK_OP_BASES_UNIQ: case BK_OP_BASES_UNIQ_F:
        need |= NEED_IDS; window_ops = true; break;
      case BK_OP_ECHO_MAP_ID_UNIQ: need |= NEED_IDS; need_ids = true; window_ops = true; break;
      case BK_OP_KTH:
        if (!(spec->op_arg[c] > 0.0 && spec->op_arg[c] < 1.0))
          return fail(ctx, BK_ERR_UNSUPPORTED, "--kth %g: the device path supports 0 < val < 1", spec->op_arg[c]);
        need |= NEED_IDS; need_mapscore = true; window_ops = true; break;
      case BK_OP_MEDIAN: need |= NEED_IDS; need_mapscore = true; window_ops = true; break;
      case BK_OP_MAD:
        if (spec->op_arg[c] < 0.0) return fail(ctx, BK_ERR_ARG, "--mad Expect 0 < val");
        need |= NEED_IDS; need_mapscore = true; window_ops = true; break;
      case BK_OP_VARIANCE: case BK_OP_STDEV: case BK_OP_CV: need |= NEED_IDS; need_mapscore = true; window_ops = true; break;
      case BK_OP_WMEAN: need |= NEED_IDS; need_mapscore = true; window_ops = true; break;
      case BK_OP_TMEAN: {  // TrimmedMean's constructor, TrimmedMeanVisitor.hpp:60-73
        const double lo = spec->op_arg[c], hi = spec->op_arg2[c];
        if (!(lo >= 0 && lo <= 1)) return fail(ctx, BK_ERR_ARG, "Expect 0 <= lowerKth <= 1");
        if (!(hi >= 0 && hi <= 1)) return fail(ctx, BK_ERR_ARG, "Expect 0 <= upperKth <= 1");
        if (!(lo + hi <= 1 + 2.220446049250313e-16)) return fail(ctx, BK_ERR_ARG, "Expect lowerKth + upperKth <= 1");
        need |= NEED_IDS; need_mapscore = true; window_ops = true; break;
      }
      case BK_OP_MAX_ELEMENT: case BK_OP_MIN_ELEMENT:
        need |= NEED_IDS; need_mapscore = true; need_mapline = true; window_ops = true; element_ops = true; break;
      default: return fail(ctx, BK_ERR_UNSUPPORTED, "bedmap operation %d is outside the device hot path", spec->ops[c]);
    }
  }
  if (need_refline && !ref->line_off && ref->nrows)
    return fail(ctx, BK_ERR_ARG, "reference file was loaded without BK_COL_LINE but --echo needs it");
  if ((need & (NEED_SUM | NEED_MAX | NEED_MIN)) && !map->score && map->nrows)
    return fail(ctx, BK_ERR_ARG, "map file was loaded without BK_COL_SCORE but a score operation needs it");
  if (need_ids && (!map->idspan || !map->line_off) && map->nrows)
    return fail(ctx, BK_ERR_ARG, "map file was loaded without BK_COL_ID|BK_COL_LINE but --echo-map-id needs it");
  if (need_mapline && !map->line_off && map->nrows)
    return fail(ctx, BK_ERR_ARG, "map file was loaded without BK_COL_LINE but --echo-map needs it");
  if (need_mapscore && !map->score && map->nrows)
    return fail(ctx, BK_ERR_ARG, "map file was loaded without BK_COL_SCORE but a score operation needs it");

  OverlapSpec ov{};
  ov.kind = spec->overlap_kind;
  switch (spec->overlap_kind) {
    case BK_OVR_BP:
    case BK_OVR_RANGE:
      if (spec->overlap_bp > 0xFFFFFFFEull) return fail(ctx, BK_ERR_UNSUPPORTED, "overlap/range value beyond 32 bits");
      ov.bp = (uint32_t)spec->overlap_bp;
      break;
    case BK_OVR_FRAC_REF: case BK_OVR_FRAC_MAP: case BK_OVR_FRAC_EITHER: case BK_OVR_FRAC_BOTH: {
      double perc = spec->overlap_frac;  // PercentOverlapMapping ctor, BedDistances.hpp:120-131
      while (perc > 1) perc /= 10.0;
      perc -= 2.220446049250313e-16;
      if (perc <= 0.0) perc = 2.220446049250313e-16;
      ov.frac = perc;
      break;
    }
    case BK_OVR_EXACT: break;
    default: return fail(ctx, BK_ERR_ARG, "unknown overlap kind %d", spec->overlap_kind);
  }

  // reference rows to process and the chromosome pairing table
  const bool all = !spec->chrom || !*spec->chrom || strcmp(spec->chrom, "all") == 0;
  std::vector<uint64_t> rrb, rmb, rme;
  uint64_t row0 = 0, row1 = 0;
  if (all) {
    row0 = 0;
    row1 = ref->nrows;
    for (auto& r : ref->runs) {
      rrb.push_back(r.row_begin);
      int j = find_run(map, r.name);
      rmb.push_back(j >= 0 ? map->runs[j].row_begin : 0);
      rme.push_back(j >= 0 ? map->runs[j].row_end : 0);
    }
  } else {
    int jr = find_run(ref, spec->chrom);
    if (jr >= 0) {
      row0 = ref->runs[jr].row_begin;
      row1 = ref->runs[jr].row_end;
      rrb.push_back(row0);
      int j = find_run(map, spec->chrom);
      rmb.push_back(j >= 0 ? map->runs[j].row_begin : 0);
      rme.push_back(j >= 0 ? map->runs[j].row_end : 0);
    }
  }
  const uint64_t n = row1 - row0;
  if (n == 0) return finish_text(ctx, nullptr, 0, 0, spec->out_on_device, out);
  rrb.push_back(row1);
  const int nruns = (int)rmb.size();

  BK_TRY(ensure_pmax(ctx, map));

  uint64_t* d_tab = dalloc<uint64_t>(ctx, 3 * (size_t)nruns + 1);
  if (!d_tab) return BK_ERR_NOMEM;
  std::vector<uint64_t> tab;
  tab.insert(tab.end(), rrb.begin(), rrb.end());
  tab.insert(tab.end(), rmb.begin(), rmb.end());
  tab.insert(tab.end(), rme.begin(), rme.end());
  BK_TRY(upload_params(ctx, d_tab, tab.data(), tab.size() * 8));

  MapStatsParams sp{};
  sp.rs = ref->start; sp.re = ref->end; sp.row0 = row0; sp.n = n;
  sp.run_ref_begin = d_tab; sp.run_map_begin = d_tab + nruns + 1; sp.run_map_end = d_tab + 2 * nruns + 1; sp.nruns = nruns;
  sp.ms = map->start; sp.me = map->end; sp.pm = map->pmax_end; sp.score = map->score;
  sp.idspan = map->idspan; sp.map_rows = map->nrows;
  sp.ov = ov; sp.need = need; sp.mdelim_len = (uint32_t)strlen(mdelim);
  sp.count = dalloc<uint32_t>(ctx, n);
  if (need & NEED_BASES) sp.bases = dalloc<uint64_t>(ctx, n);
  if (need & NEED_SUM) sp.sum = dalloc<double>(ctx, n);
  if (need & NEED_MAX) sp.vmax = dalloc<double>(ctx, n);
  if (need & NEED_MIN) sp.vmin = dalloc<double>(ctx, n);
  if (need & NEED_IDS) {
    sp.win_lo = dalloc<uint64_t>(ctx, n);
    sp.win_n = dalloc<uint32_t>(ctx, n);
    sp.idbytes = dalloc<uint32_t>(ctx, n);
  }
  sp.scratch = ctx->d_scratch;
  if (!sp.count || ((need & NEED_BASES) && !sp.bases) || ((need & NEED_SUM) && !sp.sum) || ((need & NEED_MAX) && !sp.vmax) ||
      ((need & NEED_MIN) && !sp.vmin) || ((need & NEED_IDS) && (!sp.win_lo || !sp.win_n || !sp.idbytes)))
    return BK_ERR_NOMEM;
  {
    const uint64_t batches = (n + 31) / 32, per_block = MS_THREADS / 32;
    uint64_t       blocks = (batches + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t)kSMs * 8 * 8;  // several resident waves; warps stride over the batches
    if (blocks > cap) blocks = cap;
    // the lane-per-row form (k_map_group) is the default: 2.07 ms against 3.04 ms at configuration 2, 26.9 against 43.3 ms
    // (block-skipping scan) at configuration 5; the warp-per-row forms serve map files of 2^32 rows and more and
    // BEDKIT_MAP_KERNEL=row (A/B measurements, tests)
    const char* const mk_env = getenv("BEDKIT_MAP_KERNEL");
    const bool group = !(mk_env && mk_env[0] == 'r') && map->nrows > 0 && map->nrows < 0xFFFFFF00ull;
    // the default criterion (--bp-ovr) gets its own instantiation; the other six share the generic predicate
    // dense map files (>= 32 map rows per reference row: candidate windows of several hundred rows): both forms skip
    // dead 32-row blocks with the block-max-end index (warp-per-row: its own instantiation, ~20 % slower on short windows)
    const bool dense = ov.kind == BK_OVR_BP && !(need & NEED_IDS) && map->nrows / 32 >= n;
    if (dense) {
      BK_TRY(ensure_bmax(ctx, map));
      sp.bmax = map->bmax_end;
    }
    prof_begin(ctx, group ? "k_map_group" : "k_map_stats");
    if (group && ov.kind == BK_OVR_BP) launch_map_group<BK_OVR_BP>(need, (unsigned)blocks, ctx->stream, sp);
    else if (group) launch_map_group<-1>(need, (unsigned)blocks, ctx->stream, sp);
    else if (dense) launch_map_stats<BK_OVR_BP, true>(need, (unsigned)blocks, ctx->stream, sp);
    else if (ov.kind == BK_OVR_BP) launch_map_stats<BK_OVR_BP, false>(need, (unsigned)blocks, ctx->stream, sp);
    else launch_map_stats<-1, false>(need, (unsigned)blocks, ctx->stream, sp);
    prof_end(ctx);
    BK_LAUNCHED(ctx);
  }
  dfree(ctx, d_tab);  // stream-ordered reuse; the table went through the pinned ring


  const uint64_t dl = strlen(delim);
  const uint64_t cap = 0;  // the emitter sizes the result itself (length pass)
  uint64_t*      d_rank = nullptr;
  if (rowid_ops && spec->skip_unmapped) {
    const uint64_t nranges = (n + RK_RANGE - 1) / RK_RANGE;
    d_rank = dalloc<uint64_t>(ctx, n);
    uint64_t* d_tot = dalloc<uint64_t>(ctx, nranges);
    uint64_t* d_base = dalloc<uint64_t>(ctx, nranges + 1);
    if (!d_rank || !d_tot || !d_base) return BK_ERR_NOMEM;
    const unsigned blocks = (unsigned)((nranges + 7) / 8);
    k_rank_totals<<<blocks, 256, 0, ctx->stream>>>(sp.count, n, d_tot);
    BK_LAUNCHED(ctx);
    k_scan_totals<SC_COUNT_D><<<1, 1024, 0, ctx->stream>>>(d_tot, d_base, (uint32_t)nranges, ctx->d_scratch);
    BK_LAUNCHED(ctx);
    k_rank_write<<<blocks, 256, 0, ctx->stream>>>(sp.count, n, d_base, d_rank);
    BK_LAUNCHED(ctx);
    dfree(ctx, d_tot);
    dfree(ctx, d_base);
  }

  // --max-element / --min-element throw on a row without mapped elements (unless --skip-unmapped drops it first): the
  // reference has by then printed the earlier rows and this row up to the element's column.  Find that row.
  uint64_t stop_row = ~0ull, n_emit = n;
  if (element_ops && !spec->skip_unmapped) {
    BK_TRY(reset_scratch(ctx));
    k_first_unmapped<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(sp.count, n, ctx->d_scratch);
    BK_LAUNCHED(ctx);
    BK_TRY(read_scratch(ctx));
    if (ctx->h_scratch[SC_COUNT_A]) {
      stop_row = ~ctx->h_scratch[SC_COUNT_A];
      n_emit = stop_row + 1;
    }
  }

  // one --echo-map-id column: the per-row emitter leaves its room empty, a warp per row fills it afterwards
  uint64_t* d_idfill = nullptr;
  {
    int id_cols = 0;
    for (int c = 0; c < spec->n_ops; c++) id_cols += spec->ops[c] == BK_OP_ECHO_MAP_ID;
    if (id_cols == 1 && stop_row == ~0ull) {
      d_idfill = dalloc<uint64_t>(ctx, n);
      if (!d_idfill) return BK_ERR_NOMEM;
    }
  }

  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  int      rc = BK_OK;
  auto emit = [&](auto fn) {
    fn.rtext = ref->d_text; fn.rline = ref->line_off; fn.rs = ref->start; fn.re = ref->end; fn.row0 = row0;
    fn.ref_fields = (ref->min_fields >= 5 && !ref->score) ? 4 : ref->min_fields;
    fn.rscore = ref->score;
    fn.mtext = map->d_text; fn.mline = map->line_off; fn.midspan = map->idspan; fn.ms = map->start; fn.me = map->end;
    fn.mscore = map->score;
    fn.map_fields = (map->min_fields >= 5 && !map->score) ? 4 : map->min_fields;
    fn.count = sp.count; fn.bases = sp.bases; fn.sum = sp.sum; fn.vmax = sp.vmax; fn.vmin = sp.vmin;
    fn.win_lo = sp.win_lo; fn.win_n = sp.win_n; fn.idbytes = sp.idbytes;
    fn.rank = d_rank; fn.rowid_base = spec->row_id_base; fn.rowid_ops = rowid_ops;
    fn.ov = ov; fn.n_ops = spec->n_ops;
    for (int c = 0; c < spec->n_ops; c++) {
      fn.ops[c] = (unsigned char)spec->ops[c];
      fn.kth_arg[c] = spec->op_arg[c];
      fn.arg2[c] = spec->op_arg2[c];
    }
    fn.stop_row = stop_row;
    fn.idfill_off = d_idfill;
    fn.prec = spec->precision; fn.sci = spec->sci; fn.skip_unmapped = spec->skip_unmapped;
    fn.delim_len = (int)dl; memcpy(fn.delim, delim, dl);
    fn.mdelim_len = (int)strlen(mdelim); memcpy(fn.mdelim, mdelim, fn.mdelim_len);
    fn.scratch = ctx->d_scratch;
    rc = run_emit(ctx, fn, n_emit, cap, &d_out, &bytes, &rows);
  };
  const bool rare_fields = need_echo && ref->min_fields > 3;
  if (window_ops) emit(BedmapRow<7>{});  // the list operations are not the hot path: one instantiation carries everything
  else if (!spec->sci && !rare_fields) emit(BedmapRow<0>{});
  else if (spec->sci && !rare_fields) emit(BedmapRow<1>{});
  else if (!spec->sci) emit(BedmapRow<2>{});
  else emit(BedmapRow<3>{});
  if (rc == BK_OK && d_idfill && bytes) {
    FillIdsParams fp{};
    fp.rs = ref->start; fp.re = ref->end; fp.row0 = row0; fp.n = n;
    fp.ms = map->start; fp.me = map->end; fp.midspan = map->idspan; fp.mline = map->line_off; fp.mtext = map->d_text;
    fp.win_lo = sp.win_lo; fp.win_n = sp.win_n; fp.idbytes = sp.idbytes; fp.col_off = d_idfill; fp.ov = ov;
    fp.mdelim_len = (uint32_t)strlen(mdelim); memcpy(fp.mdelim, mdelim, fp.mdelim_len);
    fp.out = d_out;
    const uint64_t want = (n + 7) / 8, cap_blocks = (uint64_t)kSMs * 64;
    prof_begin(ctx, "k_fill_ids");
    k_fill_ids<<<(unsigned)(want < cap_blocks ? want : cap_blocks), 256, 0, ctx->stream>>>(fp);
    prof_end(ctx);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) rc = BK_ERR_CUDA;
  }
  dfree(ctx, d_idfill);
  dfree(ctx, sp.count); dfree(ctx, sp.bases); dfree(ctx, sp.sum); dfree(ctx, sp.vmax); dfree(ctx, sp.vmin);
  dfree(ctx, sp.win_lo); dfree(ctx, sp.win_n); dfree(ctx, sp.idbytes); dfree(ctx, d_rank);
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return rc;
  }
  rc = finish_text(ctx, d_out, bytes, rows, spec->out_on_device, out);
  if (rc == BK_OK && stop_row != ~0ull) return fail(ctx, BK_ERR_NAN_ELEMENT, "Unable to process a 'NAN' with PrintAllScorePrecision.");
  return rc;
}
