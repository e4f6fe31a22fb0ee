// emit.cuh -- the BED writer skeleton shared by every tool: one output row per thread, two passes.
//
//   k_emit_len:  length of every row (CountSink), warp-local prefix sums over contiguous row ranges
//   k_scan_totals: one CTA, range totals -> range bases (and the output size)
//   k_emit:      write pass (MemSink) into a shared-memory stage laid out with the same 16-byte phase as the
//                destination -> 16-byte coalesced streaming stores to HBM.  Tiles are independent (no look-back).
//
// RowFn contract:  template <class Sink> __device__ void operator()(uint64_t i, Sink& s) const;
//                  writes row i (including its '\n'), or nothing at all if the row is suppressed.
#pragma once
#include "common.cuh"
#include "fmt.cuh"
#include "strtod_exact.cuh"

namespace bk {

constexpr int E_THREADS = 256;
constexpr int E_STAGE = 32 * 1024;  // bytes of shared staging per tile (tiles that exceed it write straight to HBM)

#ifdef __CUDACC__

// echo a B3Rest row: chrom \t start \t end <rest>   (Bed.hpp:316-320, :376-378): numbers are re-printed from the
// parsed values, the rest of the line (including its leading tab) is copied verbatim.
template <class Sink>
__device__ __noinline__ void echo_b3rest_slow(Sink& s, const char* __restrict__ p, uint32_t st, uint32_t en);

template <class Sink>
__device__ __forceinline__ void echo_b3rest(Sink& s, const char* __restrict__ text, uint64_t packed_off, uint32_t st, uint32_t en) {
  const char*    p = text + (packed_off & kLineOffMask);
  const uint32_t len = (uint32_t)(packed_off >> 48);
  if (len != 0xFFFFu) {  // canonical line: re-printing reproduces the input bytes (decided by the parser)
    s.copy(p, len);
    return;
  }
  echo_b3rest_slow(s, p, st, en);
}

template <class Sink>
__device__ __noinline__ void echo_b3rest_slow(Sink& s, const char* __restrict__ p, uint32_t st, uint32_t en) {
  int n = 0;
  while (is_tok((unsigned char)p[n])) n++;
  s.copy(p, n);
  s.put('\t');
  s.put_u32(st);
  s.put('\t');
  s.put_u32(en);
  const char* q = p + n;
  while (is_ws((unsigned char)*q)) q++;
  if (*q == '+') q++;
  while (is_digit((unsigned char)*q)) q++;
  while (is_ws((unsigned char)*q)) q++;
  if (*q == '+') q++;
  while (is_digit((unsigned char)*q)) q++;
  int m = 0;
  while (q[m] != '\n') m++;
  s.copy(q, m);
}

// echo a B4Rest / B5Rest row (single-file bedmap, where the reference file has the map's record type,
// Bedmap.cpp:676-700): "%s\t%lu\t%lu\t%s%s" / "%s\t%lu\t%lu\t%s\t%lf%s" (Bed.hpp:640-646, :740-743, :896-903):
// id re-printed, column 5 re-printed with "%lf" (six decimals, whatever --prec says), the remainder verbatim.
struct GlobalCursor {
  const char* p;
  __device__ __forceinline__ unsigned char at(int64_t q) const { return (unsigned char)p[q]; }
};
template <class Sink>
__device__ __noinline__ void echo_b45rest(Sink& s, const char* __restrict__ text, uint64_t packed_off, uint32_t st, uint32_t en,
                                             int fields, double score, uint64_t* scratch, uint64_t row) {
  const char* p = text + (packed_off & kLineOffMask);
  if (fields == 4 && (uint32_t)(packed_off >> 48) != 0xFFFFu) {
    s.copy(p, (uint32_t)(packed_off >> 48));
    return;
  }
  int n = 0;
  while (is_tok((unsigned char)p[n])) n++;
  s.copy(p, n);
  s.put('\t');
  s.put_u32(st);
  s.put('\t');
  s.put_u32(en);
  const char* q = p + n;
  while (is_ws((unsigned char)*q)) q++;
  if (*q == '+') q++;
  while (is_digit((unsigned char)*q)) q++;
  while (is_ws((unsigned char)*q)) q++;
  if (*q == '+') q++;
  while (is_digit((unsigned char)*q)) q++;
  while (is_ws((unsigned char)*q)) q++;
  int idn = 0;
  while (is_tok((unsigned char)q[idn])) idn++;
  s.put('\t');
  s.copy(q, idn);
  q += idn;
  if (fields >= 5) {
    while (is_ws((unsigned char)*q)) q++;
    GlobalCursor gc{q};
    int64_t      adv = 0;
    double       dummy;
    parse_decimal(gc, adv, dummy);  // only to find where strtod stops
    q += adv;
    s.put('\t');
    Fixed f;
    if (!to_fixed(score, 6, f)) {
      dev_set_error(scratch, BK_ERR_UNSUPPORTED, row);
      s.put('?');
    } else {
      put_fixed(s, f, 6);
    }
  }
  int m = 0;
  while (q[m] != '\n') m++;
  s.copy(q, m);
}

// Pass 1: byte length of every row.  A warp owns `rows_per_warp` consecutive rows (a multiple of E_THREADS, so an
// output tile never straddles two warp ranges), 32 rows per step; local_off[i] = bytes of the earlier rows of the
// same warp range, warp_total[w] = bytes of the range.  No block barrier, no atomics on the data path.
template <class RowFn>
__global__ void __launch_bounds__(E_THREADS) k_emit_len(RowFn fn, uint64_t n, uint32_t rows_per_warp,
                                                        uint32_t* __restrict__ local_off, uint64_t* __restrict__ warp_total,
                                                        uint32_t nwarps, uint64_t* scratch) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w0 = ((uint64_t)blockIdx.x * E_THREADS + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * E_THREADS) >> 5;
  for (uint64_t w = w0; w < nwarps; w += nw) {
    const uint64_t a = w * rows_per_warp, b = a + rows_per_warp < n ? a + rows_per_warp : n;
    uint64_t       run = 0;
    uint32_t       nz = 0;
    for (uint64_t r0 = a; r0 < b; r0 += 32) {
      const uint64_t i = r0 + lane;
      uint64_t       len64 = 0;
      if (i < b) {
        CountSink cs;
        fn(i, cs);
        len64 = cs.n;
      }
      if (len64 >> 32) dev_set_error(scratch, BK_ERR_UNSUPPORTED, i);  // a single row of 4 GiB: not representable
      const uint32_t len = (uint32_t)len64;
      const uint32_t incl = warp_incl_scan(len);
      if (i < b) local_off[i] = (uint32_t)run + (incl - len);
      run += __shfl_sync(0xffffffffu, incl, 31);
      nz += __popc(__ballot_sync(0xffffffffu, len != 0));
    }
    if (lane == 0) {
      warp_total[w] = run;
      if (run >> 32) dev_set_error(scratch, BK_ERR_UNSUPPORTED, a);
      if (nz) atomicAdd(reinterpret_cast<unsigned long long*>(&scratch[SC_OUT_ROWS]), (unsigned long long)nz);
    }
  }
}

// Pass 2: the byte offset of every row is known (warp_base[range] + local_off[row]), so tiles are independent: every
// thread formats its row into the shared-memory stage at its final position, then the tile leaves with 16-byte stores.
template <class RowFn>
__global__ void __launch_bounds__(E_THREADS, 4) k_emit(RowFn fn, uint64_t n, char* __restrict__ out, uint64_t out_cap,
                                                       int range_shift, const uint32_t* __restrict__ local_off,
                                                       const uint64_t* __restrict__ warp_base, uint32_t nwarps,
                                                       uint32_t ntiles, uint64_t* scratch) {
  __shared__ __align__(16) char stage[E_STAGE + 16];
  const int tid = threadIdx.x;
  auto offset_of = [&](uint64_t r) -> uint64_t {  // first output byte of row r (r == n: the total)
    return r >= n ? warp_base[nwarps] : warp_base[r >> range_shift] + local_off[r];
  };
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint64_t row0 = (uint64_t)tile * E_THREADS, i = row0 + tid;
    const uint64_t base = offset_of(row0);
    const uint32_t total = (uint32_t)(offset_of(row0 + E_THREADS) - base);  // a tile lies inside one range: < 4 GiB
    uint32_t       off = 0, len = 0;
    if (i < n) {
      const uint64_t o = offset_of(i);
      off = (uint32_t)(o - base);
      len = (uint32_t)(offset_of(i + 1) - o);
    }
    if (base + total > out_cap) {
      if (tid == 0) dev_set_error(scratch, BK_ERR_NOMEM, tile);
      continue;
    }
    const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(out) + base) & 15);
    if (total + shift <= E_STAGE) {
      if (len) {
        MemSink ms{stage + shift + off, stage + shift + off, base + off};
        fn(i, ms);
      }
      __syncthreads();
      // stage[0] corresponds to the 16-byte aligned address out + base - shift
      char*          dst0 = out + base - shift;
      const uint32_t end = shift + total;
      const uint32_t nvec = (end + 15) / 16;
      for (uint32_t v = tid; v < nvec; v += E_THREADS) {
        const uint32_t b0 = v * 16;
        if (b0 >= shift && b0 + 16 <= end) {
          stg_stream16(dst0 + b0, *reinterpret_cast<const uint4*>(stage + b0));
        } else {
          for (uint32_t b = (b0 < shift ? shift : b0); b < b0 + 16 && b < end; b++) dst0[b] = stage[b];
        }
      }
      __syncthreads();  // the stage is rewritten by the next tile
    } else if (len) {
      MemSink ms{out + base + off, out + base + off, base + off};
      fn(i, ms);
    }
  }
}

#endif

inline int grid_for_kernel(const void* kernel, int threads, uint64_t ntiles) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
  if (per_sm < 1) per_sm = 1;
  uint64_t g = (uint64_t)kSMs * per_sm;
  return (int)(ntiles < g ? (ntiles ? ntiles : 1) : g);
}

int reset_scratch(bk_ctx* ctx);
int read_scratch(bk_ctx* ctx);

#ifdef __CUDACC__
// host helper: run an emitter over n rows into a freshly allocated device buffer of out_cap bytes.
// On return *d_out holds the text, *out_bytes / *out_rows its size (the stream has been synchronised).
template <class RowFn>
int run_emit(bk_ctx* ctx, const RowFn& fn, uint64_t n, uint64_t out_cap, char** d_out, uint64_t* out_bytes,
             uint64_t* out_rows) {
  (void)out_cap;  // historical: the result is allocated exactly, after the length pass
  *out_bytes = 0;
  *out_rows = 0;
  *d_out = nullptr;
  if (n == 0) {
    *d_out = reinterpret_cast<char*>(dmalloc(ctx, 16));
    return *d_out ? BK_OK : BK_ERR_NOMEM;
  }
  const uint32_t ntiles = (uint32_t)((n + E_THREADS - 1) / E_THREADS);
  // warp ranges of 2^range_shift rows (>= one tile): up to four waves of 64 warps per SM
  int range_shift = 8;
  while (((n + (1ull << range_shift) - 1) >> range_shift) > (uint64_t)kSMs * 256 && range_shift < 20) range_shift++;
  const uint32_t nwarps = (uint32_t)((n + (1ull << range_shift) - 1) >> range_shift);
  uint32_t*      local_off = dalloc<uint32_t>(ctx, n);
  uint64_t*      wtot = dalloc<uint64_t>(ctx, nwarps);
  uint64_t*      wbase = dalloc<uint64_t>(ctx, (size_t)nwarps + 1);
  if (!local_off || !wtot || !wbase) return BK_ERR_NOMEM;
  BK_TRY(reset_scratch(ctx));
  const uint64_t len_ctas = ((uint64_t)nwarps * 32 + E_THREADS - 1) / E_THREADS;
  prof_begin(ctx, "k_emit_len");
  k_emit_len<RowFn><<<grid_for_kernel((const void*)k_emit_len<RowFn>, E_THREADS, len_ctas), E_THREADS, 0, ctx->stream>>>(
      fn, n, 1u << range_shift, local_off, wtot, nwarps, ctx->d_scratch);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  k_scan_totals<SC_OUT_BYTES><<<1, 1024, 0, ctx->stream>>>(wtot, wbase, nwarps, ctx->d_scratch);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));  // the output size: allocate exactly
  const uint64_t total = ctx->h_scratch[SC_OUT_BYTES];
  if (!ctx->h_scratch[SC_ERR_CODE]) *d_out = reinterpret_cast<char*>(dmalloc(ctx, total + 16));
  if (ctx->h_scratch[SC_ERR_CODE] || !*d_out) {
    dfree(ctx, local_off);
    dfree(ctx, wtot);
    dfree(ctx, wbase);
    if (!ctx->h_scratch[SC_ERR_CODE]) return BK_ERR_NOMEM;
    return fail(ctx, (int)ctx->h_scratch[SC_ERR_CODE], "output row too long for the device writer (row %llu)",
                (unsigned long long)ctx->h_scratch[SC_ERR_ROW]);
  }
  prof_begin(ctx, "k_emit");
  k_emit<RowFn><<<grid_for_kernel((const void*)k_emit<RowFn>, E_THREADS, ntiles), E_THREADS, 0, ctx->stream>>>(
      fn, n, *d_out, total, range_shift, local_off, wbase, nwarps, ntiles, ctx->d_scratch);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));
  dfree(ctx, local_off);
  dfree(ctx, wtot);
  dfree(ctx, wbase);
  if (ctx->h_scratch[SC_ERR_CODE]) {
    int code = (int)ctx->h_scratch[SC_ERR_CODE];
    dfree(ctx, *d_out);
    *d_out = nullptr;
    return fail(ctx, code, code == BK_ERR_NOMEM ? "output writer ran past the length pass (tile %llu)"
                                                 : "value outside the exact device formatter (row block %llu)",
                (unsigned long long)ctx->h_scratch[SC_ERR_ROW]);
  }
  *out_bytes = ctx->h_scratch[SC_OUT_BYTES];
  *out_rows = ctx->h_scratch[SC_OUT_ROWS];
  return BK_OK;
}
#endif

}  // namespace bk
