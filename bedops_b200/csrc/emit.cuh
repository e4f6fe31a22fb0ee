// emit.cuh -- the BED writer skeleton shared by every tool: one output row per thread, single pass.
//
//   length pass (CountSink) -> block scan -> decoupled look-back over tiles (byte offsets) ->
//   write pass (MemSink) into a shared-memory stage laid out with the same 16-byte phase as the destination ->
//   16-byte coalesced streaming stores to HBM.
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

template <class RowFn>
__global__ void __launch_bounds__(E_THREADS, 4) k_emit(RowFn fn, uint64_t n, char* __restrict__ out, uint64_t out_cap,
                                                    uint64_t* tile_state, uint32_t ntiles, uint64_t* scratch) {
  __shared__ __align__(16) char stage[E_STAGE + 16];
  __shared__ uint32_t           scan_sm[34];
  __shared__ uint32_t           ticket_sm;
  __shared__ uint64_t           base_sm;
  const int tid = threadIdx.x;
  while (true) {
    const uint32_t tile = next_ticket(scratch, &ticket_sm);
    if (tile >= ntiles) break;
    const uint64_t i = (uint64_t)tile * E_THREADS + tid;
    uint64_t       len64 = 0;
    if (i < n) {
      CountSink cs;
      fn(i, cs);
      len64 = cs.n;
    }
    // rows longer than 4 GiB are not representable in the block scan; rest fields are <= 1 MiB per row
    uint32_t len = (uint32_t)len64, total;
    uint32_t off = block_excl_scan(len, scan_sm, &total);
    uint32_t nz = __syncthreads_count(len != 0);
    if (tid < 32) {
      uint64_t b = lookback_sum(tile_state, tile, total);
      if (tid == 0) {
        base_sm = b;
        if (nz) atomicAdd(reinterpret_cast<unsigned long long*>(&scratch[SC_OUT_ROWS]), (unsigned long long)nz);
        if (tile == ntiles - 1) scratch[SC_OUT_BYTES] = b + total;
      }
    }
    __syncthreads();
    const uint64_t base = base_sm;
    if (base + total > out_cap) {
      if (tid == 0) dev_set_error(scratch, BK_ERR_NOMEM, tile);
      continue;
    }
    const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(out) + base) & 15);
    if (total + shift <= E_STAGE) {
      if (len) {
        MemSink ms{stage + shift + off};
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
    } else if (len) {
      MemSink ms{out + base + off};
      fn(i, ms);
    }
    __syncthreads();
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
  *out_bytes = 0;
  *out_rows = 0;
  *d_out = reinterpret_cast<char*>(dmalloc(ctx, out_cap + 16));
  if (!*d_out) return BK_ERR_NOMEM;
  if (n == 0) return BK_OK;
  const uint32_t ntiles = (uint32_t)((n + E_THREADS - 1) / E_THREADS);
  uint64_t*      state = dalloc<uint64_t>(ctx, ntiles);
  if (!state) return BK_ERR_NOMEM;
  BK_CUDA(ctx, cudaMemsetAsync(state, 0, (size_t)ntiles * 8, ctx->stream));
  BK_TRY(reset_scratch(ctx));
  prof_begin(ctx, "k_emit");
  k_emit<RowFn><<<grid_for_kernel((const void*)k_emit<RowFn>, E_THREADS, ntiles), E_THREADS, 0, ctx->stream>>>(
      fn, n, *d_out, out_cap, state, ntiles, ctx->d_scratch);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));
  dfree(ctx, state);
  if (ctx->h_scratch[SC_ERR_CODE]) {
    int code = (int)ctx->h_scratch[SC_ERR_CODE];
    return fail(ctx, code, code == BK_ERR_NOMEM ? "output exceeds the precomputed bound (tile %llu)"
                                                 : "value outside the exact device formatter (row block %llu)",
                (unsigned long long)ctx->h_scratch[SC_ERR_ROW]);
  }
  *out_bytes = ctx->h_scratch[SC_OUT_BYTES];
  *out_rows = ctx->h_scratch[SC_OUT_ROWS];
  return BK_OK;
}
#endif

}  // namespace bk
