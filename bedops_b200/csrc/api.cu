// api.cu -- context, memory and the load/free half of the C ABI (include/bedkit.h).
#include <chrono>
#include <thread>
#include <stdarg.h>
#include "common.cuh"
#include "emit.cuh"
#include "parse.cuh"

namespace bk {

int fail(bk_ctx* ctx, int code, const char* fmt, ...) {
  char    buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->last_error = buf;
  return code;
}

int cuda_fail(bk_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  return fail(ctx, e == cudaErrorMemoryAllocation ? BK_ERR_NOMEM : BK_ERR_CUDA, "CUDA error %d (%s) in %s at %s:%d", (int)e,
              cudaGetErrorString(e), what, base ? base + 1 : file, line);
}

// ---- device memory: a per-context caching allocator -------------------------------------------------------------
// Every call allocates the same handful of (large) blocks again: text, columns, result.  The driver's stream-ordered
// pool re-maps multi-GB blocks on most requests (measured: 10 ms .. 3 s per cudaMallocAsync of 3-16 GB in steady
// state), so freed blocks are kept here, keyed by a size class (8 classes per octave, <= 12.5 % slack), and handed out
// again without a driver call.  All work of a context is ordered on one stream, so immediate reuse is safe.
// cudaMalloc is called only for a class with no cached block; when it fails the cache is released and it is retried.
static size_t size_class(size_t b) {
  if (b < 512) return 512;
  int top = 63 - __builtin_clzll((unsigned long long)b);
  const size_t unit = (size_t)1 << (top - 3);
  return (b + unit - 1) & ~(unit - 1);
}

void release_cached(bk_ctx* ctx) {
  if (ctx->dev_free.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->dev_free) cudaFree(kv.second);
  ctx->dev_free.clear();
  ctx->dev_cached_bytes = 0;
}

void* dmalloc(bk_ctx* ctx, size_t bytes) {
  const size_t sz = size_class(bytes);
  // the smallest idle block of this class or, failing that, of at most twice the size: successive chromosome groups of a
  // pipelined call ask for similar, not equal, sizes, and a driver allocation costs 0.3-4 ms each
  auto it = ctx->dev_free.lower_bound(sz);
  if (it != ctx->dev_free.end() && it->first <= 2 * sz) {
    void*        p = it->second;
    const size_t have = it->first;
    ctx->dev_free.erase(it);
    ctx->dev_cached_bytes -= have;
    ctx->dev_live[p] = have;
    return p;
  }
  static const bool trace = getenv("BEDKIT_TRACE") != nullptr;  // host-side cost of new blocks, to stderr
  const auto        t0 = std::chrono::steady_clock::now();
  void*             p = nullptr;
  cudaError_t       e = cudaMalloc(&p, sz);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    release_cached(ctx);
    e = cudaMalloc(&p, sz);
  }
  if (trace) {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[bedkit] cudaMalloc(%zu bytes, class %zu) took %.2f ms; cached %zu bytes\n", bytes, sz, ms,
            ctx->dev_cached_bytes);
  }
  if (e != cudaSuccess) {
    cuda_fail(ctx, e, "cudaMalloc", __FILE__, __LINE__);
    cudaGetLastError();
    return nullptr;
  }
  ctx->dev_live[p] = sz;
  return p;
}

void dfree(bk_ctx* ctx, void* p) {
  if (!p) return;
  auto it = ctx->dev_live.find(p);
  if (it == ctx->dev_live.end()) return;  // not ours (borrowed text)
  ctx->dev_free.emplace(it->second, p);
  ctx->dev_cached_bytes += it->second;
  ctx->dev_live.erase(it);
  if (ctx->dev_cached_bytes > ctx->dev_cache_limit) release_cached(ctx);  // do not hoard: other tenants share the HBM
}

char* pinned_get(bk_ctx* ctx, size_t bytes) {
  if (bytes == 0) bytes = 16;
  int best = -1;
  for (size_t i = 0; i < ctx->pinned.size(); i++) {
    auto& b = ctx->pinned[i];
    if (!b.busy && b.cap >= bytes && (best < 0 || b.cap < ctx->pinned[best].cap)) best = (int)i;
  }
  if (best >= 0) {
    ctx->pinned[best].busy = true;
    return ctx->pinned[best].ptr;
  }
  // drop idle buffers that are too small, then allocate with headroom
  for (size_t i = 0; i < ctx->pinned.size();) {
    if (!ctx->pinned[i].busy) {
      cudaFreeHost(ctx->pinned[i].ptr);
      ctx->pinned.erase(ctx->pinned.begin() + i);
    } else {
      i++;
    }
  }
  size_t cap = bytes + bytes / 8 + 4096;
  char*  p = nullptr;
  if (cudaHostAlloc(reinterpret_cast<void**>(&p), cap, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    fail(ctx, BK_ERR_NOMEM, "cudaHostAlloc(%zu) failed", cap);
    return nullptr;
  }
  ctx->pinned.push_back({p, cap, true});
  return p;
}

void pinned_put(bk_ctx* ctx, char* p) {
  for (auto& b : ctx->pinned)
    if (b.ptr == p) b.busy = false;
}

int upload(bk_ctx* ctx, char* d_dst, const char* src, size_t n, cudaStream_t st) {
  if (n == 0) return BK_OK;
  constexpr size_t kChunk = 32u << 20;
  cudaPointerAttributes at;
  bool direct = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type != cudaMemoryTypeUnregistered;
  cudaGetLastError();
  if (direct || n < (4u << 20)) {
    BK_CUDA(ctx, cudaMemcpyAsync(d_dst, src, n, cudaMemcpyDefault, st));
    return BK_OK;
  }
  for (int k = 0; k < 3; k++)
    if (!ctx->stage[k]) {
      BK_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->stage[k]), kChunk, cudaHostAllocDefault));
      BK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming));
    }
  unsigned nt = std::thread::hardware_concurrency() / 2;
  nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
  size_t off = 0;
  for (int i = 0; off < n; i++) {
    const int    k = i % 3;
    const size_t len = n - off < kChunk ? n - off : kChunk;
    if (i >= 3) BK_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[k]));  // the DMA that last read this buffer is done
    {  // fill the staging buffer with nt threads (page faults and memcpy of the mapped file in parallel)
      std::vector<std::thread> th;
      const size_t per = ((len + nt - 1) / nt + 4095) & ~(size_t)4095;
      for (unsigned t = 1; t < nt; t++) {
        const size_t a = (size_t)t * per;
        if (a >= len) break;
        const size_t m = len - a < per ? len - a : per;
        th.emplace_back([=]() { memcpy(ctx->stage[k] + a, src + off + a, m); });
      }
      memcpy(ctx->stage[k], src + off, len < per ? len : per);
      for (auto& t : th) t.join();
    }
    BK_CUDA(ctx, cudaMemcpyAsync(d_dst + off, ctx->stage[k], len, cudaMemcpyHostToDevice, st));
    BK_CUDA(ctx, cudaEventRecord(ctx->stage_ev[k], st));
    off += len;
  }
  // the staging ring is reused by the next upload of this ctx: wait for the last chunks here
  for (int k = 0; k < 3; k++) BK_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[k]));
  return BK_OK;
}

int upload_params(bk_ctx* ctx, void* d_dst, const void* src, size_t n) {
  if (n == 0) return BK_OK;
  const size_t need = (n + 63) & ~(size_t)63;
  if (!ctx->h_params || need > kParamRing) {  // no ring (allocation failed) or a table larger than it: the plain way
    BK_CUDA(ctx, cudaMemcpyAsync(d_dst, src, n, cudaMemcpyHostToDevice, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BK_OK;
  }
  if (ctx->h_params_used + need > kParamRing) {  // earlier tables may still be in flight: drain, then start over
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->h_params_used = 0;
  }
  char* h = ctx->h_params + ctx->h_params_used;
  memcpy(h, src, n);
  ctx->h_params_used += need;
  BK_CUDA(ctx, cudaMemcpyAsync(d_dst, h, n, cudaMemcpyHostToDevice, ctx->stream));
  return BK_OK;
}

static cudaEvent_t prof_event(bk_ctx* ctx) {
  if (!ctx->prof_free.empty()) {
    cudaEvent_t e = ctx->prof_free.back();
    ctx->prof_free.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(bk_ctx* ctx, const char* name) {
  if (!ctx->prof_on) return;
  bk_ctx::ProfRec r{name, prof_event(ctx), prof_event(ctx)};
  cudaEventRecord(r.a, ctx->stream);
  ctx->prof.push_back(r);
}
void prof_end(bk_ctx* ctx) {
  if (!ctx->prof_on || ctx->prof.empty()) return;
  cudaEventRecord(ctx->prof.back().b, ctx->stream);
}

// hand a device result buffer to the caller: either as-is (on_device) or copied into pinned host memory
int finish_text(bk_ctx* ctx, char* d_out, uint64_t bytes, uint64_t rows, int on_device, bk_text* out) {
  out->len = bytes;
  out->rows = rows;
  out->on_device = on_device ? 1 : 0;
  if (on_device) {
    out->ptr = d_out;
    return BK_OK;
  }
  out->ptr = pinned_get(ctx, bytes);
  if (!out->ptr) {
    dfree(ctx, d_out);
    return BK_ERR_NOMEM;
  }
  if (bytes) BK_CUDA(ctx, cudaMemcpyAsync(out->ptr, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  dfree(ctx, d_out);
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

}  // namespace bk

using namespace bk;

extern "C" int bk_abi_version(void) { return BEDKIT_ABI_VERSION; }

extern "C" const char* bk_strerror(int code) {
  switch (code) {
    case BK_OK: return "ok";
    case BK_ERR_CUDA: return "CUDA failure or no usable sm_100 device (this library has no CPU fallback)";
    case BK_ERR_NOMEM: return "out of memory";
    case BK_ERR_ARG: return "bad argument";
    case BK_ERR_PARSE: return "BED parse error";
    case BK_ERR_COORD_RANGE: return "coordinate outside the 32-bit device layout";
    case BK_ERR_UNSUPPORTED: return "option outside the device hot path";
    case BK_ERR_STARCH: return "Starch archive where plain BED text is expected (see bk_unstarch), or an archive that is not read (v1, --header, damaged)";
    case BK_ERR_UNSORTED: return "input is not sorted per sort-bed";
    case BK_ERR_CHECK: return "error-check (--ec) failure";
    case BK_ERR_NAN_ELEMENT: return "Unable to process a 'NAN' with PrintAllScorePrecision.";
  }
  return "unknown error";
}

extern "C" const char* bk_last_error(const bk_ctx* ctx) {
  bk::DeviceGuard device_guard(ctx); return ctx ? ctx->last_error.c_str() : ""; }
extern "C" uint64_t    bk_launch_count(const bk_ctx* ctx) {
  bk::DeviceGuard device_guard(ctx); return ctx ? ctx->launches : 0; }

extern "C" int bk_init(bk_ctx** out, int device) {
  if (!out) return BK_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return BK_ERR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BK_ERR_CUDA;
  if (prop.major != 10) return BK_ERR_CUDA;  // kernels are built for sm_100a only
  if (cudaSetDevice(device) != cudaSuccess) return BK_ERR_CUDA;
  bk_ctx* ctx = new bk_ctx();
  ctx->device = device;
  ctx->sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : kSMs;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return BK_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) ctx->dev_cache_limit = total_b / 4;
  if (cudaMalloc(reinterpret_cast<void**>(&ctx->d_scratch), SC_N * sizeof(uint64_t)) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_scratch), SC_N * sizeof(uint64_t) + kHostScratchExtra, cudaHostAllocDefault) != cudaSuccess) {
    bk_destroy(ctx);
    return BK_ERR_NOMEM;
  }
  if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_params), kParamRing, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    ctx->h_params = nullptr;  // upload_params falls back to synchronous copies
  }
  *out = ctx;
  return BK_OK;
}

extern "C" void bk_destroy(bk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  release_cached(ctx);
  for (auto& kv : ctx->dev_live) cudaFree(kv.first);  // blocks of handles the caller never freed
  for (auto& b : ctx->pinned) cudaFreeHost(b.ptr);
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : ctx->prof_free) cudaEventDestroy(e);
  for (int k = 0; k < 3; k++) {
    if (ctx->stage[k]) cudaFreeHost(ctx->stage[k]);
    if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
  }
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
  if (ctx->h_params) cudaFreeHost(ctx->h_params);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int bk_set_stream(bk_ctx* ctx, void* cuda_stream) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx) return BK_ERR_ARG;
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return BK_OK;
}

extern "C" int bk_release_cached(bk_ctx* ctx) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx) return BK_ERR_ARG;
  release_cached(ctx);
  return BK_OK;
}

extern "C" int bk_profile(bk_ctx* ctx, int enable) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx) return BK_ERR_ARG;
  cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->prof) {
    ctx->prof_free.push_back(r.a);
    ctx->prof_free.push_back(r.b);
  }
  ctx->prof.clear();
  ctx->prof_on = enable != 0;
  return BK_OK;
}

extern "C" int bk_profile_query(bk_ctx* ctx, const char* kernel, double* total_ms, uint64_t* launches) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !kernel) return BK_ERR_ARG;
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double   ms = 0;
  uint64_t n = 0;
  for (auto& r : ctx->prof) {
    if (strcmp(r.name, kernel) != 0) continue;
    float t = 0;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms += t;
      n++;
    }
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  return BK_OK;
}

extern "C" int bk_copy(bk_ctx* ctx, void* dst, const void* src, size_t nbytes) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx) return BK_ERR_ARG;
  if (nbytes) BK_CUDA(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDefault, ctx->stream));
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

extern "C" int bk_sync(bk_ctx* ctx) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx) return BK_ERR_ARG;
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

static int load_common(bk_ctx* ctx, bk_bed* bed, size_t nbytes, int min_fields, unsigned cols, bk_bed** out) {
  if (min_fields < 3 || min_fields > 5) {
    delete bed;
    return fail(ctx, BK_ERR_ARG, "min_fields must be 3, 4 or 5");
  }
  if ((cols & BK_COL_SCORE) && min_fields < 5) cols &= ~BK_COL_SCORE;
  if ((cols & BK_COL_ID) && min_fields < 4 && !(cols & BK_LOAD_SORTBED)) cols &= ~BK_COL_ID;
  if ((cols & BK_LOAD_SORTBED) && (min_fields != 3 || !(cols & BK_COL_LINE))) {
    delete bed;
    return fail(ctx, BK_ERR_ARG, "BK_LOAD_SORTBED needs min_fields == 3 and BK_COL_LINE");
  }
  if (cols & BK_COL_ID) cols |= BK_COL_LINE;
  bed->min_fields = min_fields;
  bed->cols = cols;
  int rc = parse_bed(ctx, bed, nbytes);
  if (rc != BK_OK) {
    bk_free_bed(ctx, bed);
    return rc;
  }
  *out = bed;
  return BK_OK;
}

static bool looks_like_starch(const unsigned char* h, size_t n) {
  // starch v2 magic ca5cade5; v1 archives are bzip2 ("BZh") or gzip (1f 8b) streams
  // (interfaces/general-headers/data/starch/starchApi.hpp:645-700)
  if (n >= 4 && h[0] == 0xca && h[1] == 0x5c && h[2] == 0xad && h[3] == 0xe5) return true;
  if (n >= 3 && h[0] == 'B' && h[1] == 'Z' && h[2] == 'h') return true;
  if (n >= 2 && h[0] == 0x1f && h[1] == 0x8b) return true;
  return false;
}

extern "C" int bk_load_bed(bk_ctx* ctx, const char* host_text, size_t nbytes, int min_fields, unsigned cols, bk_bed** out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !out || (!host_text && nbytes)) return BK_ERR_ARG;
  ctx->last_error.clear();
  *out = nullptr;
  if (looks_like_starch(reinterpret_cast<const unsigned char*>(host_text), nbytes))
    return fail(ctx, BK_ERR_STARCH, "input is a Starch/compressed archive: expand it with bk_unstarch first");
  bk_bed* bed = new bk_bed();
  char*   d = reinterpret_cast<char*>(dmalloc(ctx, nbytes + 64));
  if (!d) {
    delete bed;
    return BK_ERR_NOMEM;
  }
  bed->d_text = d;
  bed->owns_text = true;
  {
    const int rc = upload(ctx, d, host_text, nbytes, ctx->stream);
    if (rc != BK_OK) {
      bk_free_bed(ctx, bed);
      return rc;
    }
  }
  return load_common(ctx, bed, nbytes, min_fields, cols, out);
}

extern "C" int bk_load_bed_device(bk_ctx* ctx, const char* dev_text, size_t nbytes, int min_fields, unsigned cols, bk_bed** out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !out || (!dev_text && nbytes)) return BK_ERR_ARG;
  ctx->last_error.clear();
  *out = nullptr;
  if (reinterpret_cast<uintptr_t>(dev_text) & 15) return fail(ctx, BK_ERR_ARG, "device text must be 16-byte aligned");
  bk_bed* bed = new bk_bed();
  bed->d_text = dev_text;
  bed->owns_text = false;
  return load_common(ctx, bed, nbytes, min_fields, cols, out);
}

extern "C" void bk_free_bed(bk_ctx* ctx, bk_bed* bed) {
  bk::DeviceGuard device_guard(ctx);
  if (!bed || !ctx) return;
  if (bed->owns_text) dfree(ctx, const_cast<char*>(bed->d_text));
  dfree(ctx, bed->start);
  dfree(ctx, bed->end);
  dfree(ctx, bed->score);
  dfree(ctx, bed->line_off);
  dfree(ctx, bed->idspan);
  dfree(ctx, bed->pmax_end);
  dfree(ctx, bed->bmax_end);
  delete bed;
}

extern "C" uint64_t    bk_bed_rows(const bk_bed* bed) { return bed ? bed->nrows : 0; }
extern "C" int         bk_bed_nchrom(const bk_bed* bed) { return bed ? (int)bed->runs.size() : 0; }
extern "C" const char* bk_bed_chrom_name(const bk_bed* bed, int i) {
  return (bed && i >= 0 && i < (int)bed->runs.size()) ? bed->runs[i].name.c_str() : "";
}
extern "C" uint64_t bk_bed_chrom_rows(const bk_bed* bed, int i) {
  return (bed && i >= 0 && i < (int)bed->runs.size()) ? bed->runs[i].row_end - bed->runs[i].row_begin : 0;
}

extern "C" int bk_bed_copy_columns(bk_ctx* ctx, const bk_bed* bed, uint32_t* start, uint32_t* end, double* score,
                                   uint64_t* line_off) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !bed) return BK_ERR_ARG;
  const uint64_t n = bed->nrows;
  if (n == 0) return BK_OK;
  if (start) BK_CUDA(ctx, cudaMemcpyAsync(start, bed->start, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (end) BK_CUDA(ctx, cudaMemcpyAsync(end, bed->end, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (score) {
    if (!bed->score) return fail(ctx, BK_ERR_ARG, "no score column was parsed");
    BK_CUDA(ctx, cudaMemcpyAsync(score, bed->score, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (line_off) {
    if (!bed->line_off) return fail(ctx, BK_ERR_ARG, "no line-offset column was parsed");
    BK_CUDA(ctx, cudaMemcpyAsync(line_off, bed->line_off, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (line_off)
    for (uint64_t i = 0; i < n; i++) line_off[i] &= kLineOffMask;  // the high 16 bits carry the echo fast-path length
  return BK_OK;
}

extern "C" void bk_free_text(bk_ctx* ctx, bk_text* text) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !text || !text->ptr) return;
  if (text->on_device) dfree(ctx, text->ptr); else pinned_put(ctx, text->ptr);
  text->ptr = nullptr;
  text->len = 0;
}

// ---- synthetic BED writer (bench / tests) ---------------------------------------------------------------
namespace bk {
struct FormatBedRow {
  const uint32_t* start;
  const uint32_t* end;
  const uint32_t* score;
  int64_t         id_base;
  char            chrom[128];
  int             chrom_len;
  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& s) const {
    s.puts_(chrom, chrom_len);
    s.put('\t');
    s.put_u32(start[i]);
    s.put('\t');
    s.put_u32(end[i]);
    if (id_base >= 0) {
      s.puts_("\tid", 3);
      s.put_u64((uint64_t)id_base + i);
      s.put('\t');
      s.put_u32(score[i]);
    }
    s.put('\n');
  }
};
}  // namespace bk

extern "C" int bk_format_bed_device(bk_ctx* ctx, const char* chrom, const uint32_t* d_start, const uint32_t* d_end,
                                    const uint32_t* d_score, uint64_t n, int64_t id_base, bk_text* out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !chrom || !out || strlen(chrom) > 127) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  FormatBedRow fn{};
  fn.start = d_start; fn.end = d_end; fn.score = d_score; fn.id_base = id_base;
  fn.chrom_len = (int)strlen(chrom);
  memcpy(fn.chrom, chrom, fn.chrom_len);
  uint64_t cap = n * (fn.chrom_len + 1 + 10 + 1 + 10 + 3 + 20 + 1 + 10 + 1) + 64;
  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  int      rc = run_emit(ctx, fn, n, cap, &d_out, &bytes, &rows);
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return rc;
  }
  return finish_text(ctx, d_out, bytes, rows, 1, out);
}
