// starch.cu -- Starch v2 archives as input.
//
// Replaces the reading side of interfaces/src/data/starch/unstarchHelpers.c (UNSTARCH_extractDataWithBzip2 :265-355,
// UNSTARCH_extractDataWithGzip :57-263, UNSTARCH_reverseTransformHeaderlessInput :1161-1238) behind
// allocate_iterator_starch_bed's archive detection (AllocateIterator_BED_starch.hpp:100-112; starchApi.hpp:1427).
//
// An archive is: magic ca5cade5 | one bzip2 or gzip stream per chromosome | JSON metadata | 127-byte footer (20 decimal
// digits = offset of the metadata, base64 SHA-1 of it, padding).  A decompressed stream is delta-coded text:
//     p<len>            the length of the rows that follow
//     <d>[\t<rest>]     a row: start = end of the previous row + d (d may be negative), end = start + len
// Split of the work: the host walks the container and inflates the streams (bzip2 / gzip are serial bit streams: one
// host thread per stream); the DEVICE undoes the delta coding for all chromosomes at once -- "end of row i" is a prefix
// sum of (d + len) over the rows of a chromosome and "len of row i" is the value of the last p-line before it (a
// running-maximum scan over p-line indices) -- and writes plain BED text with the two-pass emitter.  The text then goes
// through the ordinary reader like any other input.
#include <dlfcn.h>
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include "common.cuh"
#include "emit.cuh"
#include "parse.cuh"
#include "sort.cuh"

namespace bk {

// ---------------------------------------------------------------------------------------------------------
// generic u64 scans over warp ranges (count -> one-CTA scan -> apply), no look-back
// ---------------------------------------------------------------------------------------------------------
constexpr int SCN_RANGE = 2048;
struct OpAdd {
  __device__ static uint64_t f(uint64_t a, uint64_t b) { return a + b; }
};
struct OpMax {
  __device__ static uint64_t f(uint64_t a, uint64_t b) { return a > b ? a : b; }
};
template <class OP>
__device__ __forceinline__ uint64_t warp_incl_scan64(uint64_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v = OP::f(t, v);
  }
  return v;
}
template <class OP>
__global__ void __launch_bounds__(256) k_scan64_totals(const uint64_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ tot) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * SCN_RANGE, b = a + SCN_RANGE < n ? a + SCN_RANGE : n;
  if (a >= n) return;
  uint64_t acc = 0;
  for (uint64_t k = a + lane; k < b; k += 32) acc = OP::f(acc, in[k]);
  acc = warp_incl_scan64<OP>(acc, lane);
  if (lane == 31) tot[w] = acc;
}
template <class OP>
__global__ void __launch_bounds__(1024) k_scan64_bases(uint64_t* __restrict__ tot, uint64_t n) {  // in-place exclusive scan, one CTA
  __shared__ uint64_t part[1024];
  const uint32_t      tid = threadIdx.x;
  const uint64_t      per = (n + 1023) / 1024;
  const uint64_t      b = (uint64_t)tid * per < n ? (uint64_t)tid * per : n, e = b + per < n ? b + per : n;
  uint64_t            s = 0;
  for (uint64_t i = b; i < e; i++) s = OP::f(s, tot[i]);
  part[tid] = s;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    const uint64_t v = tid >= d ? part[tid - d] : 0;
    __syncthreads();
    part[tid] = OP::f(v, part[tid]);
    __syncthreads();
  }
  uint64_t run = tid ? part[tid - 1] : 0;
  for (uint64_t i = b; i < e; i++) {
    const uint64_t v = tot[i];
    tot[i] = run;
    run = OP::f(run, v);
  }
}
template <class OP>
__global__ void __launch_bounds__(256) k_scan64_apply(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t n,
                                                      const uint64_t* __restrict__ base) {  // inclusive
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * SCN_RANGE, b = a + SCN_RANGE < n ? a + SCN_RANGE : n;
  if (a >= n) return;
  uint64_t run = base[w];
  for (uint64_t k0 = a; k0 < b; k0 += 32) {
    const uint64_t k = k0 + lane;
    uint64_t       v = k < b ? in[k] : 0;
    v = OP::f(run, warp_incl_scan64<OP>(v, lane));
    if (k < b) out[k] = v;
    run = __shfl_sync(0xffffffffu, v, 31);
  }
}
template <class OP>
static int inclusive_scan64(bk_ctx* ctx, const uint64_t* in, uint64_t* out, uint64_t n) {
  if (n == 0) return BK_OK;
  const uint64_t nw = (n + SCN_RANGE - 1) / SCN_RANGE;
  uint64_t*      tot = dalloc<uint64_t>(ctx, nw);
  if (!tot) return BK_ERR_NOMEM;
  const unsigned grid = (unsigned)((nw * 32 + 255) / 256);
  k_scan64_totals<OP><<<grid, 256, 0, ctx->stream>>>(in, n, tot);
  BK_LAUNCHED(ctx);
  k_scan64_bases<OP><<<1, 1024, 0, ctx->stream>>>(tot, nw);
  BK_LAUNCHED(ctx);
  k_scan64_apply<OP><<<grid, 256, 0, ctx->stream>>>(in, out, n, tot);
  BK_LAUNCHED(ctx);
  dfree(ctx, tot);
  return BK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// the delta-coded text -> per-line values
// ---------------------------------------------------------------------------------------------------------
// NL count of every SCN_RANGE-byte range of the text (as u64, scanned by the kernels above)
__global__ void __launch_bounds__(256) k_st_count_nl(const char* __restrict__ t, uint64_t n, uint64_t* __restrict__ cnt) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * SCN_RANGE, b = a + SCN_RANGE < n ? a + SCN_RANGE : n;
  if (a >= n) return;
  uint32_t c = 0;
  for (uint64_t k = a + lane; k < b; k += 32) c += t[k] == '\n' ? 1u : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) cnt[w] = c;
}
// line_end[j] = offset of the j-th NL
__global__ void __launch_bounds__(256) k_st_line_ends(const char* __restrict__ t, uint64_t n, const uint64_t* __restrict__ incl,
                                                      uint64_t* __restrict__ line_end) {
  const int      lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const uint64_t a = w * SCN_RANGE, b = a + SCN_RANGE < n ? a + SCN_RANGE : n;
  if (a >= n) return;
  uint64_t run = w ? incl[w - 1] : 0;
  for (uint64_t k0 = a; k0 < b; k0 += 32) {
    const uint64_t k = k0 + lane;
    const bool     f = k < b && t[k] == '\n';
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (f) line_end[run + __popc(m & ((1u << lane) - 1u))] = k;
    run += __popc(m);
  }
}

struct StLines {
  const char*     text;
  const uint64_t* line_end;  // [L]
  uint64_t        L;
  uint64_t*       pmark;     // [L] p-line: index + 1, else 0            -> running max = governing p-line + 1
  uint64_t*       step;      // [L] row: d + len (filled after the max scan), p-line: 0   -> prefix sum = end of the row
  int64_t*        delta;     // [L]
  uint64_t*       plen;      // [L] p-line: its value
  uint32_t*       rest;      // [L] offset of the rest from the line start (0: none)
  uint64_t*       scratch;
};
__device__ __forceinline__ uint64_t st_line_start(const StLines& p, uint64_t i) { return i ? p.line_end[i - 1] + 1 : 0; }

__global__ void __launch_bounds__(256) k_st_parse(StLines p) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.L) return;
  const char* s = p.text + st_line_start(p, i);
  uint64_t    v = 0;
  uint32_t    k = 0;
  if (s[0] == 'p') {
    k = 1;
    while (is_digit((unsigned char)s[k])) v = v * 10 + (uint64_t)(s[k++] - '0');
    if (k == 1 || s[k] != '\n') dev_set_error(p.scratch, BK_ERR_PARSE, i);
    p.pmark[i] = i + 1;
    p.plen[i] = v;
    p.delta[i] = 0;
    p.rest[i] = 0;
    return;
  }
  const bool neg = s[0] == '-';
  k = neg ? 1 : 0;
  const uint32_t k0 = k;
  while (is_digit((unsigned char)s[k])) v = v * 10 + (uint64_t)(s[k++] - '0');
  if (k == k0 || (s[k] != '\n' && s[k] != '\t')) dev_set_error(p.scratch, BK_ERR_PARSE, i);
  p.pmark[i] = 0;
  p.plen[i] = 0;
  p.delta[i] = neg ? -(int64_t)v : (int64_t)v;
  p.rest[i] = s[k] == '\t' ? k + 1 : 0;
}
// after the running max over pmark (in place): step = d + governing length for rows
__global__ void __launch_bounds__(256) k_st_steps(StLines p, const uint64_t* __restrict__ gov) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.L) return;
  const uint64_t g = gov[i];
  if (g == i + 1) {  // a p-line
    p.step[i] = 0;
    return;
  }
  if (g == 0) {  // a row before any p-line: not a Starch stream
    dev_set_error(p.scratch, BK_ERR_PARSE, i);
    p.step[i] = 0;
    return;
  }
  p.step[i] = (uint64_t)(p.delta[i] + (int64_t)p.plen[g - 1]);
}

struct StRow {  // "%s\t%ld\t%ld[\t%s]\n", unstarchHelpers.c:1183, :1189, :1224
  StLines         p;
  const uint64_t* gov;         // governing p-line + 1
  const uint64_t* ends;        // inclusive prefix sum of step over ALL lines
  const uint64_t* chr_line0;   // [nchr+1] first line of every chromosome's stream
  const char*     names;       // [nchr][128]
  const uint32_t* name_len;
  int             nchr;
  template <class Sink>
  __device__ void operator()(uint64_t i, Sink& s) const {
    const uint64_t g = gov[i];
    if (g == i + 1 || g == 0) return;  // p-line
    int l = 0, h = nchr;
    while (h - l > 1) {
      const int mid = (l + h) >> 1;
      if (chr_line0[mid] <= i) l = mid; else h = mid;
    }
    const uint64_t c0 = chr_line0[l];
    const uint64_t end = ends[i] - (c0 ? ends[c0 - 1] : 0);  // lastEnd restarts at 0 with every chromosome
    const uint64_t start = end - p.plen[g - 1];
    s.puts_(names + (size_t)l * 128, (int)name_len[l]);
    s.put('\t');
    s.put_u64(start);
    s.put('\t');
    s.put_u64(end);
    const uint32_t r = p.rest[i];
    if (r) {
      const uint64_t ls = st_line_start(p, i);
      s.put('\t');
      s.copy(p.text + ls + r, p.line_end[i] - (ls + r));
    }
    s.put('\n');
  }
};

// the transformed text of all chromosomes (concatenated, every stream NL-terminated) -> BED text
static int untransform(bk_ctx* ctx, const char* d_text, uint64_t nbytes, const std::vector<uint64_t>& chr_off,
                       const std::vector<std::string>& chr_names, int on_device, bk_text* out) {
  if (nbytes == 0) return finish_text(ctx, nullptr, 0, 0, on_device, out);
  std::vector<void*> tmp;
  auto done = [&](int rc) {
    for (void* q : tmp) dfree(ctx, q);
    return rc;
  };
  const uint64_t nw = (nbytes + SCN_RANGE - 1) / SCN_RANGE;
  uint64_t*      cnt = dalloc<uint64_t>(ctx, nw);
  tmp.push_back(cnt);
  if (!cnt) return done(BK_ERR_NOMEM);
  const unsigned gridb = (unsigned)((nw * 32 + 255) / 256);
  prof_begin(ctx, "k_st_count_nl");
  k_st_count_nl<<<gridb, 256, 0, ctx->stream>>>(d_text, nbytes, cnt);
  prof_end(ctx);
  ctx->launches++;
  int rc = inclusive_scan64<OpAdd>(ctx, cnt, cnt, nw);
  if (rc != BK_OK) return done(rc);
  uint64_t L = 0;
  if (cudaMemcpyAsync(&L, cnt + nw - 1, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess)
    return done(BK_ERR_CUDA);
  if (L == 0) return done(finish_text(ctx, nullptr, 0, 0, on_device, out));
  StLines p{};
  p.text = d_text; p.L = L; p.scratch = ctx->d_scratch;
  uint64_t* line_end = dalloc<uint64_t>(ctx, L);
  p.pmark = dalloc<uint64_t>(ctx, L); p.step = dalloc<uint64_t>(ctx, L); p.delta = dalloc<int64_t>(ctx, L);
  p.plen = dalloc<uint64_t>(ctx, L); p.rest = dalloc<uint32_t>(ctx, L);
  tmp.insert(tmp.end(), {line_end, p.pmark, p.step, p.delta, p.plen, p.rest});
  if (!line_end || !p.pmark || !p.step || !p.delta || !p.plen || !p.rest) return done(BK_ERR_NOMEM);
  p.line_end = line_end;
  k_st_line_ends<<<gridb, 256, 0, ctx->stream>>>(d_text, nbytes, cnt, line_end);
  ctx->launches++;
  if ((rc = reset_scratch(ctx)) != BK_OK) return done(rc);
  const unsigned gridl = (unsigned)((L + 255) / 256);
  prof_begin(ctx, "k_st_parse");
  k_st_parse<<<gridl, 256, 0, ctx->stream>>>(p);
  prof_end(ctx);
  ctx->launches++;
  if ((rc = inclusive_scan64<OpMax>(ctx, p.pmark, p.pmark, L)) != BK_OK) return done(rc);
  k_st_steps<<<gridl, 256, 0, ctx->stream>>>(p, p.pmark);
  ctx->launches++;
  if ((rc = inclusive_scan64<OpAdd>(ctx, p.step, p.step, L)) != BK_OK) return done(rc);
  // first line of every chromosome's stream: the number of NLs before its byte offset (host: bisect the line ends)
  std::vector<uint64_t> ends_h;  // only the few offsets are needed: count NLs with the range counts + a partial range
  std::vector<uint64_t> line0(chr_off.size(), 0);
  {
    // line0[c] = #NL in [0, chr_off[c]) = first j with line_end[j] >= chr_off[c]: bisect on the device array from the host
    for (size_t c = 0; c < chr_off.size(); c++) {
      uint64_t lo = 0, hi = L;
      while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        uint64_t       v = 0;
        if (cudaMemcpyAsync(&v, line_end + mid, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
          return done(BK_ERR_CUDA);
        if (v < chr_off[c]) lo = mid + 1; else hi = mid;
      }
      line0[c] = lo;
    }
  }
  if ((rc = read_scratch(ctx)) != BK_OK) return done(rc);
  if (ctx->h_scratch[SC_ERR_CODE])
    return done(fail(ctx, BK_ERR_STARCH, "Starch stream line %llu is neither p<len> nor <delta>[<TAB>rest]", (unsigned long long)ctx->h_scratch[SC_ERR_ROW] + 1));
  std::vector<char>     nm(chr_names.size() * 128, 0);
  std::vector<uint32_t> nl(chr_names.size());
  for (size_t c = 0; c < chr_names.size(); c++) {
    if (chr_names[c].size() > 127) return done(fail(ctx, BK_ERR_STARCH, "chromosome name longer than 127 bytes in the archive metadata"));
    memcpy(&nm[c * 128], chr_names[c].data(), chr_names[c].size());
    nl[c] = (uint32_t)chr_names[c].size();
  }
  StRow fn{};
  fn.p = p; fn.gov = p.pmark; fn.ends = p.step; fn.nchr = (int)chr_names.size();
  uint64_t* d_l0 = dalloc<uint64_t>(ctx, line0.size());
  char*     d_nm = dalloc<char>(ctx, nm.size());
  uint32_t* d_nl = dalloc<uint32_t>(ctx, nl.size());
  tmp.insert(tmp.end(), {d_l0, d_nm, d_nl});
  if (!d_l0 || !d_nm || !d_nl) return done(BK_ERR_NOMEM);
  if (cudaMemcpyAsync(d_l0, line0.data(), line0.size() * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(d_nm, nm.data(), nm.size(), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(d_nl, nl.data(), nl.size() * 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return done(BK_ERR_CUDA);
  fn.chr_line0 = d_l0; fn.names = d_nm; fn.name_len = d_nl;
  char*    d_out = nullptr;
  uint64_t bytes = 0, rows = 0;
  rc = run_emit(ctx, fn, L, 0, &d_out, &bytes, &rows);  // syncs: the host vectors above are consumed
  if (rc != BK_OK) {
    dfree(ctx, d_out);
    return done(rc);
  }
  return done(finish_text(ctx, d_out, bytes, rows, on_device, out));
}

// ---------------------------------------------------------------------------------------------------------
// host: the container
// ---------------------------------------------------------------------------------------------------------
struct StStream {
  std::string chrom;
  uint64_t    offset, size;
};
// the few fields the reader needs from the jansson-written metadata (starchMetadataHelpers.c:485-757): a flat scan for
// keys in document order is enough -- "streams" is an array of objects whose "chromosome" precedes their "size"
static bool json_string_after(const std::string& js, size_t& pos, const char* key, std::string* val, bool quoted_number) {
  const std::string k = std::string("\"") + key + "\"";
  const size_t      at = js.find(k, pos);
  if (at == std::string::npos) return false;
  size_t q = js.find(':', at + k.size());
  if (q == std::string::npos) return false;
  q++;
  while (q < js.size() && (js[q] == ' ' || js[q] == '\n' || js[q] == '\t' || js[q] == '\r')) q++;
  val->clear();
  if (q < js.size() && js[q] == '"') {
    q++;
    while (q < js.size() && js[q] != '"') {
      if (js[q] == '\\' && q + 1 < js.size()) q++;
      val->push_back(js[q++]);
    }
    q++;
  } else {
    (void)quoted_number;
    while (q < js.size() && js[q] != ',' && js[q] != '}' && js[q] != '\n' && js[q] != ' ') val->push_back(js[q++]);
  }
  pos = q;
  return true;
}

typedef int (*bz_buff_fn)(char*, unsigned int*, char*, unsigned int, int, int);
struct BzStream {  // bz_stream of bzlib.h (1.0.x ABI)
  char*        next_in;
  unsigned int avail_in, total_in_lo32, total_in_hi32;
  char*        next_out;
  unsigned int avail_out, total_out_lo32, total_out_hi32;
  void*        state;
  void* (*bzalloc)(void*, int, int);
  void (*bzfree)(void*, void*);
  void* opaque;
};
typedef int (*bz_init_fn)(BzStream*, int, int);
typedef int (*bz_step_fn)(BzStream*);
struct BzLib {
  bz_init_fn init = nullptr;
  bz_step_fn step = nullptr, end = nullptr;
  bool       ok() const { return init && step && end; }
};
static BzLib load_bz() {
  static BzLib lib = [] {
    BzLib l;
    void* h = nullptr;
    for (const char* n : {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"})
      if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (h) {
      l.init = reinterpret_cast<bz_init_fn>(dlsym(h, "BZ2_bzDecompressInit"));
      l.step = reinterpret_cast<bz_step_fn>(dlsym(h, "BZ2_bzDecompress"));
      l.end = reinterpret_cast<bz_step_fn>(dlsym(h, "BZ2_bzDecompressEnd"));
    }
    return l;
  }();
  return lib;
}
static bool inflate_bz(const char* in, size_t n, std::string* out) {
  const BzLib lib = load_bz();
  if (!lib.ok()) return false;
  BzStream s{};
  if (lib.init(&s, 0, 0) != 0) return false;
  s.next_in = const_cast<char*>(in);
  s.avail_in = (unsigned)n;
  std::vector<char> buf(1u << 20);
  int               rc = 0;
  while (true) {
    s.next_out = buf.data();
    s.avail_out = (unsigned)buf.size();
    rc = lib.step(&s);
    out->append(buf.data(), buf.size() - s.avail_out);
    if (rc == 4 /* BZ_STREAM_END */) break;
    if (rc != 0 || (s.avail_in == 0 && s.avail_out != 0)) {
      lib.end(&s);
      return false;
    }
  }
  lib.end(&s);
  return true;
}
static bool inflate_gz(const char* in, size_t n, std::string* out) {
  z_stream s{};
  if (inflateInit2(&s, 15 + 32) != Z_OK) return false;  // zlib or gzip wrapper
  s.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in));
  s.avail_in = (uInt)n;
  std::vector<char> buf(1u << 20);
  int               rc = Z_OK;
  while (rc != Z_STREAM_END) {
    s.next_out = reinterpret_cast<Bytef*>(buf.data());
    s.avail_out = (uInt)buf.size();
    rc = inflate(&s, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END) {
      inflateEnd(&s);
      return false;
    }
    out->append(buf.data(), buf.size() - s.avail_out);
    if (rc == Z_OK && s.avail_in == 0 && s.avail_out != 0) {
      inflateEnd(&s);
      return false;
    }
  }
  inflateEnd(&s);
  return true;
}

// host stage: walk the container, inflate the wanted streams (in archive order).  err receives the message on failure.
static int starch_inflate(const char* archive, size_t nbytes, const char* chrom, std::vector<std::string>* parts,
                          std::vector<std::string>* names, std::string* err) {
  auto bad_archive = [&](const std::string& m) {
    *err = m;
    return BK_ERR_STARCH;
  };
  if (!bk_is_starch(archive, nbytes))
    return bad_archive("not a Starch v2 archive (v1 archives -- bare bzip2/gzip streams with a leading metadata block -- are not read)");
  constexpr size_t kFooter = 127;  // STARCH2_MD_FOOTER_LENGTH - 1 bytes on disk (starchMetadataHelpers.c:1113-1119, :1178)
  if (nbytes < 4 + kFooter) return bad_archive("truncated Starch archive");
  // footer: "%020llu%s" offset and base64 SHA-1, space padded, NL (starchHelpers.c STARCH2_writeStarchFooter)
  const char* foot = archive + nbytes - kFooter;
  uint64_t    md_off = 0;
  for (int k = 0; k < 20; k++) {
    if (foot[k] < '0' || foot[k] > '9') return bad_archive("Starch footer does not hold a metadata offset");
    md_off = md_off * 10 + (uint64_t)(foot[k] - '0');
  }
  if (md_off < 4 || md_off > nbytes - kFooter) return bad_archive("Starch metadata offset outside the archive");
  const std::string js(archive + md_off, nbytes - kFooter - md_off);
  size_t            pos = 0;
  std::string       v;
  int               comp = 0;  // kBzip2 = 0, kGzip = 1 (starchMetadataHelpers.h CompressionType)
  {
    size_t p0 = 0;
    if (json_string_after(js, p0, "compressionFormat", &v, false)) comp = atoi(v.c_str());
    p0 = 0;
    if (json_string_after(js, p0, "customUCSCHeaders", &v, false) && v.find("true") == 0)
      return bad_archive("Starch archive made with --header (UCSC header lines inside the streams) is not read");
  }
  std::vector<StStream> streams;
  uint64_t              at = 4;
  pos = js.find("\"streams\"");
  if (pos == std::string::npos) return bad_archive("Starch metadata has no stream list");
  while (true) {
    StStream st;
    if (!json_string_after(js, pos, "chromosome", &st.chrom, false)) break;
    if (!json_string_after(js, pos, "size", &v, true)) return bad_archive("Starch metadata: stream without size");
    st.size = strtoull(v.c_str(), nullptr, 10);
    st.offset = at;
    if (st.size > md_off - at) return bad_archive("Starch metadata: streams overrun the archive");  // (no wrap: at <= md_off)
    at += st.size;
    streams.push_back(st);
  }
  const bool all = !chrom || !*chrom || strcmp(chrom, "all") == 0;
  std::vector<StStream> want;
  for (auto& s : streams)
    if (all || s.chrom == chrom) want.push_back(s);
  // inflate: one host thread per stream (bounded), in archive order = chromosome order
  parts->assign(want.size(), std::string());
  std::atomic<size_t> next{0};
  std::atomic<int>    bad{-1};
  const unsigned      nthreads = std::max(1u, std::min<unsigned>((unsigned)want.size(), std::min(32u, std::thread::hardware_concurrency())));
  auto work = [&]() {
    while (true) {
      const size_t k = next.fetch_add(1);
      if (k >= want.size()) break;
      std::string& dst = (*parts)[k];
      const bool   ok = comp == 1 ? inflate_gz(archive + want[k].offset, want[k].size, &dst) : inflate_bz(archive + want[k].offset, want[k].size, &dst);
      if (!ok) bad = (int)k;
      else if (!dst.empty() && dst.back() != '\n') dst.push_back('\n');
    }
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  if (bad >= 0)
    return bad_archive(std::string("could not inflate the ") + (comp == 1 ? "gzip" : "bzip2") + " stream of chromosome '" + want[bad].chrom + "'" +
                       ((comp != 1 && !load_bz().ok()) ? " (libbz2.so.1.0 not found)" : ""));
  names->clear();
  for (auto& w : want) names->push_back(w.chrom);
  return BK_OK;
}

}  // namespace bk

using namespace bk;

extern "C" int bk_is_starch(const char* bytes, size_t n) {
  const unsigned char* h = reinterpret_cast<const unsigned char*>(bytes);
  return n >= 4 && h[0] == 0xca && h[1] == 0x5c && h[2] == 0xad && h[3] == 0xe5;
}

extern "C" int bk_starch_inflate_host(const char* archive, size_t nbytes, const char* chrom, char** text, size_t* len) {
  if (!archive || !text || !len) return BK_ERR_ARG;
  *text = nullptr;
  *len = 0;
  std::vector<std::string> parts, names;
  std::string              err;
  const int                rc = starch_inflate(archive, nbytes, chrom, &parts, &names, &err);
  if (rc != BK_OK) return rc;
  std::string all;
  for (size_t k = 0; k < parts.size(); k++) all += ">" + names[k] + "\n" + parts[k];
  *text = static_cast<char*>(malloc(all.size() + 1));
  if (!*text) return BK_ERR_NOMEM;
  memcpy(*text, all.data(), all.size());
  *len = all.size();
  return BK_OK;
}
extern "C" void bk_host_free(void* p) { free(p); }

extern "C" int bk_unstarch(bk_ctx* ctx, const char* archive, size_t nbytes, const char* chrom, int out_on_device, bk_text* out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !out || !archive) return BK_ERR_ARG;
  ctx->last_error.clear();
  memset(out, 0, sizeof(*out));
  std::vector<std::string> parts, want_names;
  {
    std::string err;
    const int   rc = starch_inflate(archive, nbytes, chrom, &parts, &want_names, &err);
    if (rc != BK_OK) return fail(ctx, rc, "%s", err.c_str());
  }
  struct W {
    std::string chrom;
  };
  std::vector<W> want;
  for (auto& n : want_names) want.push_back(W{n});
  std::vector<uint64_t>    off;
  std::vector<std::string> names;
  uint64_t                 total = 0;
  for (size_t k = 0; k < want.size(); k++) {
    if (parts[k].empty()) continue;
    off.push_back(total);
    names.push_back(want[k].chrom);
    total += parts[k].size();
  }
  if (total == 0) return finish_text(ctx, nullptr, 0, 0, out_on_device, out);
  char* d = reinterpret_cast<char*>(dmalloc(ctx, total + 64));
  if (!d) return BK_ERR_NOMEM;
  {
    size_t k2 = 0;
    for (size_t k = 0; k < want.size(); k++) {
      if (parts[k].empty()) continue;
      if (cudaMemcpyAsync(d + off[k2], parts[k].data(), parts[k].size(), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        dfree(ctx, d);
        return BK_ERR_CUDA;
      }
      k2++;
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      dfree(ctx, d);
      return BK_ERR_CUDA;
    }
  }
  const int rc = untransform(ctx, d, total, off, names, out_on_device, out);
  dfree(ctx, d);
  return rc;
}
