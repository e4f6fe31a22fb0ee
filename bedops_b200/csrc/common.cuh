// common.cuh -- context, error plumbing and small device primitives shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/bedkit.h"

namespace bk {

constexpr int kSMs = 148;
constexpr uint64_t kLineOffMask = (1ull << 48) - 1;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct ChromRun {
  std::string name;
  uint64_t    row_begin;
  uint64_t    row_end;
};

}  // namespace bk

struct bk_ctx {
  int          device = 0;
  int          sms = 148;  // multiProcessorCount of the device (148 on B200)
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string  last_error;
  uint64_t     launches = 0;
  // small device scratch reused by every call: error record, counters
  uint64_t* d_scratch = nullptr;  // 64 x u64
  uint64_t* h_scratch = nullptr;  // pinned mirror
  // pinned ring for the small host tables a call hands to its kernels (chromosome run tables ...): a copy out of it is truly
  // asynchronous (a copy from pageable memory first waits for the stream) and the table need not outlive the call
  char*  h_params = nullptr;
  size_t h_params_used = 0;
  // pinned host staging pool (result text / input upload)
  struct Pinned {
    char*  ptr;
    size_t cap;
    bool   busy;
  };
  std::vector<Pinned> pinned;
  // staging ring for uploads from pageable host memory (api.cu: upload)
  char*       stage[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[3] = {nullptr, nullptr, nullptr};
  // device block cache (api.cu): freed blocks by size class, live blocks by address
  std::multimap<size_t, void*>      dev_free;
  std::unordered_map<void*, size_t> dev_live;
  size_t                            dev_cached_bytes = 0;
  size_t                            dev_cache_limit = (size_t)32 << 30;  // idle bytes kept at most (bk_init: a quarter of the HBM)
  // optional per-kernel timing (bk_profile): CUDA event pairs recorded on the launching stream
  struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
  };
  bool                 prof_on = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_free;
};

struct bk_bed {
  const char* d_text = nullptr;  // device text (owned iff owns_text)
  uint64_t    nbytes = 0;        // effective length: up to and including the last '\n'
  bool        owns_text = false;
  int         min_fields = 3;
  unsigned    cols = 0;
  uint64_t    nrows = 0;
  uint32_t*   start = nullptr;
  uint32_t*   end = nullptr;
  double*     score = nullptr;
  uint64_t*   line_off = nullptr;  // [nrows+1]; low 48 bits = offset of the row's chromosome token, high 16 = line length
                                   // (bytes up to the NL) when echoing the row is a verbatim copy, else 0xFFFF
  uint32_t*   idspan = nullptr;    // (rel_off << 16) | len  relative to line_off
  uint32_t*   pmax_end = nullptr;  // inclusive running max of end within the chromosome run (lazy)
  uint32_t*   bmax_end = nullptr;  // max end of every 32-row block (global row index / 32; lazy, with pmax_end)
  std::vector<bk::ChromRun> runs;
  bool        pad_tie_disorder = false;  // bk_bed_pad: rows clamped to start 0 with equal ends are not in rest order (--everything over several files refuses)
};

namespace bk {

// Every entry point that takes a ctx runs on the ctx's device whatever the calling thread's current device is (a host
// with one ctx per GPU may call them from one thread), and leaves the caller's current device as it found it.
struct DeviceGuard {
  int  prev = -1;
  bool switched = false;
  explicit DeviceGuard(const bk_ctx* c) {
    if (!c) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int  fail(bk_ctx* ctx, int code, const char* fmt, ...);
int  cuda_fail(bk_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);
void  release_cached(bk_ctx* ctx);         // give every cached device block back to the driver
void* dmalloc(bk_ctx* ctx, size_t bytes);  // stream-ordered; returns nullptr and sets last_error on failure
void  dfree(bk_ctx* ctx, void* p);
// host -> device copy of n bytes on `st`.  Pinned / registered / device sources go down as one asynchronous copy; pageable
// memory (a tool's mmap of its input file) is moved through a ring of pinned staging buffers filled by several host
// threads, so that the copy runs at PCIe speed rather than at one core's page-fault-and-memcpy speed.  Returns after
// the last chunk has been QUEUED (the source may be reused once the stream reaches that point -- callers sync).
int   upload(bk_ctx* ctx, char* d_dst, const char* src, size_t n, cudaStream_t st);
constexpr size_t kParamRing = 256 << 10;
// device copy of a small host table, queued on the ctx stream without any synchronisation (through the pinned ring; tables
// that do not fit wait for the stream and restart the ring)
int   upload_params(bk_ctx* ctx, void* d_dst, const void* src, size_t n);
char* pinned_get(bk_ctx* ctx, size_t bytes);
void  pinned_put(bk_ctx* ctx, char* p);

#define BK_CUDA(ctx, expr)                                                      \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) return bk::cuda_fail((ctx), _e, #expr, __FILE__, __LINE__); \
  } while (0)

#define BK_LAUNCHED(ctx)                                                        \
  do {                                                                          \
    (ctx)->launches++;                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) return bk::cuda_fail((ctx), _e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define BK_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != BK_OK) return _rc; \
  } while (0)

void prof_begin(bk_ctx* ctx, const char* name);
void prof_end(bk_ctx* ctx);

template <typename T>
inline T* dalloc(bk_ctx* ctx, size_t n) {
  return reinterpret_cast<T*>(dmalloc(ctx, (n ? n : 1) * sizeof(T)));
}

// device scratch slots (u64 each)
enum {
  SC_ERR_CODE = 0,   // first error code seen by any kernel (atomicCAS from 0)
  SC_ERR_ROW = 1,    // row / byte offset attached to it
  SC_SPARE = 2,      // (was the tile ticket of the look-back kernels: no kernel waits on another CTA any more)
  SC_COUNT_A = 3,    // generic counters
  SC_COUNT_B = 4,
  SC_COUNT_C = 5,
  SC_COUNT_D = 6,
  SC_EFFLEN = 7,    // parser: bytes up to and including the last newline
  SC_NHEADS = 8,    // parser: chromosome run heads found
  SC_NROWS = 9,     // parser: total rows
  SC_OUT_BYTES = 10, // emitters: total bytes written
  SC_OUT_ROWS = 11,  // emitters: total rows written
  SC_N = 64
};
constexpr size_t kHostScratchExtra = 16384;  // pinned bytes behind the SC_N words of h_scratch (parser: first chromosome heads)

// ---------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void dev_set_error(uint64_t* scratch, int code, uint64_t where) {
  if (atomicCAS(reinterpret_cast<unsigned long long*>(&scratch[SC_ERR_CODE]), 0ull, (unsigned long long)code) == 0ull)
    scratch[SC_ERR_ROW] = where;
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only); both addresses 16-byte aligned
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- bulk asynchronous copy global -> shared (cp.async.bulk, SASS UBLKCP) completed on an mbarrier -------------------
// One elected thread arms the barrier with the byte count and issues ONE instruction for the whole window; the copy
// engine moves the bytes, no thread spends issue slots on them.  Addresses 16-byte aligned, size a multiple of 16.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst), b = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc),
               "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {  // release.cta: the arriving thread's earlier writes are visible to waiters
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
#ifndef BK_MBAR_HINT_NS
#define BK_MBAR_HINT_NS 0  // > 0: suspend-time hint of try_wait (the thread sleeps in hardware until the phase completes or the time is up)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
#if BK_MBAR_HINT_NS > 0
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(b),
      "r"(parity), "r"((uint32_t)BK_MBAR_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(b),
      "r"(parity)
      : "memory");
#endif
}

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {  // streaming 16-byte load, no L1 allocation
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

__device__ __forceinline__ bool is_ws(unsigned char c) {  // isspace() minus '\n'
  return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f';
}
__device__ __forceinline__ bool is_tok(unsigned char c) { return !(is_ws(c) || c == '\n'); }
__device__ __forceinline__ bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }

// warp inclusive scan (sum) of a 32-bit value
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// ---- binary searches over sorted u32 arrays -----------------------------------------------------------------
// first index in [lo,hi) with a[i] >= key
__device__ __forceinline__ uint64_t lower_bound_u32(const uint32_t* __restrict__ a, uint64_t lo, uint64_t hi, uint64_t key) {
  while (lo < hi) {
    uint64_t mid = lo + ((hi - lo) >> 1);
    if ((uint64_t)__ldg(&a[mid]) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first index in [lo,hi) with a[i] > key
__device__ __forceinline__ uint64_t upper_bound_u32(const uint32_t* __restrict__ a, uint64_t lo, uint64_t hi, uint64_t key) {
  while (lo < hi) {
    uint64_t mid = lo + ((hi - lo) >> 1);
    if ((uint64_t)__ldg(&a[mid]) <= key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// First index i in [from, n) with a[i] >= key (n if none); a is non-decreasing and `from` is a lower bound of the
// answer.  All 32 lanes call it with the same arguments.  Probe 1: 32 consecutive elements (the common case when
// the answer moved a few rows since the previous reference row).  Probe 2: the last element of each of the next 32
// blocks of 32 (1024 rows per probe), then one consecutive probe inside the block found.  Far jumps fall back to a
// binary search.
__device__ __forceinline__ uint32_t warp_gallop(const uint32_t* __restrict__ a, uint32_t from, uint32_t n, uint32_t key,
                                                int lane, bool near_first) {
  if (near_first) {
    if (from >= n) return n;
    const uint32_t k = from + lane;
    const uint32_t v = k < n ? __ldg(&a[k]) : 0xFFFFFFFFu;
    const unsigned m = __ballot_sync(0xffffffffu, k >= n || v >= key);
    if (m) {
      const uint32_t r = from + (__ffs(m) - 1);
      return r < n ? r : n;
    }
    from += 32;
  }
#pragma unroll 1
  for (int round = 0; round < 2; round++) {
    if (from >= n) return n;
    const uint32_t k = from + 32u * lane + 31u;
    const uint32_t v = k < n ? __ldg(&a[k]) : 0xFFFFFFFFu;
    const unsigned m = __ballot_sync(0xffffffffu, k >= n || v >= key);
    if (m) {
      const uint32_t base = from + 32u * (__ffs(m) - 1);
      const uint32_t k2 = base + lane;
      const uint32_t v2 = k2 < n ? __ldg(&a[k2]) : 0xFFFFFFFFu;
      const unsigned m2 = __ballot_sync(0xffffffffu, k2 >= n || v2 >= key);
      const uint32_t r = base + (__ffs(m2) - 1);
      return r < n ? r : n;
    }
    from += 1024;
  }
  uint32_t lo = from, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(&a[mid]) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// First index i in [0, n) with a[i] >= key (n if none), a non-decreasing: 32-ary search, every round the 32 lanes
// probe the last element of 32 equal sub-ranges (5 rounds for 2^25 rows instead of 25 dependent binary-search loads).
__device__ __forceinline__ uint32_t warp_search32(const uint32_t* __restrict__ a, uint32_t n, uint32_t key, int lane) {
  // invariant: the answer is in [from, from+len]; from+len means "every element of [from, from+len) is < key"
  uint32_t from = 0, len = n;
  while (len) {
    const uint32_t step = (len + 31u) >> 5;
    const uint32_t k = from + step * (uint32_t)(lane + 1) - 1u;  // last element of sub-range `lane`
    const bool     in = k < from + len;
    const uint32_t v = in ? __ldg(&a[k]) : 0xFFFFFFFFu;
    const unsigned m = __ballot_sync(0xffffffffu, !in || v >= key);
    if (m == 0) return from + len;
    const uint32_t j = (uint32_t)(__ffs(m) - 1);
    const uint32_t ub = min(from + step * (j + 1u) - 1u, from + len);  // a[ub] >= key, or ub is the old bound
    from += step * j;
    len = ub - from;
  }
  return from;
}

// exclusive scan of per-warp-range totals (one CTA of 1024 threads; n is a few thousand):
// base[i] = sum total[0..i), base[n] = scratch[SLOT] = the grand total
template <int SLOT>
__global__ void __launch_bounds__(1024) k_scan_totals(const uint64_t* __restrict__ total, uint64_t* __restrict__ base,
                                                      uint32_t n, uint64_t* scratch) {
  __shared__ uint64_t part[1024];
  const uint32_t tid = threadIdx.x, per = (n + 1023) / 1024;
  const uint32_t b = tid * per < n ? tid * per : n, e = b + per < n ? b + per : n;
  uint64_t       s = 0;
  for (uint32_t i = b; i < e; i++) s += total[i];
  part[tid] = s;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {  // Hillis-Steele over the 1024 partials
    uint64_t v = tid >= d ? part[tid - d] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  uint64_t runv = tid ? part[tid - 1] : 0;
  for (uint32_t i = b; i < e; i++) {
    base[i] = runv;
    runv += total[i];
  }
  if (tid == 1023) {
    base[n] = part[1023];
    scratch[SLOT] = part[1023];
  }
}

// ---- warp ranges inside chromosome runs (prefix-max index, segment compaction) ----------------------------------
// Rows are cut into ranges of PM_RANGE rows that never cross a chromosome boundary; one warp owns a range.
constexpr int PM_THREADS = 256, PM_RANGE = 2048;
struct PmRun {
  uint64_t row_begin, row_end, first_range;  // first_range: index of the chromosome's first warp range
};

// range r -> its rows [a,b), the first range of its chromosome, and the index of that chromosome in runs[]
__device__ __forceinline__ int pm_locate(const PmRun* __restrict__ runs, int nruns, uint64_t r, uint64_t& a, uint64_t& b,
                                         uint64_t& first) {
  int lo = 0, hi = nruns;  // last run with first_range <= r (empty runs own no range and are skipped by "last")
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (runs[mid].first_range <= r) lo = mid; else hi = mid;
  }
  first = runs[lo].first_range;
  a = runs[lo].row_begin + (r - first) * PM_RANGE;
  b = a + PM_RANGE < runs[lo].row_end ? a + PM_RANGE : runs[lo].row_end;
  return lo;
}

// sum of a 64-bit value over the warp from four 16-bit limbs (REDUX is 32-bit)
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
  const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  const uint64_t r0 = __reduce_add_sync(0xffffffffu, lo & 0xFFFFu), r1 = __reduce_add_sync(0xffffffffu, lo >> 16);
  const uint64_t r2 = __reduce_add_sync(0xffffffffu, hi & 0xFFFFu), r3 = __reduce_add_sync(0xffffffffu, hi >> 16);
  return r0 + (r1 << 16) + (r2 << 32) + (r3 << 48);
}

#endif  // __CUDACC__

}  // namespace bk
