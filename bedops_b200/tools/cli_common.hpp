// cli_common.hpp -- shared host-side plumbing of the drop-in command-line tools (bedmap, bedops,
// closest-features).  The tools keep the reference's argv grammar, stdout/stderr text and exit codes
//  and hand all compute to libbedkit.so through include/bedkit.h.  There is no CPU compute path:
// when the library cannot reach a B200 the tool fails with the library's message.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <atomic>
#include <thread>
#include <string>
#include <vector>
#include <memory>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cerrno>
#include <sys/mman.h>
#include <condition_variable>
#include <mutex>
#include <ctime>
#include "bedkit.h"

namespace cli {

static const char* kVersion = "2.4.26";  // interfaces/general-headers/suite/BEDOPS.Version.hpp:36-42
static const char* kCitation = "http://bioinformatics.oxfordjournals.org/content/28/14/1919.abstract";
static const char* kAuthors = "Shane Neph & Scott Kuehn";

struct UserError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void banner(FILE* f, const char* prog) {
  std::fprintf(f, "%s\n  citation: %s\n  version:  %s\n  authors:  %s\n", prog, kCitation, kVersion, kAuthors);
}

inline bool only_chars(const std::string& s, const char* set) { return s.find_first_not_of(set) == std::string::npos; }

// A whole input: a regular file is mapped (the library stages it to the GPU in parallel chunks, no intermediate copy in
// the tool), stdin ("-") and pipes are read into memory.
struct Input {
  const char*       data = nullptr;
  size_t            size = 0;
  std::vector<char> owned;
  void*             map = nullptr;
  size_t            map_len = 0;
  Input() = default;
  Input(const Input&) = delete;
  Input& operator=(const Input&) = delete;
  ~Input() {
    if (map) ::munmap(map, map_len);
  }
  void settle() {}  // (a background page-toucher lived here: it contended with the CUDA context creation for the mm lock and lost)
  bool open(const std::string& name) {  // false if the file cannot be opened
    int fd = name == "-" ? 0 : ::open(name.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fd && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
      void* p = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p != MAP_FAILED) {
        ::madvise(p, (size_t)st.st_size, MADV_WILLNEED);
        map = p;
        map_len = (size_t)st.st_size;
        data = static_cast<const char*>(p);
        size = map_len;
        ::close(fd);
        return true;
      }
    }
    owned.resize(1u << 20);
    size_t n = 0;
    while (true) {
      if (n == owned.size()) owned.resize(owned.size() * 2);
      ssize_t r = ::read(fd, owned.data() + n, owned.size() - n);
      if (r < 0) {
        if (fd) ::close(fd);
        return false;
      }
      if (r == 0) break;
      n += (size_t)r;
    }
    owned.resize(n);
    if (fd) ::close(fd);
    data = owned.data();
    size = n;
    return true;
  }
  // the checking iterator of the reference also reads an unterminated last line: give the text its final NL
  void ensure_final_newline() {
    if (size == 0 || data[size - 1] == '\n') return;
    if (owned.empty() || data != owned.data()) owned.assign(data, data + size);
    owned.push_back('\n');
    data = owned.data();
    size = owned.size();
  }
};

// BEDKIT_TRACE: wall clock of the tool's phases to stderr (process start = first call)
inline void trace_lap(const char* what) {
  static const bool on = std::getenv("BEDKIT_TRACE") != nullptr;
  if (!on) return;
  static timespec t0{};
  timespec        t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  if (!t0.tv_sec) t0 = t;
  std::fprintf(stderr, "[bedkit-tool] %-28s %8.1f ms\n", what, (t.tv_sec - t0.tv_sec) * 1e3 + (t.tv_nsec - t0.tv_nsec) * 1e-6);
}

// The result is on stdout: leave without unwinding the CUDA context block by block (the driver releases everything with
// the process; the orderly teardown costs 0.2-0.6 s of a run that computes for 0.1 s).
[[noreturn]] inline void finish_now(int code) {
  std::fflush(stdout);
  std::fflush(stderr);
  trace_lap("exit");
  ::_exit(code);
}

inline void write_all(const char* p, size_t n) {  // straight to fd 1: the result text is not copied again
  std::fflush(stdout);
  while (n) {
    ssize_t w = ::write(1, p, n > (1u << 30) ? (1u << 30) : n);
    if (w < 0) {
      if (errno == EINTR) continue;
      break;
    }
    p += w;
    n -= (size_t)w;
  }
}

struct Engine {
  bk_ctx* ctx = nullptr;
  explicit Engine(int shard = 0) {
    const char* dev = std::getenv("BEDKIT_DEVICE");
    if (std::getenv("BEDKIT_SHARE_DEVICE")) shard = 0;  // all shards on one GPU (tests on a single-GPU box)
    int         rc = bk_init(&ctx, (dev ? std::atoi(dev) : 0) + shard);
    if (rc != BK_OK) throw std::runtime_error(bk_strerror(rc));
  }
  ~Engine() {
    bk_destroy(ctx);
    trace_lap("bk_destroy done");
  }
  [[noreturn]] void raise(int rc) const {
    const char* detail = bk_last_error(ctx);
    throw std::runtime_error(detail && *detail ? std::string(detail) : std::string(bk_strerror(rc)));
  }
  bk_bed* load(const Input& text, int min_fields, unsigned cols) const { return load(text.data, text.size, min_fields, cols); }
  bk_bed* load(const char* p, size_t n, int min_fields, unsigned cols) const {
    bk_bed* b = nullptr;
    int     rc = bk_load_bed(ctx, p, n, min_fields, cols, &b);
    if (rc != BK_OK) raise(rc);
    return b;
  }
};

// Starch archives are accepted wherever BED is (AllocateIterator_BED_starch.hpp:100-112): the library inflates and
// un-transforms the archive (bk_unstarch), the tool goes on with the BED text
inline void unstarch_if_archive(const Engine& eng, Input& in) {
  if (!bk_is_starch(in.data, in.size)) return;
  bk_text out{};
  int     rc = bk_unstarch(eng.ctx, in.data, in.size, nullptr, 0, &out);
  if (rc != BK_OK) eng.raise(rc);
  std::vector<char> text(out.ptr, out.ptr + out.len);
  bk_free_text(eng.ctx, &out);
  in.settle();
  if (in.map) ::munmap(in.map, in.map_len);
  in.map = nullptr;
  in.owned.swap(text);
  in.data = in.owned.data();
  in.size = in.owned.size();
}
inline bool any_archive(std::initializer_list<const Input*> ins) {
  for (const Input* i : ins)
    if (bk_is_starch(i->data, i->size)) return true;
  return false;
}

// --ec / --header: run the device validation (bk_check_text) on a text that ends with NL (Input::ensure_final_newline);
// failures are reported the way BedCheckIterator.hpp:589-593 words them.
inline void ec_prepare(Input& text) { text.ensure_final_newline(); }
inline void ec_check(const Engine& eng, const Input& text, const std::string& name, int n_fields, bool has_rest, bool nest_check) {
  int rc = bk_check_text(eng.ctx, text.data, text.size, n_fields, has_rest ? 1 : 0, nest_check ? 1 : 0);
  if (rc == BK_ERR_CHECK) throw std::runtime_error("in " + name + "\n" + bk_last_error(eng.ctx));
  if (rc != BK_OK) eng.raise(rc);
}

// ---- multi-GPU: BEDKIT_GPUS=N shards the inputs by contiguous chromosome groups, one host thread + one bk_ctx per
// GPU, outputs concatenated in shard order (byte-identical to the unsharded run) -------------------------
struct Slice {
  const char* ptr;
  size_t      len;
};

inline int gpus_requested() {
  const char* g = std::getenv("BEDKIT_GPUS");
  int         n = g ? std::atoi(g) : 1;
  return n < 1 ? 1 : n;
}

// slices[shard][file]
// Whole-chromosome shards (bedops, closest-features).  Returns an empty plan when a file's chromosomes are not in strictly
// ascending strcmp order (a UCSC header line read as a chromosome, or unsorted input): the caller then runs unsharded and
// the single-GPU path reports what is wrong.
inline std::vector<std::vector<Slice>> plan_slices(const std::vector<const Input*>& files, int n_shards) {
  std::vector<std::vector<bk_chrom_span>> idx(files.size());
  std::vector<std::string>                names;
  for (size_t f = 0; f < files.size(); f++) {
    int n = 0, cap = 64;
    while (true) {
      idx[f].resize(cap);
      int rc = bk_chrom_index(files[f]->data, files[f]->size, idx[f].data(), cap, &n);
      if (rc == BK_OK) break;
      if (rc != BK_ERR_NOMEM) throw std::runtime_error("chromosome index failed");
      cap = n + 8;
    }
    idx[f].resize(n);
    for (int k = 1; k < n; k++)
      if (std::strcmp(idx[f][k - 1].name, idx[f][k].name) >= 0) return {};
    for (auto& s : idx[f]) names.push_back(s.name);
  }
  std::sort(names.begin(), names.end(), [](const std::string& a, const std::string& b) { return std::strcmp(a.c_str(), b.c_str()) < 0; });
  names.erase(std::unique(names.begin(), names.end()), names.end());
  std::vector<uint64_t> load(names.size(), 0);
  auto pos_of = [&](const char* nm) {
    return (size_t)(std::lower_bound(names.begin(), names.end(), std::string(nm),
                                     [](const std::string& a, const std::string& b) { return std::strcmp(a.c_str(), b.c_str()) < 0; }) -
                    names.begin());
  };
  for (auto& ix : idx)
    for (auto& s : ix) load[pos_of(s.name)] += s.end - s.begin;
  std::vector<int> first(n_shards + 1, 0);
  if (bk_plan_shards(load.data(), (int)load.size(), n_shards, first.data()) != BK_OK) throw std::runtime_error("shard plan failed");
  std::vector<std::vector<Slice>> out(n_shards, std::vector<Slice>(files.size(), Slice{nullptr, 0}));
  for (int s = 0; s < n_shards; s++)
    for (size_t f = 0; f < files.size(); f++) {
      uint64_t b = 0, e = 0;
      bool     any = false;
      for (auto& sp : idx[f]) {
        size_t p = pos_of(sp.name);
        if ((int)p >= first[s] && (int)p < first[s + 1]) {
          if (!any) b = sp.begin;
          e = sp.end;
          any = true;
        }
      }
      out[s][f] = Slice{files[f]->data + b, (size_t)(e - b)};
    }
  return out;
}

// run fn(engine, slices_of_this_shard) -> result text on n_shards GPUs concurrently; write the outputs in shard order
template <class Fn>
inline void run_sharded(const std::vector<std::vector<Slice>>& slices, Fn fn, int n_gpus) {
  // more shards than GPUs (plan_slices is asked for 4 per GPU): every GPU thread pulls the next shard, which evens
  // out chromosomes of very different sizes; the outputs are written in shard (= chromosome) order
  const int                n = (int)slices.size();
  std::vector<std::string> outs(n), errs(n_gpus);
  std::atomic<int>         next{0};
  std::vector<std::thread> th;
  for (int g = 0; g < n_gpus; g++)
    th.emplace_back([&, g]() {
      try {
        Engine eng(g);
        for (int s = next.fetch_add(1); s < n; s = next.fetch_add(1)) outs[s] = fn(eng, slices[s]);
      } catch (const std::exception& e) {
        errs[g] = e.what();
        if (errs[g].empty()) errs[g] = "unknown error";
        next.store(n);  // the other threads stop at their next pull
      }
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw std::runtime_error(e);
  for (auto& o : outs) write_all(o.data(), o.size());
}

// ---- bedmap over N GPUs by genomic RANGES (cuts inside chromosomes, boundary halos; include/bedkit.h) -----------------
// One host thread + one bk_ctx per GPU.  begin (upload + parse own slices, report halo coordinates) -> all threads meet ->
// finish (left halo, map) -> outputs written in rank order straight from the pinned result buffers.
// Returns false when the inputs cannot be range-sharded (chromosomes not strictly ascending): run unsharded.
inline bool run_range_sharded_bedmap(const Input& ref, const Input& map, int ref_fields, unsigned ref_cols, int map_fields,
                                     unsigned map_cols, const bk_mapspec& spec, int n_gpus) {
  bk_shard_plan plan;
  if (n_gpus > BK_MAX_SHARDS || bk_shard_plan_make(ref.data, ref.size, map.data, map.size, n_gpus, &plan) != BK_OK) return false;
  std::vector<uint64_t>    reach((size_t)n_gpus * n_gpus, ~0ull);
  std::vector<std::string> errs(n_gpus);
  std::vector<bk_text>     outs(n_gpus);
  std::vector<std::unique_ptr<Engine>> engines(n_gpus);
  std::mutex              mu;
  std::condition_variable cv;
  int                     arrived = 0;
  bool                    failed = false;
  std::vector<std::thread> th;
  for (int g = 0; g < n_gpus; g++)
    th.emplace_back([&, g]() {
      bk_shard* sh = nullptr;
      std::memset(&outs[g], 0, sizeof(bk_text));
      try {
        engines[g].reset(new Engine(g));
        int rc = bk_bedmap_shard_begin(engines[g]->ctx, &plan, g, ref.data, ref.size, ref_fields, ref_cols, map.data, map.size,
                                       map_fields, map_cols, nullptr, nullptr, &spec, &sh, &reach[(size_t)g * n_gpus]);
        if (rc != BK_OK) engines[g]->raise(rc);
      } catch (const std::exception& e) {
        errs[g] = *e.what() ? e.what() : "unknown error";
      }
      {  // everyone meets here, failed or not: the reach matrix is complete behind this point
        std::unique_lock<std::mutex> lk(mu);
        failed = failed || !errs[g].empty();
        if (++arrived == n_gpus) cv.notify_all();
        else cv.wait(lk, [&] { return arrived == n_gpus; });
      }
      if (!failed && sh) {
        try {
          int rc = bk_bedmap_shard_finish(engines[g]->ctx, sh, reach.data(), &outs[g]);
          if (rc != BK_OK) engines[g]->raise(rc);
        } catch (const std::exception& e) {
          errs[g] = *e.what() ? e.what() : "unknown error";
        }
      }
      if (sh && engines[g]) bk_shard_free(engines[g]->ctx, sh);
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw std::runtime_error(e);
  for (int g = 0; g < n_gpus; g++) {
    write_all(outs[g].ptr, outs[g].len);
    bk_free_text(engines[g]->ctx, &outs[g]);
  }
  return true;
}

}  // namespace cli
