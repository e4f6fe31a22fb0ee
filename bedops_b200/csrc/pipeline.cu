// pipeline.cu -- bk_bedmap_host: bedmap over HOST text with the PCIe transfers overlapped with the kernels.
//
// The path shards by chromosome, so one call can be cut into independent chromosome groups:
//   copy stream    : H2D of every group's slices of the two files, back to back (enqueued up front)
//   compute stream : for each group, as soon as its slices have arrived: parse both, map, emit (result stays in HBM)
//   result stream  : D2H of each group's output text into one pinned host buffer at its final offset
// PCIe is full duplex, the kernels take ~10 % of the transfer time, so the call approaches the H2D time of the inputs.
// Output = the groups' outputs in chromosome order = the output of one unsplit call (same contract as the reference's
// per-chromosome scale-out, bedmap/src/Input.hpp:117-122).  Inputs should be pinned for the overlap to happen.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include "common.cuh"

using namespace bk;

namespace {

struct Group {
  uint64_t rb, re, mb, me;  // byte ranges in the reference / map text
  uint64_t d_ref, d_map;    // offsets of the slices in the device staging buffer (256-byte aligned)
};

bool strictly_sorted(const std::vector<bk_chrom_span>& ix) {
  for (size_t i = 1; i < ix.size(); i++)
    if (strcmp(ix[i - 1].name, ix[i].name) >= 0) return false;
  return true;
}

int index_of(const char* text, size_t n, std::vector<bk_chrom_span>* out) {
  int cnt = 0, cap = 64;
  while (true) {
    out->assign(cap, bk_chrom_span{});
    int rc = bk_chrom_index(text, n, out->data(), cap, &cnt);
    if (rc == BK_OK) {
      out->resize(cnt);
      return BK_OK;
    }
    if (rc != BK_ERR_NOMEM) return rc;
    cap = cnt + 8;
  }
}

int plain(bk_ctx* ctx, const char* ref_text, size_t ref_len, const char* map_text, size_t map_len, int ref_fields,
          unsigned ref_cols, int map_fields, unsigned map_cols, const bk_mapspec* spec, bk_text* out) {
  bk_bed *ref = nullptr, *map = nullptr;
  int     rc = bk_load_bed(ctx, ref_text, ref_len, ref_fields, ref_cols, &ref);
  if (rc == BK_OK) rc = bk_load_bed(ctx, map_text, map_len, map_fields, map_cols, &map);
  if (rc == BK_OK) rc = bk_bedmap(ctx, ref, map, spec, out);
  bk_free_bed(ctx, ref);
  bk_free_bed(ctx, map);
  return rc;
}

}  // namespace

extern "C" int bk_bedmap_host(bk_ctx* ctx, const char* ref_text, size_t ref_len, int ref_fields, unsigned ref_cols,
                              const char* map_text, size_t map_len, int map_fields, unsigned map_cols,
                              const bk_mapspec* spec, bk_text* out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !spec || !out || (!ref_text && ref_len) || (!map_text && map_len)) return BK_ERR_ARG;
  memset(out, 0, sizeof(*out));
  const bool one_chrom = spec->chrom && strcmp(spec->chrom, "all") != 0;
  bool       element_ops = false;  // --max/min-element may abort the run in the middle of the output: one unsplit call
  for (int c = 0; c < spec->n_ops && c < BK_MAX_OPS; c++)
    element_ops |= spec->ops[c] == BK_OP_MAX_ELEMENT || spec->ops[c] == BK_OP_MIN_ELEMENT;
  std::vector<bk_chrom_span> rix, mix;
  if (one_chrom || element_ops || spec->out_on_device || index_of(ref_text, ref_len, &rix) != BK_OK ||
      index_of(map_text, map_len, &mix) != BK_OK || !strictly_sorted(rix) || !strictly_sorted(mix) || rix.size() < 2)
    return plain(ctx, ref_text, ref_len, map_text, map_len, ref_fields, ref_cols, map_fields, map_cols, spec, out);

  // chromosome groups of roughly 1/16 of the bytes each (>= 32 MiB); chromosomes without reference rows are not uploaded
  uint64_t total = 0;
  std::vector<std::pair<uint64_t, uint64_t>> mspan(rix.size(), {0, 0});
  {
    size_t j = 0;
    for (size_t i = 0; i < rix.size(); i++) {
      while (j < mix.size() && strcmp(mix[j].name, rix[i].name) < 0) j++;
      if (j < mix.size() && strcmp(mix[j].name, rix[i].name) == 0) mspan[i] = {mix[j].begin, mix[j].end};
      total += (rix[i].end - rix[i].begin) + (mspan[i].second - mspan[i].first);
    }
  }
  const uint64_t     target = std::max<uint64_t>(total / 16, 32ull << 20);
  std::vector<Group> groups;
  uint64_t           dev_bytes = 0;
  for (size_t i = 0; i < rix.size();) {
    Group    g{rix[i].begin, rix[i].end, 0, 0, 0, 0};
    uint64_t load = 0;
    bool     have_map = false;
    size_t   k = i;
    for (; k < rix.size(); k++) {
      const uint64_t add = (rix[k].end - rix[k].begin) + (mspan[k].second - mspan[k].first);
      if (k > i && load + add > target) break;
      // the map slice of a group must be one contiguous byte range: stop in front of a gap made by chromosomes that
      // exist only in the map file
      if (mspan[k].second > mspan[k].first) {
        if (have_map && mspan[k].first != g.me) break;
        if (!have_map) g.mb = mspan[k].first;
        g.me = mspan[k].second;
        have_map = true;
      }
      g.re = rix[k].end;
      load += add;
    }
    auto up = [](uint64_t x) { return (x + 255) & ~255ull; };
    g.d_ref = dev_bytes;
    dev_bytes = up(dev_bytes + (g.re - g.rb)) + 256;
    g.d_map = dev_bytes;
    dev_bytes = up(dev_bytes + (g.me - g.mb)) + 256;
    groups.push_back(g);
    i = k;
  }

  ctx->last_error.clear();
  char* d_stage = reinterpret_cast<char*>(dmalloc(ctx, dev_bytes + 256));
  if (!d_stage) return BK_ERR_NOMEM;
  cudaStream_t copy_s = nullptr, out_s = nullptr;
  std::vector<cudaEvent_t> ev_in(groups.size(), nullptr), ev_done(groups.size(), nullptr);
  std::vector<bk_text>     parts(groups.size());
  for (auto& t : parts) memset(&t, 0, sizeof(t));
  int  rc = BK_OK;
  auto cuda_ok = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == BK_OK) rc = cuda_fail(ctx, e, what, __FILE__, __LINE__);
    return e == cudaSuccess;
  };
  cuda_ok(cudaStreamCreateWithFlags(&copy_s, cudaStreamNonBlocking), "cudaStreamCreate");
  cuda_ok(cudaStreamCreateWithFlags(&out_s, cudaStreamNonBlocking), "cudaStreamCreate");
  // the staging block may still be in use by earlier work of this ctx (cached allocator, stream order): start the
  // copies after everything already queued on the compute stream
  cudaEvent_t ev_start = nullptr;
  cuda_ok(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming), "cudaEventCreate");
  if (rc == BK_OK) {
    cuda_ok(cudaEventRecord(ev_start, ctx->stream), "cudaEventRecord");
    cuda_ok(cudaStreamWaitEvent(copy_s, ev_start, 0), "cudaStreamWaitEvent");
  }
  for (size_t g = 0; g < groups.size() && rc == BK_OK; g++) {
    cuda_ok(cudaEventCreateWithFlags(&ev_in[g], cudaEventDisableTiming), "cudaEventCreate");
    cuda_ok(cudaEventCreateWithFlags(&ev_done[g], cudaEventDisableTiming), "cudaEventCreate");
  }
  // The uploads run on their own host thread: from pinned memory they are plain asynchronous copies, from pageable memory
  // (the tools' mmap of the input files) they are staged through pinned buffers by upload(), which keeps a host thread
  // busy -- either way the calling thread is free to queue the kernels of the groups that have arrived.
  std::vector<std::atomic<int>> arrived(groups.size());
  for (auto& a : arrived) a.store(0);
  std::atomic<int> up_rc{BK_OK};
  std::thread      uploader;
  if (rc == BK_OK)
    uploader = std::thread([&]() {
      cudaSetDevice(ctx->device);
      for (size_t g = 0; g < groups.size(); g++) {
        const Group& G = groups[g];
        int          r = BK_OK;
        if (up_rc.load() == BK_OK) {
          r = upload(ctx, d_stage + G.d_ref, ref_text + G.rb, G.re - G.rb, copy_s);
          if (r == BK_OK) r = upload(ctx, d_stage + G.d_map, map_text + G.mb, G.me - G.mb, copy_s);
          if (r == BK_OK && cudaEventRecord(ev_in[g], copy_s) != cudaSuccess) r = BK_ERR_CUDA;
          if (r != BK_OK) up_rc.store(r);
        }
        arrived[g].store(1, std::memory_order_release);
      }
    });

  // compute, group by group; results stay in HBM until their D2H has been queued with a known offset
  bk_mapspec dspec = *spec;
  dspec.out_on_device = 1;
  char*    h_out = nullptr;
  uint64_t h_cap = 0, h_off = 0, rows = 0, rowids = 0;
  int      rowid_ops = 0;
  for (int c = 0; c < spec->n_ops; c++) rowid_ops += spec->ops[c] == BK_OP_ECHO_REF_ROW_ID;
  size_t   flushed = 0;  // parts [0, flushed) have their D2H queued
  uint64_t ref_done = 0, out_done = 0;
  auto     flush = [&](size_t upto) {
    for (; flushed < upto && rc == BK_OK; flushed++) {
      bk_text& t = parts[flushed];
      if (t.len && h_off + t.len <= h_cap) {
        cuda_ok(cudaStreamWaitEvent(out_s, ev_done[flushed], 0), "cudaStreamWaitEvent");
        cuda_ok(cudaMemcpyAsync(h_out + h_off, t.ptr, t.len, cudaMemcpyDeviceToHost, out_s), "D2H");
      } else if (t.len) {
        break;  // does not fit the estimate: the tail is copied after the final size is known
      }
      h_off += t.len;
    }
  };
  for (size_t g = 0; g < groups.size() && rc == BK_OK; g++) {
    const Group& G = groups[g];
    while (!arrived[g].load(std::memory_order_acquire)) std::this_thread::yield();  // its copies and event are queued
    if (up_rc.load() != BK_OK) {
      rc = up_rc.load();
      break;
    }
    cuda_ok(cudaStreamWaitEvent(ctx->stream, ev_in[g], 0), "cudaStreamWaitEvent");
    bk_bed *ref = nullptr, *map = nullptr;
    if (rc == BK_OK) rc = bk_load_bed_device(ctx, d_stage + G.d_ref, G.re - G.rb, ref_fields, ref_cols, &ref);
    if (rc == BK_OK) rc = bk_load_bed_device(ctx, d_stage + G.d_map, G.me - G.mb, map_fields, map_cols, &map);
    dspec.row_id_base = spec->row_id_base + rowids;  // ids already issued by the earlier groups (each printed id bumps the counter)
    if (rc == BK_OK) rc = bk_bedmap(ctx, ref, map, &dspec, &parts[g]);
    if (rc == BK_OK) rowids += parts[g].rows * (uint64_t)rowid_ops;
    if (rc == BK_OK) cuda_ok(cudaEventRecord(ev_done[g], ctx->stream), "cudaEventRecord");
    bk_free_bed(ctx, ref);
    bk_free_bed(ctx, map);
    if (rc != BK_OK) break;
    rows += parts[g].rows;
    ref_done += G.re - G.rb;
    out_done += parts[g].len;
    if (!h_out && (out_done || g + 1 == groups.size())) {  // size the host buffer from the output density so far, +25 %
      const double per_byte = (double)out_done / (double)std::max<uint64_t>(1, ref_done);
      h_cap = (uint64_t)(per_byte * 1.25 * (double)ref_len) + (1u << 20);
      h_out = pinned_get(ctx, h_cap);
      if (!h_out) rc = BK_ERR_NOMEM;
    }
    if (rc == BK_OK && h_out) flush(g + 1);
  }
  if (rc == BK_OK) {
    uint64_t need = 0;
    for (auto& t : parts) need += t.len;
    if (flushed < parts.size()) {  // the estimate was too small: move to an exact buffer
      cuda_ok(cudaStreamSynchronize(out_s), "cudaStreamSynchronize");
      char* exact = pinned_get(ctx, need);
      if (!exact) rc = BK_ERR_NOMEM;
      if (rc == BK_OK) {
        memcpy(exact, h_out, h_off);
        pinned_put(ctx, h_out);
        h_out = exact;
        h_cap = need;
        flush(parts.size());
      }
    }
    cuda_ok(cudaStreamSynchronize(out_s), "cudaStreamSynchronize");
    if (rc == BK_OK) {
      out->ptr = h_out;
      out->len = need;
      out->rows = rows;
      out->on_device = 0;
    }
  }
  // teardown: nothing may be reused while a stream still reads it
  if (uploader.joinable()) uploader.join();
  if (copy_s) cudaStreamSynchronize(copy_s);
  if (out_s) cudaStreamSynchronize(out_s);
  cudaStreamSynchronize(ctx->stream);
  for (auto& t : parts)
    if (t.ptr) bk_free_text(ctx, &t);
  dfree(ctx, d_stage);
  for (auto e : ev_in) if (e) cudaEventDestroy(e);
  for (auto e : ev_done) if (e) cudaEventDestroy(e);
  if (ev_start) cudaEventDestroy(ev_start);
  if (copy_s) cudaStreamDestroy(copy_s);
  if (out_s) cudaStreamDestroy(out_s);
  if (rc != BK_OK && h_out) pinned_put(ctx, h_out);
  // a failing group reports row numbers relative to its slice: let the unsplit path produce the message
  if (rc != BK_OK && rc != BK_ERR_NOMEM && rc != BK_ERR_CUDA)
    return plain(ctx, ref_text, ref_len, map_text, map_len, ref_fields, ref_cols, map_fields, map_cols, spec, out);
  return rc;
}

// ---- range-sharded bedmap (protocol: include/bedkit.h) ---------------------------------------------------------------
struct bk_shard {
  bk_shard_plan plan;
  int           rank = 0;
  const char *  map_text = nullptr, *map_src = nullptr;
  int           map_fields = 5;
  unsigned      map_cols = 0;
  bk_mapspec    spec;
  std::string   delim, mdelim;
  bk_bed *      ref = nullptr, *map = nullptr;
  uint64_t      bytes_in = 0;
};

namespace {

// copy [off, off+len) of src (host, pinned host or device memory) into a fresh device block and parse it
int load_slice(bk_ctx* ctx, const char* src, uint64_t off, uint64_t len, int fields, unsigned cols, bk_bed** out) {
  *out = nullptr;
  char* d = reinterpret_cast<char*>(dmalloc(ctx, len + 64));
  if (!d) return BK_ERR_NOMEM;
  {
    const int rc = upload(ctx, d, src + off, len, ctx->stream);
    if (rc != BK_OK) {
      dfree(ctx, d);
      return rc;
    }
  }
  int rc = bk_load_bed_device(ctx, d, len, fields, cols, out);
  if (rc != BK_OK) {
    dfree(ctx, d);
    return rc;
  }
  (*out)->owns_text = true;  // the slice belongs to the bed from here on
  return BK_OK;
}

uint64_t range_pad(const bk_mapspec& s) { return s.overlap_kind == BK_OVR_RANGE ? s.overlap_bp : 0; }

struct Lap {  // BEDKIT_TRACE: host wall clock of the sub-steps of a call, to stderr
  const char* what;
  bool        on = getenv("BEDKIT_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  explicit Lap(const char* w) : what(w) {}
  void operator()(const char* step) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[bedkit] %s: %s %.3f ms\n", what, step, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

}  // namespace

extern "C" int bk_shard_plan_make(const char* ref_text, size_t ref_len, const char* map_text, size_t map_len, int n_shards,
                                  bk_shard_plan* plan) {
  if (!plan || n_shards < 1 || n_shards > BK_MAX_SHARDS || (!ref_text && ref_len) || (!map_text && map_len)) return BK_ERR_ARG;
  memset(plan, 0, sizeof(*plan));
  plan->n_shards = n_shards;
  std::vector<bk_chrom_span> rix, mix;
  int rc = index_of(ref_text, ref_len, &rix);
  if (rc == BK_OK) rc = index_of(map_text, map_len, &mix);
  if (rc != BK_OK) return rc;
  if (!strictly_sorted(rix) || !strictly_sorted(mix)) return BK_ERR_UNSORTED;
  const bool by_map = map_len >= ref_len;  // cut where the bytes are
  if (n_shards > 1) {
    rc = by_map ? bk_plan_cuts(map_text, map_len, mix.data(), (int)mix.size(), n_shards, plan->cuts)
                : bk_plan_cuts(ref_text, ref_len, rix.data(), (int)rix.size(), n_shards, plan->cuts);
    if (rc != BK_OK) return rc;
  }
  plan->ref_off[0] = rix.empty() ? 0 : rix.front().begin;
  plan->map_off[0] = mix.empty() ? 0 : mix.front().begin;
  for (int k = 0; k + 1 < n_shards; k++) {
    plan->ref_off[k + 1] = bk_cut_offset(ref_text, ref_len, rix.data(), (int)rix.size(), &plan->cuts[k]);
    plan->map_off[k + 1] = bk_cut_offset(map_text, map_len, mix.data(), (int)mix.size(), &plan->cuts[k]);
    plan->map_chrom_begin[k] = plan->map_chrom_end[k] = plan->map_off[k + 1];
    for (auto& sp : mix)
      if (!plan->cuts[k].at_end && strcmp(sp.name, plan->cuts[k].chrom) == 0) {
        plan->map_chrom_begin[k] = sp.begin;
        plan->map_chrom_end[k] = sp.end;
      }
  }
  plan->ref_off[n_shards] = rix.empty() ? 0 : rix.back().end;
  plan->map_off[n_shards] = mix.empty() ? 0 : mix.back().end;
  return BK_OK;
}

extern "C" void bk_shard_free(bk_ctx* ctx, bk_shard* sh) {
  bk::DeviceGuard device_guard(ctx);
  if (!sh) return;
  bk_free_bed(ctx, sh->ref);
  bk_free_bed(ctx, sh->map);
  delete sh;
}

extern "C" uint64_t bk_shard_bytes_in(const bk_shard* sh) { return sh ? sh->bytes_in : 0; }

extern "C" int bk_bedmap_shard_begin(bk_ctx* ctx, const bk_shard_plan* plan, int rank, const char* ref_text, size_t ref_len,
                                     int ref_fields, unsigned ref_cols, const char* map_text, size_t map_len, int map_fields,
                                     unsigned map_cols, const char* ref_src, const char* map_src, const bk_mapspec* spec,
                                     bk_shard** out, uint64_t* reach) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !plan || !spec || !out || !reach || rank < 0 || rank >= plan->n_shards) return BK_ERR_ARG;
  (void)ref_len;
  (void)map_len;
  *out = nullptr;
  ctx->last_error.clear();
  const int n = plan->n_shards;
  for (int j = 0; j < n; j++) reach[j] = ~0ull;
  if (spec->chrom && strcmp(spec->chrom, "all") != 0) return fail(ctx, BK_ERR_ARG, "range sharding and --chrom exclude each other");
  bk_shard* sh = new bk_shard();
  sh->plan = *plan;
  sh->rank = rank;
  sh->map_text = map_text;
  sh->map_src = map_src ? map_src : map_text;
  sh->map_fields = map_fields;
  sh->map_cols = map_cols;
  sh->spec = *spec;
  sh->delim = spec->delim ? spec->delim : "|";
  sh->mdelim = spec->multidelim ? spec->multidelim : ";";
  sh->spec.delim = sh->delim.c_str();
  sh->spec.multidelim = sh->mdelim.c_str();
  sh->spec.chrom = nullptr;
  const uint64_t pad = range_pad(*spec);
  Lap            lap("shard_begin");
  // this rank's reference rows: the records whose start lies in its range
  const uint64_t r0 = plan->ref_off[rank], r1 = plan->ref_off[rank + 1];
  int rc = load_slice(ctx, ref_src ? ref_src : ref_text, r0, r1 - r0, ref_fields, ref_cols, &sh->ref);
  sh->bytes_in += r1 - r0;
  lap("ref slice");
  // right halo: map records behind the right cut that start before the largest reference end of the shard
  uint64_t m0 = plan->map_off[rank], m1 = plan->map_off[rank + 1];
  if (rc == BK_OK && rank + 1 < n && !plan->cuts[rank].at_end && plan->cuts[rank].coord > 0) {
    uint64_t maxend = 0;
    rc = bk_bed_chrom_max_end(ctx, sh->ref, plan->cuts[rank].chrom, &maxend);
    if (rc == BK_OK && maxend + pad > plan->cuts[rank].coord)
      m1 = bk_find_start(map_text, m1, plan->map_chrom_end[rank], maxend + pad);
  }
  lap("right halo bound");
  if (rc == BK_OK) rc = load_slice(ctx, sh->map_src, m0, m1 - m0, map_fields, map_cols, &sh->map);
  sh->bytes_in += m1 - m0;
  lap("map slice");
  // what the later shards need from this one: where, among these records, their left halo would begin
  for (int j = rank + 1; j < n && rc == BK_OK; j++) {
    const bk_cut& c = plan->cuts[j - 1];
    if (c.at_end || c.coord == 0) continue;
    uint64_t s = ~0ull;
    rc = bk_bed_reach_start(ctx, sh->map, c.chrom, c.coord > pad ? c.coord - pad : 0, &s);
    if (s < c.coord) reach[j] = s;
  }
  lap("reach");
  if (rc != BK_OK) {
    bk_shard_free(ctx, sh);
    return rc;
  }
  *out = sh;
  return BK_OK;
}

extern "C" int bk_bedmap_shard_finish(bk_ctx* ctx, bk_shard* sh, const uint64_t* all_reach, bk_text* out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !sh || !out || (!all_reach && sh->rank > 0)) return BK_ERR_ARG;
  memset(out, 0, sizeof(*out));
  const int n = sh->plan.n_shards, rank = sh->rank;
  int       rc = BK_OK;
  Lap       lap("shard_finish");
  if (rank > 0) {
    const bk_cut& c = sh->plan.cuts[rank - 1];
    uint64_t      s = ~0ull;
    for (int i = 0; i < rank; i++) s = std::min(s, all_reach[(size_t)i * n + rank]);
    if (!c.at_end && c.coord > 0 && s < c.coord) {
      // left halo: from the first record that starts at s to the cut.  Records in between that do not reach the cut
      // are carried along: they are map rows like any other and simply overlap nothing here.
      const uint64_t cut_off = sh->plan.map_off[rank];
      const uint64_t h0 = bk_find_start(sh->map_text, sh->plan.map_chrom_begin[rank - 1], cut_off, s);
      if (h0 < cut_off) {
        bk_bed *halo = nullptr, *both = nullptr;
        rc = load_slice(ctx, sh->map_src, h0, cut_off - h0, sh->map_fields, sh->map_cols, &halo);
        sh->bytes_in += cut_off - h0;
        if (rc == BK_OK) rc = bk_bed_concat(ctx, halo, sh->map, &both);
        bk_free_bed(ctx, halo);
        if (rc == BK_OK) {
          bk_free_bed(ctx, sh->map);
          sh->map = both;
        }
      }
    }
  }
  lap("left halo");
  if (rc == BK_OK) rc = bk_bedmap(ctx, sh->ref, sh->map, &sh->spec, out);
  lap("bedmap");
  bk_free_bed(ctx, sh->ref);  // the columns go back to the cache now; the handle itself lives until bk_shard_free
  bk_free_bed(ctx, sh->map);
  sh->ref = sh->map = nullptr;
  return rc;
}
