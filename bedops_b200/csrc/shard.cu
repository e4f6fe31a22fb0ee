// shard.cu -- the device half of genomic-range sharding with boundary halos (protocol in include/bedkit.h):
// where a shard's left halo must begin (prefix-max-end index), how far its right halo reaches (largest reference end),
// and the concatenation halo ++ own records.  The host half (cuts, byte offsets) is hostplan.cu.
#include <algorithm>
#include "common.cuh"
#include "parse.cuh"

namespace bk {

// first row of [0,n) whose running-max end exceeds pos -> its start (sorted by start: the smallest start of any row that
// ends beyond pos); one warp
__global__ void k_reach_start(const uint32_t* __restrict__ pm, const uint32_t* __restrict__ start, uint32_t n, uint64_t pos,
                              uint64_t* __restrict__ scratch) {
  const int      lane = threadIdx.x;
  const uint32_t key = pos >= 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)pos + 1u;
  const uint32_t k = pos >= 0xFFFFFFFFull ? n : warp_search32(pm, n, key, lane);
  if (lane == 0) scratch[SC_COUNT_A] = k < n ? (uint64_t)start[k] : ~0ull;
}

__global__ void k_shift_line_off(uint64_t* __restrict__ line_off, uint64_t n, uint64_t delta) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) line_off[k] += delta;  // low 48 bits: the offset
}

static const ChromRun* run_named(const bk_bed* b, const char* name) {
  for (auto& r : b->runs)
    if (r.name == name) return &r;
  return nullptr;
}

}  // namespace bk

using namespace bk;

extern "C" int bk_bed_reach_start(bk_ctx* ctx, const bk_bed* bed, const char* chrom, uint64_t pos, uint64_t* start_out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !bed || !chrom || !start_out) return BK_ERR_ARG;
  *start_out = ~0ull;
  const ChromRun* r = run_named(bed, chrom);
  if (!r || r->row_end == r->row_begin) return BK_OK;
  BK_TRY(ensure_pmax(ctx, bed));
  k_reach_start<<<1, 32, 0, ctx->stream>>>(bed->pmax_end + r->row_begin, bed->start + r->row_begin,
                                           (uint32_t)(r->row_end - r->row_begin), pos, ctx->d_scratch);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));
  *start_out = ctx->h_scratch[SC_COUNT_A];
  return BK_OK;
}

extern "C" int bk_bed_chrom_max_end(bk_ctx* ctx, const bk_bed* bed, const char* chrom, uint64_t* end_out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !bed || !chrom || !end_out) return BK_ERR_ARG;
  *end_out = 0;
  const ChromRun* r = run_named(bed, chrom);
  if (!r || r->row_end == r->row_begin) return BK_OK;
  BK_TRY(ensure_pmax(ctx, bed));
  uint32_t v = 0;
  BK_CUDA(ctx, cudaMemcpyAsync(&v, bed->pmax_end + r->row_end - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *end_out = v;
  return BK_OK;
}

extern "C" int bk_bed_concat(bk_ctx* ctx, const bk_bed* head, const bk_bed* tail, bk_bed** out) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || !head || !tail || !out) return BK_ERR_ARG;
  *out = nullptr;
  ctx->last_error.clear();
  if (head->min_fields != tail->min_fields || head->cols != tail->cols)
    return fail(ctx, BK_ERR_ARG, "bk_bed_concat: the two parts were loaded with different record types / columns");
  if (!head->runs.empty() && !tail->runs.empty()) {
    const int c = strcmp(head->runs.back().name.c_str(), tail->runs.front().name.c_str());
    if (c > 0) return fail(ctx, BK_ERR_UNSORTED, "bk_bed_concat: chromosome '%s' follows '%s'", tail->runs.front().name.c_str(),
                           head->runs.back().name.c_str());
  }
  bk_bed* b = new bk_bed();
  b->min_fields = head->min_fields;
  b->cols = head->cols;
  const uint64_t na = head->nrows, nb = tail->nrows, n = na + nb;
  b->nrows = n;
  auto fail_mem = [&]() {
    bk_free_bed(ctx, b);
    return BK_ERR_NOMEM;
  };
  auto cat = [&](auto*& dst, const auto* a, const auto* t, uint64_t extra) -> bool {
    using T = typename std::remove_reference<decltype(*dst)>::type;
    dst = dalloc<T>(ctx, n + extra);
    if (!dst) return false;
    if (na && cudaMemcpyAsync(dst, a, na * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) return false;
    if (nb && cudaMemcpyAsync(dst + na, t, nb * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) return false;
    return true;
  };
  if (!cat(b->start, head->start, tail->start, 0) || !cat(b->end, head->end, tail->end, 0)) return fail_mem();
  if (head->score || tail->score) {
    if ((na && !head->score) || (nb && !tail->score) || !cat(b->score, head->score, tail->score, 0)) return fail_mem();
  }
  if (head->idspan || tail->idspan) {
    if ((na && !head->idspan) || (nb && !tail->idspan) || !cat(b->idspan, head->idspan, tail->idspan, 0)) return fail_mem();
  }
  if (head->line_off || tail->line_off) {  // lines are kept: one text buffer, the tail's offsets move behind the head's text
    if ((na && !head->line_off) || (nb && !tail->line_off) || !cat(b->line_off, head->line_off, tail->line_off, 1)) return fail_mem();
    const uint64_t ha = (head->nbytes + 15) & ~15ull;  // keep the tail's text 16-byte aligned
    char* text = reinterpret_cast<char*>(dmalloc(ctx, ha + tail->nbytes + 64));
    if (!text) return fail_mem();
    b->d_text = text;
    b->owns_text = true;
    b->nbytes = ha + tail->nbytes;
    if (head->nbytes) BK_CUDA(ctx, cudaMemcpyAsync(text, head->d_text, head->nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (ha > head->nbytes) BK_CUDA(ctx, cudaMemsetAsync(text + head->nbytes, '\n', ha - head->nbytes, ctx->stream));
    if (tail->nbytes) BK_CUDA(ctx, cudaMemcpyAsync(text + ha, tail->d_text, tail->nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (nb) {
      const uint64_t blocks = std::min<uint64_t>((nb + 255) / 256, (uint64_t)ctx->sms * 16);
      k_shift_line_off<<<(unsigned)blocks, 256, 0, ctx->stream>>>(b->line_off + na, nb, ha);
      BK_LAUNCHED(ctx);
    }
    const uint64_t endoff = b->nbytes;
    BK_CUDA(ctx, cudaMemcpyAsync(b->line_off + n, &endoff, 8, cudaMemcpyHostToDevice, ctx->stream));
    BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  } else {
    b->nbytes = head->nbytes + tail->nbytes;
  }
  b->runs = head->runs;
  for (const ChromRun& r : tail->runs) {
    if (!b->runs.empty() && b->runs.back().name == r.name) b->runs.back().row_end = na + r.row_end;
    else b->runs.push_back({r.name, na + r.row_begin, na + r.row_end});
  }
  *out = b;
  return BK_OK;
}
