// hostplan.cu -- host-side planning for multi-GPU runs: chromosome byte index of a sorted BED text and
// a balanced contiguous partition of chromosomes over GPUs.  Pure host code (no kernel launches): the path shards by
// genomic range with no data-path collective, so all that is needed is WHERE to cut; every shard is then an
// ordinary single-GPU call and the outputs are concatenated in shard order.  This is the B200 counterpart of the
// reference's own scale-out, one process per chromosome with --chrom (bedmap/src/Input.hpp:117-122;
// data/bed/AllocateIterator_BED_starch.hpp:113-160 seeks by binary search over byte offsets, FindBedRange.hpp:68).
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../include/bedkit.h"

namespace {

inline bool is_ws(unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// first line start >= pos whose line is not blank; returns n if none.  `pos` need not be a line start.
size_t next_record(const char* t, size_t n, size_t pos, bool pos_is_line_start) {
  if (!pos_is_line_start) {
    const void* nl = memchr(t + pos, '\n', n - pos);
    if (!nl) return n;
    pos = (const char*)nl - t + 1;
  }
  while (pos < n) {
    size_t q = pos;
    while (q < n && is_ws((unsigned char)t[q])) q++;
    if (q < n && t[q] != '\n') return pos;
    const void* nl = memchr(t + q, '\n', n - q);
    if (!nl) return n;
    pos = (const char*)nl - t + 1;
  }
  return n;
}

// chromosome token of the record that starts at line start `pos`
void token_at(const char* t, size_t n, size_t pos, const char** tok, size_t* len) {
  while (pos < n && is_ws((unsigned char)t[pos])) pos++;
  size_t e = pos;
  while (e < n && !is_ws((unsigned char)t[e]) && t[e] != '\n') e++;
  *tok = t + pos;
  *len = e - pos;
}

bool same_tok(const char* a, size_t la, const char* b, size_t lb) { return la == lb && memcmp(a, b, la) == 0; }

}  // namespace

extern "C" int bk_chrom_index(const char* text, size_t nbytes, bk_chrom_span* out, int cap, int* n_out) {
  if (!n_out || (!text && nbytes) || (cap > 0 && !out)) return BK_ERR_ARG;
  *n_out = 0;
  // only complete lines are records
  size_t n = nbytes;
  while (n > 0 && text[n - 1] != '\n') n--;
  size_t pos = next_record(text, n, 0, true);
  int    k = 0;
  while (pos < n) {
    const char* tok;
    size_t      len;
    token_at(text, n, pos, &tok, &len);
    // gallop to a record of another chromosome, then bisect on line-aligned probes
    size_t lo = pos, hi = n, step = 1 << 16;  // lo: a record of this chromosome; hi: first byte known to be beyond the run
    while (true) {
      size_t probe = lo + step;
      if (probe >= n) break;
      size_t r = next_record(text, n, probe, false);
      if (r >= n) break;
      const char* t2;
      size_t      l2;
      token_at(text, n, r, &t2, &l2);
      if (same_tok(tok, len, t2, l2)) {
        lo = r;
        step *= 2;
      } else {
        hi = r;
        break;
      }
    }
    // invariant: record at lo belongs to the run; hi is a record of another chromosome or n
    while (true) {
      size_t after_lo = next_record(text, n, lo, false);  // the record following lo
      if (after_lo >= hi) break;
      size_t mid = lo + (hi - lo) / 2;
      size_t r = next_record(text, n, mid, false);
      if (r >= hi) r = after_lo;
      const char* t2;
      size_t      l2;
      token_at(text, n, r, &t2, &l2);
      if (same_tok(tok, len, t2, l2)) lo = r; else hi = r;
    }
    if (k < cap) {
      size_t cl = len < 127 ? len : 127;
      memcpy(out[k].name, tok, cl);
      out[k].name[cl] = 0;
      out[k].begin = pos;
      out[k].end = hi;
    }
    k++;
    pos = hi;
  }
  *n_out = k;
  return k <= cap ? BK_OK : BK_ERR_NOMEM;
}

// Partition items 0..n-1 (in order) into n_shards contiguous groups minimising the largest group load.
// first_item[s] .. first_item[s+1] are the items of shard s (first_item has n_shards + 1 entries).
extern "C" int bk_plan_shards(const uint64_t* load, int n_items, int n_shards, int* first_item) {
  if (!load || !first_item || n_items < 0 || n_shards < 1) return BK_ERR_ARG;
  uint64_t total = 0, biggest = 0;
  for (int i = 0; i < n_items; i++) {
    total += load[i];
    biggest = std::max(biggest, load[i]);
  }
  auto groups_needed = [&](uint64_t cap, int* cuts) {
    int      g = 1;
    uint64_t acc = 0;
    if (cuts) cuts[0] = 0;
    for (int i = 0; i < n_items; i++) {
      if (acc + load[i] > cap && acc > 0) {
        if (cuts && g <= n_shards) cuts[g] = i;
        g++;
        acc = 0;
      }
      acc += load[i];
    }
    return g;
  };
  uint64_t lo = biggest, hi = total;  // smallest cap that fits in n_shards groups
  while (lo < hi) {
    uint64_t mid = lo + (hi - lo) / 2;
    if (groups_needed(mid, nullptr) <= n_shards) hi = mid; else lo = mid + 1;
  }
  std::vector<int> cuts(n_shards + 2, n_items);
  int g = groups_needed(lo, cuts.data());
  for (int s = 0; s <= n_shards; s++) first_item[s] = s < g ? cuts[s] : n_items;
  first_item[n_shards] = n_items;
  return BK_OK;
}

// ---- cuts inside a chromosome (balanced genomic ranges with boundary halos) ------------------------------
namespace {

// start coordinate of the record that begins at line start `pos` (second token, strtoul-like); false if malformed
bool start_at(const char* t, size_t n, size_t pos, uint64_t* out) {
  while (pos < n && is_ws((unsigned char)t[pos])) pos++;
  while (pos < n && !is_ws((unsigned char)t[pos]) && t[pos] != '\n') pos++;  // chromosome token
  while (pos < n && is_ws((unsigned char)t[pos])) pos++;
  if (pos < n && t[pos] == '+') pos++;
  if (pos >= n || t[pos] < '0' || t[pos] > '9') return false;
  uint64_t v = 0;
  while (pos < n && t[pos] >= '0' && t[pos] <= '9') v = v * 10 + (uint64_t)(t[pos++] - '0');
  *out = v;
  return true;
}

int cmp_name(const char* a, const char* b) { return strcmp(a, b); }

}  // namespace

// First record of [begin,end) -- the lines of ONE chromosome, sorted by start -- whose start >= coord: bisection on
// line-aligned probes, O(log bytes) lines touched.  Returns `end` if there is none.
extern "C" uint64_t bk_find_start(const char* text, uint64_t begin, uint64_t end, uint64_t coord) {
  if (!text || begin >= end) return end;
  // invariant: every record that begins before lo has start < coord; the record at hi (or hi == end) has start >= coord
  uint64_t lo = next_record(text, end, begin, true), hi = end;
  if (lo >= end) return end;
  {
    uint64_t s0;
    if (!start_at(text, end, lo, &s0) || s0 >= coord) return lo;
  }
  // now: the record at lo has start < coord
  while (true) {
    const uint64_t after_lo = next_record(text, end, lo, false);
    if (after_lo >= hi) return hi;
    uint64_t mid = lo + (hi - lo) / 2;
    uint64_t r = next_record(text, end, mid, false);
    if (r >= hi) r = after_lo;
    uint64_t s;
    if (start_at(text, end, r, &s) && s < coord) lo = r; else hi = r;
  }
}

// n_shards - 1 cut points that balance the BYTES of a sorted BED text: cut k is the (chromosome, start) of the first
// record at or after byte k * nbytes / n_shards.  Cuts are non-decreasing; equal cuts make empty shards.
extern "C" int bk_plan_cuts(const char* text, size_t nbytes, const bk_chrom_span* idx, int n_idx, int n_shards, bk_cut* cuts) {
  if ((!text && nbytes) || (n_idx > 0 && !idx) || n_shards < 1 || (n_shards > 1 && !cuts)) return BK_ERR_ARG;
  const uint64_t lo_all = n_idx ? idx[0].begin : 0, hi_all = n_idx ? idx[n_idx - 1].end : 0;
  for (int k = 1; k < n_shards; k++) {
    bk_cut& c = cuts[k - 1];
    memset(&c, 0, sizeof(c));
    const uint64_t target = lo_all + (hi_all - lo_all) / (uint64_t)n_shards * (uint64_t)k;
    int j = 0;
    while (j < n_idx && idx[j].end <= target) j++;
    if (j >= n_idx) {  // beyond the last record: an empty tail shard
      c.at_end = 1;
      continue;
    }
    uint64_t pos = target <= idx[j].begin ? idx[j].begin : next_record(text, idx[j].end, target, false);
    if (pos >= idx[j].end) {  // the target fell into the chromosome's last line: cut at the next chromosome
      j++;
      if (j >= n_idx) {
        c.at_end = 1;
        continue;
      }
      pos = idx[j].begin;
    }
    memcpy(c.chrom, idx[j].name, sizeof(c.chrom));
    uint64_t s = 0;
    if (pos > idx[j].begin && start_at(text, idx[j].end, pos, &s)) c.coord = s;  // at a chromosome's first line the cut is (chrom, 0)
  }
  return BK_OK;
}

// Byte offset at which a cut falls in another sorted BED text of the same genome: the first record whose
// (chromosome, start) >= (cut.chrom, cut.coord) in sort-bed order.
extern "C" uint64_t bk_cut_offset(const char* text, size_t nbytes, const bk_chrom_span* idx, int n_idx, const bk_cut* cut) {
  size_t n = nbytes;
  while (n > 0 && text[n - 1] != '\n') n--;
  const uint64_t end_all = n_idx ? idx[n_idx - 1].end : n;
  if (!cut || cut->at_end) return end_all;
  for (int j = 0; j < n_idx; j++) {
    const int c = cmp_name(idx[j].name, cut->chrom);
    if (c < 0) continue;
    if (c > 0) return idx[j].begin;
    return cut->coord == 0 ? idx[j].begin : bk_find_start(text, idx[j].begin, idx[j].end, cut->coord);
  }
  return end_all;
}
