// check.cu -- the --ec input validation as one device pass over the raw text.
//
// Replaces Bed::bed_check_iterator<T*>::check (interfaces/general-headers/data/bed/BedCheckIterator.hpp:326-634):
// per line, the column-by-column format rules (tabs only, digits only, <= 12 digits, id / measurement syntax),
// then the ordering rules against the previous data line (chromosome by strcmp, start, end, rest; "fully nested"
// for --faster) and end > start.  Header lines (UCSC browser/track, '@', '#') are skipped as the reference does.
// One thread per line (the thread that owns the line's first byte); the FIRST offending line in file order wins
// (atomicMin over (byte offset, code, character)); the host turns the code into the reference's message text.
#include "common.cuh"
#include "parse.cuh"

namespace bk {

enum {
  EC_OK = 0, EC_EMPTY = 1, EC_CHR_SPACE, EC_CHR_TAB0, EC_NO_TABS, EC_CHR_LONG,
  EC_ST_NOCOORD, EC_ST_NEG, EC_ST_SPACE, EC_ST_NONNUM, EC_ST_NOTAB, EC_ST_DIGITS, EC_ST_MAX,
  EC_EN_NOCOORD, EC_EN_NEG, EC_EN_SPACE, EC_EN_NONNUM, EC_EN_ONLY3, EC_EN_DIGITS, EC_EN_MAX,
  EC_ID_NOID, EC_ID_SPACE, EC_ID_ONLY4, EC_ID_EMPTY, EC_ID_LONG,
  EC_SC_NONE, EC_SC_2DEC, EC_SC_DECEXP, EC_SC_2EXP, EC_SC_SPACE, EC_SC_SIGNPOS, EC_SC_2SIGN, EC_SC_EXPSIGN, EC_SC_NONNUM,
  EC_SC_EMPTY, EC_SC_ENDMINUS,
  EC_SORT_CHR = 40, EC_SORT_START, EC_SORT_END, EC_SORT_REST, EC_NESTED, EC_END_LE_START,
  EC_HEADER = 255
};

struct LineInfo {
  uint32_t chromlen;
  uint64_t start, end;
  uint32_t rest;  // offset of the end-coordinate field (restMarker, BedCheckIterator.hpp:413)
};

__device__ __forceinline__ bool lower_is(const unsigned char* t, uint64_t n, const char* w, int wl) {
  if (n != (uint64_t)wl) return false;
  for (int i = 0; i < wl; i++) {
    unsigned char c = t[i];
    if (c >= 'A' && c <= 'Z') c += 32;
    if (c != (unsigned char)w[i]) return false;
  }
  return true;
}
__device__ __forceinline__ bool is_ucsc(const unsigned char* t, uint64_t n) {  // isUCSCheader, :315-318
  return lower_is(t, n, "browser", 7) || lower_is(t, n, "track", 5);
}

// format rules of one line t[0..sz); returns code << 8 | character
__device__ uint32_t check_line(const unsigned char* t, uint64_t sz, int nf, LineInfo& li) {
  if (sz == 0) return EC_EMPTY << 8;
  if (is_ucsc(t, sz)) return EC_HEADER << 8;
  uint64_t marker = 0;
  while (marker < sz) {
    const unsigned char c = t[marker];
    if (c == ' ') {
      if (is_ucsc(t, marker)) return EC_HEADER << 8;
      return EC_CHR_SPACE << 8;
    } else if (marker == 0 && (c == '@' || c == '#')) {
      return EC_HEADER << 8;
    } else if (c == '\t') {
      if (marker == 0) return EC_CHR_TAB0 << 8;
      if (is_ucsc(t, marker)) return EC_HEADER << 8;
      break;
    }
    ++marker;
  }
  if (sz <= marker) return EC_NO_TABS << 8;
  if (marker > 127) return EC_CHR_LONG << 8;
  li.chromlen = (uint32_t)marker;
  ++marker;
  // start coordinate
  uint64_t pos = marker, v = 0;
  while (marker < sz) {
    const unsigned char c = t[marker];
    if (!is_digit(c)) {
      if (c == '\t' && pos != marker) break;
      if (c == '\t') return EC_ST_NOCOORD << 8;
      if (c == '-' && marker == pos) return EC_ST_NEG << 8;
      if (c == ' ') return EC_ST_SPACE << 8;
      return (EC_ST_NONNUM << 8) | c;
    }
    if (marker - pos < 19) v = v * 10 + (c - '0');
    ++marker;
  }
  if (sz <= marker) return EC_ST_NOTAB << 8;
  if (marker - pos > 12) return EC_ST_DIGITS << 8;
  li.start = v;
  ++marker;
  // end coordinate
  pos = marker;
  li.rest = (uint32_t)marker;
  v = 0;
  while (marker < sz) {
    const unsigned char c = t[marker];
    if (!is_digit(c)) {
      if (c == '\t' && pos != marker) break;
      if (c == '\t') return EC_EN_NOCOORD << 8;
      if (c == '-' && marker == pos) return EC_EN_NEG << 8;
      if (c == ' ') return EC_EN_SPACE << 8;
      return (EC_EN_NONNUM << 8) | c;
    }
    if (marker - pos < 19) v = v * 10 + (c - '0');
    ++marker;
  }
  if (sz <= marker && nf > 3) return EC_EN_ONLY3 << 8;
  if (marker - pos > 12) return EC_EN_DIGITS << 8;
  li.end = v;
  ++marker;
  if (nf > 3) {  // id
    pos = marker;
    while (marker < sz) {
      const unsigned char c = t[marker];
      if (c == '\t' && pos != marker) break;
      if (c == '\t') return EC_ID_NOID << 8;
      if (c == ' ') return EC_ID_SPACE << 8;
      ++marker;
    }
    if (sz <= marker && nf > 4) return EC_ID_ONLY4 << 8;
    if (pos == marker) return EC_ID_EMPTY << 8;
    if (marker - pos > 16383) return EC_ID_LONG << 8;
    ++marker;
    if (nf > 4) {  // measurement
      pos = marker;
      int      dec = 0, ex = 0, minus = 0;
      uint64_t expos = 0, minuspos = 0;
      while (marker < sz) {
        const unsigned char c = t[marker];
        if (!is_digit(c)) {
          if (c == '\t' && pos != marker) break;
          if (c == '\t') return EC_SC_NONE << 8;
          if (c == '.') {
            if (++dec > 1) return EC_SC_2DEC << 8;
            if (ex > 0) return EC_SC_DECEXP << 8;
          } else if (c == 'e' || c == 'E') {
            if (++ex > 1) return EC_SC_2EXP << 8;
            expos = marker;
          } else if (c == ' ') {
            return EC_SC_SPACE << 8;
          } else if (c == '-' || c == '+') {
            if (marker != pos && ex < 1) return EC_SC_SIGNPOS << 8;
            if (marker != pos) {
              if (++minus > 1) return EC_SC_2SIGN << 8;
              if (expos + 1 != marker) return EC_SC_EXPSIGN << 8;
              minuspos = marker;
            }
          } else {
            return (EC_SC_NONNUM << 8) | c;
          }
        }
        ++marker;
      }
      if (pos == marker) return EC_SC_EMPTY << 8;
      if (minuspos > 0 && minuspos + 1 == marker) return EC_SC_ENDMINUS << 8;
    }
  }
  return EC_OK;
}

// strcmp of two byte ranges (no NULs inside BED text)
__device__ int range_cmp(const unsigned char* a, uint64_t la, const unsigned char* b, uint64_t lb) {
  const uint64_t n = la < lb ? la : lb;
  for (uint64_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return la < lb ? -1 : (la > lb ? 1 : 0);
}

__global__ void __launch_bounds__(256) k_check_text(const unsigned char* __restrict__ text, uint64_t nbytes, int nf, int has_rest,
                                                    int nest_check, uint64_t* scratch) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 32;
  for (uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 32; p0 < nbytes; p0 += stride) {
    // line starts in my 32 bytes
    uint32_t smask = 0;
    for (int j = 0; j < 32 && p0 + j < nbytes; j++) {
      const uint64_t g = p0 + j;
      if (g == 0 || text[g - 1] == '\n') smask |= 1u << j;
    }
    for (uint32_t m = smask; m; m &= m - 1) {
      const uint64_t ls = p0 + (__ffs(m) - 1);
      uint64_t       le = ls;
      while (le < nbytes && text[le] != '\n') le++;
      if (le >= nbytes) continue;  // callers terminate the text with a NL; anything after the last NL is not a line
      LineInfo li{};
      uint32_t rc = check_line(text + ls, le - ls, nf, li);
      if ((rc >> 8) == EC_HEADER) continue;
      if (rc == EC_OK) {
        // previous DATA line (headers are skipped; a previous line with an error has a smaller offset and wins anyway)
        uint64_t pe = ls;  // one past the NL that ends the candidate previous line
        LineInfo pl{};
        bool     have_prev = false;
        uint64_t pls = 0, ple = 0;
        while (pe > 0) {
          ple = pe - 1;  // its NL
          pls = ple;
          while (pls > 0 && text[pls - 1] != '\n') pls--;
          const uint32_t prc = check_line(text + pls, ple - pls, nf, pl);
          if (prc == EC_OK) { have_prev = true; break; }
          if ((prc >> 8) != EC_HEADER) break;  // an offending line: it reports itself
          pe = pls;
        }
        if (have_prev) {
          const int cmp = range_cmp(text + ls, li.chromlen, text + pls, pl.chromlen);
          if (cmp < 0) rc = EC_SORT_CHR << 8;
          else if (cmp == 0) {
            if (li.start < pl.start) rc = EC_SORT_START << 8;
            else if (li.start == pl.start) {
              if (li.end < pl.end) rc = EC_SORT_END << 8;
              else if (has_rest && li.end == pl.end &&
                       range_cmp(text + ls + li.rest, (le - ls) - li.rest, text + pls + pl.rest, (ple - pls) - pl.rest) < 0)
                rc = EC_SORT_REST << 8;
            }
            if (rc == EC_OK && nest_check && li.end < pl.end) rc = EC_NESTED << 8;
          }
        }
        if (rc == EC_OK && li.end <= li.start) rc = EC_END_LE_START << 8;
      }
      if (rc != EC_OK) atomicMin(reinterpret_cast<unsigned long long*>(&scratch[SC_COUNT_D]), (unsigned long long)((ls << 16) | rc));
    }
  }
}

// number of NLs in text[0, upto)  (1-based line number of the offending line = this + 1)
__global__ void k_count_nl_before(const unsigned char* __restrict__ text, uint64_t upto, uint64_t* scratch) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t       c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < upto; i += stride) c += text[i] == '\n';
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(&scratch[SC_COUNT_A]), (unsigned long long)c);
}

static std::string ec_message(int code, int ch, int nf) {
  auto num = [](long v) { return std::to_string(v); };
  switch (code) {
    case EC_EMPTY: return "Empty line found.";
    case EC_CHR_SPACE:
      return "First column should not have spaces.  Consider 'chr1' vs. 'chr1 '.  These are different names.\nsort-bed can correct this for you.";
    case EC_CHR_TAB0: return "First column name should not start with a tab.";
    case EC_NO_TABS: return "No tabs found in BED row.";
    case EC_CHR_LONG:
      return "Chromosome name does not fit in MAXCHROMSIZE chars.\nIncrease TOKEN_CHR_MAX_LENGTH in BEDOPS.Constants.hpp and recompile BEDOPS.";
    case EC_ST_NOCOORD: return "Two or more consecutive tabs.  No start coordinate.";
    case EC_ST_NEG: return "Start coordinate cannot be < 0: ";
    case EC_ST_SPACE: return "Start coordinate may not contain a space: ";
    case EC_ST_NONNUM: return std::string("Start coordinate contains non-numeric character: ") + (char)ch;
    case EC_ST_NOTAB: return "No tabs after start coordinate.";
    case EC_ST_DIGITS: case EC_EN_DIGITS:
      return "Sanity check failure - start coordinate has too many digits as defined by MAX_DEC_INTEGERS in BEDOPS.Constants.hpp";
    case EC_ST_MAX: case EC_EN_MAX:
      return "Sanity check failure - start coordinate is more than allowed by MAX_COORD_VALUE in BEDOPS.Constants.hpp";
    case EC_EN_NOCOORD: return "Two or more consecutive tabs.  No end coordinate.";
    case EC_EN_NEG: return "End coordinate cannot be < 0: ";
    case EC_EN_SPACE: return "End coordinate may not contain a space: ";
    case EC_EN_NONNUM: return std::string("End coordinate contains non-numeric character: ") + (char)ch;
    case EC_EN_ONLY3: return "Only 3 columns given.  Require at least " + num(nf);
    case EC_ID_NOID: return "Two or more consecutive tabs.  No ID field.";
    case EC_ID_SPACE: return "ID field may not contain a space.";
    case EC_ID_ONLY4: return "Only 4 columns given.  Require at least " + num(nf);
    case EC_ID_EMPTY: return "Fourth (id) column is empty.";
    case EC_ID_LONG:
      return "ID field does not fit in MAXCHROMSIZE chars.\nIncrease TOKEN_ID_MAX_LENGTH in BEDOPS.Constants.hpp and recompile BEDOPS.";
    case EC_SC_NONE: return "Two or more consecutive tabs.  No measurement given.";
    case EC_SC_2DEC: return "More than one decimal point in measurement field.";
    case EC_SC_DECEXP: return "Bad decimal point - part of exponent.";
    case EC_SC_2EXP: return "Measurement value contains non-numeric character (multiple 'E' or 'e' characters detected).";
    case EC_SC_SPACE: return "Measurement value may not contain a space.";
    case EC_SC_SIGNPOS: return "Measurement value has '-' or '+' in wrong place.";
    case EC_SC_2SIGN: return "Measurement value has multiple '-' and/or '+' characters.";
    case EC_SC_EXPSIGN: return "Measurement value has bad '-' in the exponent.";
    case EC_SC_NONNUM: return std::string("Measurement value contains non-numeric character: ") + (char)ch;
    case EC_SC_EMPTY: return "Fifth (measure) column is empty.";
    case EC_SC_ENDMINUS: return "Measurement value ends with a '-'.";
    case EC_SORT_CHR: return "Bed file not properly sorted by first column.";
    case EC_SORT_START: return "Bed file not properly sorted by start coordinates.";
    case EC_SORT_END: return "Bed file not properly sorted by end coordinates when start coordinates are identical.";
    case EC_SORT_REST: return "Bed file not sorted by information following the 3rd column (columns 1-3 equal to previous row).";
    case EC_NESTED: return "Fully nested component found.";
    case EC_END_LE_START: return "End coordinates must be greater than start coordinates.";
  }
  return "Unknown input error.";
}

}  // namespace bk

using namespace bk;

// Validate device-resident text.  On failure returns BK_ERR_CHECK and bk_last_error() holds
// "<message>\nSee row: <line>" (the tool prefixes "in <file>\n" exactly like BedCheckIterator.hpp:589-593).
extern "C" int bk_check_text_device(bk_ctx* ctx, const char* dev_text, size_t nbytes, int n_fields, int has_rest, int nest_check) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || (!dev_text && nbytes)) return BK_ERR_ARG;
  ctx->last_error.clear();
  if (nbytes == 0) return BK_OK;
  const unsigned char* t = reinterpret_cast<const unsigned char*>(dev_text);
  BK_TRY(reset_scratch(ctx));
  const uint64_t big = ~0ull;
  BK_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch + SC_COUNT_D, &big, 8, cudaMemcpyHostToDevice, ctx->stream));
  BK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t blocks = (nbytes / 32 + 255) / 256 + 1, cap = (uint64_t)ctx->sms * 16;
  prof_begin(ctx, "k_check_text");
  k_check_text<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(t, nbytes, n_fields, has_rest, nest_check, ctx->d_scratch);
  prof_end(ctx);
  BK_LAUNCHED(ctx);
  BK_TRY(read_scratch(ctx));
  const uint64_t key = ctx->h_scratch[SC_COUNT_D];
  if (key == big) return BK_OK;
  const uint64_t off = key >> 16;
  const int      code = (int)((key >> 8) & 0xFF), ch = (int)(key & 0xFF);
  BK_TRY(reset_scratch(ctx));
  if (off) {
    uint64_t b = (off + 255) / 256, c2 = (uint64_t)ctx->sms * 16;
    k_count_nl_before<<<(unsigned)(b < c2 ? b : c2), 256, 0, ctx->stream>>>(t, off, ctx->d_scratch);
    BK_LAUNCHED(ctx);
  }
  BK_TRY(read_scratch(ctx));
  const uint64_t line = ctx->h_scratch[SC_COUNT_A] + 1;
  return fail(ctx, BK_ERR_CHECK, "%s\nSee row: %llu", ec_message(code, ch, n_fields).c_str(), (unsigned long long)line);
}

extern "C" int bk_check_text(bk_ctx* ctx, const char* host_text, size_t nbytes, int n_fields, int has_rest, int nest_check) {
  bk::DeviceGuard device_guard(ctx);
  if (!ctx || (!host_text && nbytes)) return BK_ERR_ARG;
  ctx->last_error.clear();
  if (nbytes == 0) return BK_OK;
  char* d = reinterpret_cast<char*>(dmalloc(ctx, nbytes + 64));
  if (!d) return BK_ERR_NOMEM;
  BK_CUDA(ctx, cudaMemcpyAsync(d, host_text, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = bk_check_text_device(ctx, d, nbytes, n_fields, has_rest, nest_check);
  dfree(ctx, d);
  return rc;
}
