// parse.cuh -- shared declarations of the BED reader kernels
#pragma once
#include "common.cuh"

namespace bk {

constexpr int P_THREADS = 256;
#ifndef BK_P_MINBLOCKS
#define BK_P_MINBLOCKS 6
#endif
constexpr int P_MINBLOCKS = BK_P_MINBLOCKS;  // resident CTAs per SM the compiler must leave registers for
constexpr int P_TILE = P_THREADS * 32;  // bytes of text whose line STARTS one tile owns
constexpr int P_PRE = 128;    // halo before the tile (previous line's chromosome token)
constexpr int P_POST = 384;   // halo after the tile (tail of the last line that starts in the tile)
constexpr int P_BUF = P_PRE + P_TILE + P_POST;
constexpr int P_MAXROWS = P_TILE / 2;

struct HeadRec {  // first row of a chromosome run, discovered by the parser
  uint64_t row;
  uint32_t len;
  char     name[132];
};

struct ParseParams {
  const char* text;
  uint64_t    nbytes_raw;
  int         min_fields;
  unsigned    cols;
  uint32_t*   start;
  uint32_t*   end;
  double*     score;
  uint64_t*   line_off;
  uint32_t*   idspan;
  uint64_t    cap;
  const uint64_t* tile_base;     // [ntiles] first row of every tile (pass 1: warp-range base + prefix inside the range)
  uint32_t    ntiles;
  uint32_t    bulk_tiles;        // tiles 1..bulk_tiles have their whole window inside the text (staged by cp.async.bulk)
  const uint32_t* cm_arr;        // pass 1: packed control-byte mask, one word per 32 text bytes
  const uint32_t* ls_arr;        // pass 1: packed line-start mask (rows only: blank and header lines already dropped)
  uint64_t    nwords;            // words in each of the two arrays
  uint64_t*   scratch;
  HeadRec*    heads;
  uint32_t    heads_cap;
};

int reset_scratch(bk_ctx* ctx);
int read_scratch(bk_ctx* ctx);
int parse_bed(bk_ctx* ctx, bk_bed* bed, uint64_t nbytes_raw);
int ensure_pmax(bk_ctx* ctx, const bk_bed* bed);
int ensure_bmax(bk_ctx* ctx, const bk_bed* bed);  // max end per 32-row block (dense-map variant of the window scan)
int seg_prefix_max(bk_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n, const std::vector<ChromRun>& runs);

}  // namespace bk
