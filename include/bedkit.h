/* bedkit.h -- C ABI of the B200-native sorted-interval engine (libbedkit.so).
 *
 * The reference (BEDOPS v2.4.26) has no library boundary: its algorithms are templates compiled into each
 * tool.  The three seams this ABI replaces are the calls a maintainer would swap out in the reference:
 *
 *   bk_load_bed / bk_load_bed_device   replaces  Bed::allocate_iterator_starch_bed<T*> + T::readline(FILE*)
 *                                      (interfaces/general-headers/data/bed/AllocateIterator_BED_starch.hpp:60-187,
 *                                       data/bed/Bed.hpp:244-255, 343-360, 577-606, 829-860)
 *   bk_bedmap                          replaces  WindowSweep::sweep(refI, refEnd, mapI, mapEnd, st, multiv, sweepAll)
 *                                      (applications/bed/bedmap/src/Bedmap.cpp:288, single-file form :229) together
 *                                      with the MultiVisitor print chain (visitors/other/MultiVisitor.hpp:83-98)
 *   bk_setop                           replaces  selectWork(...) -> doMerge/doIntersection/doElementOf
 *                                      (applications/bed/bedops/src/Bedops.cpp:1523-1577, :538-606)
 *   bk_closest                         replaces  findDistances(ref, nonRef, allowOverlaps, printer)
 *                                      (applications/bed/closestfeats/src/ClosestFeature.cpp:260-413)
 *
 * Conventions: plain pointers and sizes, integer return codes (0 = ok), never exceptions; the caller owns input
 * text; the library owns device memory and returns result text as bk_text, released with bk_free_text.  One host
 * thread per bk_ctx (one ctx per GPU; contexts are independent).  There is no CPU fallback: every entry point
 * fails with BK_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef BEDKIT_H
#define BEDKIT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEDKIT_ABI_VERSION 3

typedef struct bk_ctx bk_ctx; /* one per GPU: stream, cached device blocks, pinned staging, last error text */
typedef struct bk_bed bk_bed; /* a parsed, device-resident sorted BED file: SoA columns + chromosome runs */

/* result text; ptr is pinned host memory (on_device == 0) or device memory (on_device == 1) */
typedef struct bk_text {
  char*    ptr;
  uint64_t len;
  int      on_device;
  uint64_t rows; /* number of output lines */
} bk_text;

/* ---- error codes ----------------------------------------------------------------------------------------- */
enum {
  BK_OK = 0,
  BK_ERR_CUDA = 1,        /* CUDA runtime failure or no usable device */
  BK_ERR_NOMEM = 2,
  BK_ERR_ARG = 3,         /* bad argument to the ABI */
  BK_ERR_PARSE = 4,       /* a BED line the device parser does not accept (message names the row) */
  BK_ERR_COORD_RANGE = 5, /* coordinate >= 2^32 - 1: outside this build's 32-bit device layout */
  BK_ERR_UNSUPPORTED = 6, /* option combination outside the hot path (named in the message) */
  BK_ERR_STARCH = 7,      /* a Starch archive where plain text is required (bk_load_bed), or one that is not read: v1, --header, damaged */
  BK_ERR_UNSORTED = 8,    /* chromosome runs not in strcmp order / repeated chromosome run */
  BK_ERR_CHECK = 9,       /* --ec validation failed (message mirrors BedCheckIterator.hpp:589-593) */
  BK_ERR_NAN_ELEMENT = 10 /* bedmap --max-element/--min-element reached a reference row with no mapped element: the result text
                             holds what the reference had printed when it threw (ProcessBedVisitorRow.hpp:206-208); print it, then fail */
};

/* ---- context --------------------------------------------------------------------------------------------- */
int         bk_init(bk_ctx** out, int device);
void        bk_destroy(bk_ctx* ctx);
/* run all work of this ctx on an existing CUDA stream (cudaStream_t); NULL restores the ctx's own stream */
int         bk_set_stream(bk_ctx* ctx, void* cuda_stream);
int         bk_sync(bk_ctx* ctx);
const char* bk_strerror(int code);
const char* bk_last_error(const bk_ctx* ctx); /* detail for the last failing call on this ctx ("" if none) */
int         bk_abi_version(void);
/* number of kernels this ctx has launched since creation (bench.py reports it as gpu_launches) */
uint64_t    bk_launch_count(const bk_ctx* ctx);

/* per-kernel device timing for bench.py's roofline: enable (and reset) recording of a CUDA event pair around every
 * kernel launch; query sums the elapsed time of the launches named `kernel` ("k_parse", "k_map_stats", "k_emit", ...)
 * since the last reset.  Off by default; costs two event records per launch when on. */
int bk_profile(bk_ctx* ctx, int enable);
int bk_profile_query(bk_ctx* ctx, const char* kernel, double* total_ms, uint64_t* launches);
/* cudaMemcpyAsync(cudaMemcpyDefault) + sync on the ctx stream (bench/test plumbing for device-resident text) */
int bk_copy(bk_ctx* ctx, void* dst, const void* src, size_t nbytes);
/* Device memory is cached per ctx: blocks freed by bk_free_bed / bk_free_text / internal temporaries are kept (by size
 * class) and reused by later calls without a driver call.  bk_release_cached hands every idle block back to the driver
 * (bk_destroy does it too); the library does it by itself when an allocation fails or when the idle blocks exceed a
 * quarter of the device memory. */
int bk_release_cached(bk_ctx* ctx);

/* ---- BED reader ------------------------------------------------------------------------------ */
/* which per-row columns the parser materialises besides start/end */
enum {
  BK_COL_LINE = 1,  /* byte offset of each line (needed to echo a row or copy its id) */
  BK_COL_SCORE = 2, /* column 5 as double (strtod-exact); requires min_fields == 5 */
  BK_COL_ID = 4,    /* (offset,len) of column 4; requires min_fields >= 4 and BK_COL_LINE */
  BK_LOAD_HEADERS = 8, /* --ec/--header: UCSC browser/track lines and lines starting with '@' or '#' are not records
                         (BedCheckIterator.hpp:315-360) */
  BK_LOAD_SORTBED = 16 /* sort-bed's reader (SortDetails.cpp:617-629): rows in any order, only empty lines are skipped, rows
                          the tokeniser does not take are kept for bk_sort_bed's own validation; used by bk_sort_bed only */
};
/* min_fields = 3|4|5 selects the reference record type B3Rest / B4Rest / B5Rest (bedmap/src/Bedmap.cpp:623-654).
 * host text: copied to the device (pinned staging, async), then parsed.  The text must stay valid until return. */
int bk_load_bed(bk_ctx* ctx, const char* host_text, size_t nbytes, int min_fields, unsigned cols, bk_bed** out);
/* device text: borrowed, 16-byte aligned, must outlive the bk_bed */
int bk_load_bed_device(bk_ctx* ctx, const char* dev_text, size_t nbytes, int min_fields, unsigned cols, bk_bed** out);
void        bk_free_bed(bk_ctx* ctx, bk_bed* bed);
uint64_t    bk_bed_rows(const bk_bed* bed);
int         bk_bed_nchrom(const bk_bed* bed);
const char* bk_bed_chrom_name(const bk_bed* bed, int i);
uint64_t    bk_bed_chrom_rows(const bk_bed* bed, int i);
/* test/debug accessor: copy parsed columns to host arrays of bk_bed_rows() elements (any pointer may be NULL) */
int bk_bed_copy_columns(bk_ctx* ctx, const bk_bed* bed, uint32_t* start, uint32_t* end, double* score,
                        uint64_t* line_off);

/* ---- --ec input validation -------------------------------------------------------------------------- */
/* Replaces Bed::bed_check_iterator<T*>::check (data/bed/BedCheckIterator.hpp:326-634): format rules per line, then
 * sort order against the previous data line, "fully nested" (nest_check, bedmap --faster) and end > start.
 * n_fields = NumFields of the reference record type (3|4|5), has_rest = its UseRest.  The text must end with '\n'
 * (the checking iterator also reads an unterminated last line; callers append the '\n').  On failure returns
 * BK_ERR_CHECK and bk_last_error() = "<message>\nSee row: <n>"; the caller prefixes "in <file>\n" (:589-593). */
int bk_check_text(bk_ctx* ctx, const char* host_text, size_t nbytes, int n_fields, int has_rest, int nest_check);
int bk_check_text_device(bk_ctx* ctx, const char* dev_text, size_t nbytes, int n_fields, int has_rest, int nest_check);

/* ---- bedmap ------------------------------------------------------------------------------ */
enum { /* operations, printed left to right in the order given (MultiVisitor.hpp:83-98) */
  BK_OP_ECHO = 1,        /* --echo          EchoVisitor.hpp:39-65 */
  BK_OP_COUNT = 2,       /* --count         CountVisitor.hpp:34-64 */
  BK_OP_INDICATOR = 3,   /* --indicator     IndicatorVisitor.hpp:37-56 */
  BK_OP_BASES = 4,       /* --bases         OvrAggregateVisitor.hpp:41-108 */
  BK_OP_SUM = 5,         /* --sum           SumVisitor.hpp:36-68 */
  BK_OP_MEAN = 6,        /* --mean          AverageVisitor.hpp:35-76 */
  BK_OP_MAX = 7,         /* --max           ExtremeVisitor.hpp:84-134 */
  BK_OP_MIN = 8,         /* --min */
  BK_OP_ECHO_MAP_ID = 9, /* --echo-map-id   EchoMapBedVisitor.hpp:39-66 */
  BK_OP_ECHO_REF_SIZE = 10,  /* --echo-ref-size  */
  BK_OP_ECHO_REF_NAME = 11,  /* --echo-ref-name  chrom:start-end */
  BK_OP_ECHO_REF_ROW_ID = 12, /* --echo-ref-row-id  id-<row> */
  /* lists over the qualifying map rows, joined by --multidelim (EchoMapBedVisitor.hpp:39-66, helpers :140-170) */
  BK_OP_ECHO_MAP = 13,          /* --echo-map           the rows as their record type prints them (map needs BK_COL_LINE) */
  BK_OP_ECHO_MAP_SCORE = 14,    /* --echo-map-score     scores with --prec/--sci (map needs BK_COL_SCORE) */
  BK_OP_ECHO_MAP_SIZE = 15,     /* --echo-map-size      end - start */
  BK_OP_ECHO_OVERLAP_SIZE = 16, /* --echo-overlap-size  EchoMapIntersectLengthVisitor.hpp:64-73 */
  BK_OP_ECHO_MAP_RANGE = 17,    /* --echo-map-range     one chrom<TAB>min start<TAB>max end (ProcessBedVisitorRow.hpp:433-456) */
  BK_OP_BASES_UNIQ = 18,        /* --bases-uniq         OvrUniqueVisitor.hpp:62-77 */
  BK_OP_BASES_UNIQ_F = 19,      /* --bases-uniq-f       OvrUniqueFractionVisitor.hpp:47-50 */
  /* (n*sum(x^2) - sum(x)^2) / (n*(n-1)); NAN for fewer than two hits (map needs BK_COL_SCORE) */
  BK_OP_VARIANCE = 20,          /* --variance           VarianceVisitor.hpp:58-66 */
  BK_OP_STDEV = 21,             /* --stdev              StdevVisitor.hpp */
  BK_OP_CV = 22,                /* --cv                 CoeffVariationVisitor.hpp (NAN when the mean is 0) */
  BK_OP_ECHO_MAP_ID_UNIQ = 23,  /* --echo-map-id-uniq   distinct ids in strcmp order (ProcessBedVisitorRow.hpp:361-389); BK_COL_ID|BK_COL_LINE */
  /* generalised median of the scores (RollingKthAverageVisitor.hpp:37-68 over RollingKthVisitor.hpp:75-95) */
  BK_OP_MEDIAN = 24,            /* --median             = --kth 0.5 */
  BK_OP_KTH = 25,               /* --kth <val>          0 < val < 1 in bk_mapspec.op_arg[] of the same slot */
  BK_OP_MAD = 26,               /* --mad [mult]         median absolute deviation * mult (op_arg[], 0 = default 1); MedianAbsoluteDeviationVisitor.hpp:57-93 */
  BK_OP_WMEAN = 27,             /* --wmean              scores weighted by overlap / reference length (WeightedAverageVisitor.hpp:55-70) */
  BK_OP_TMEAN = 28,             /* --tmean <low> <hi>   mean after trimming the fractions op_arg[] / op_arg2[] of the sorted scores (TrimmedMeanVisitor.hpp:93-141) */
  BK_OP_MAX_ELEMENT = 29,       /* --max-element        the highest-scoring map row, ties to the genomically last (ExtremeVisitor.hpp:84-134,
                                                        BedCompare.hpp:263-288), printed chrom start end id score rest (ProcessBedVisitorRow.hpp:181-222) */
  BK_OP_MIN_ELEMENT = 30        /* --min-element        the lowest-scoring map row, ties to the genomically first.  Both abort the run at the
                                                        first reference row without a mapped element (BK_ERR_NAN_ELEMENT), as the reference does */
};
enum { /* overlap criterion (Bedmap.cpp:107-156; BedDistances.hpp:41-317) */
  BK_OVR_BP = 0,          /* --bp-ovr N (default N = 1) */
  BK_OVR_RANGE = 1,       /* --range N */
  BK_OVR_FRAC_REF = 2,    /* --fraction-ref F */
  BK_OVR_FRAC_MAP = 3,    /* --fraction-map F */
  BK_OVR_FRAC_EITHER = 4, /* --fraction-either F */
  BK_OVR_FRAC_BOTH = 5,   /* --fraction-both F */
  BK_OVR_EXACT = 6        /* --exact */
};
#define BK_MAX_OPS 64
typedef struct bk_mapspec {
  int         n_ops;
  int         ops[BK_MAX_OPS];
  int         overlap_kind;
  uint64_t    overlap_bp;    /* BK_OVR_BP: required bases; BK_OVR_RANGE: padding */
  double      overlap_frac;  /* BK_OVR_FRAC_* */
  int         precision;     /* --prec, default 6 */
  int         sci;           /* --sci */
  int         skip_unmapped; /* --skip-unmapped */
  const char* delim;         /* --delim, default "|" (already unescaped) */
  const char* multidelim;    /* --multidelim, default ";" */
  const char* chrom;         /* --chrom, NULL or "all" = every chromosome */
  int         out_on_device; /* leave result text in HBM (bench: device-resident timing) */
  double      op_arg[BK_MAX_OPS]; /* per-operation argument (BK_OP_KTH: the fraction, BK_OP_MAD: the multiplier); 0 otherwise */
  uint64_t    row_id_base;   /* --echo-ref-row-id: rows already printed by earlier calls of the same command (the reference
                                counts printed rows, ProcessBedVisitorRow.hpp:347-354); 0 for a whole-file call */
  double      op_arg2[BK_MAX_OPS]; /* second per-operation argument (BK_OP_TMEAN: <hi>); 0 otherwise */
} bk_mapspec;
void bk_mapspec_default(bk_mapspec* spec);
/* map == NULL: single-file mode, ref is mapped onto itself (Input.hpp:359-364) */
int bk_bedmap(bk_ctx* ctx, const bk_bed* ref, const bk_bed* map, const bk_mapspec* spec, bk_text* out);

/* The whole bedmap call over HOST text (what the tool does: read two files, map, print), cut into chromosome groups so
 * that the host->device copies of later groups and the device->host copies of earlier results overlap the kernels
 * (PCIe is full duplex; pin the inputs for the overlap to happen).  Same output as
 * bk_load_bed x2 + bk_bedmap; falls back to exactly that for --chrom, single-chromosome or unsorted input, and to
 * produce error messages.  ref_fields/ref_cols/map_fields/map_cols as for bk_load_bed.  Result: host text. */
int bk_bedmap_host(bk_ctx* ctx, const char* ref_text, size_t ref_len, int ref_fields, unsigned ref_cols,
                   const char* map_text, size_t map_len, int map_fields, unsigned map_cols, const bk_mapspec* spec,
                   bk_text* out);

/* ---- bedops set operators ------------------------------------------------------------------- */
enum {
  BK_SETOP_MERGE = 1,
  BK_SETOP_INTERSECT = 2,
  BK_SETOP_ELEMENT_OF = 3,
  BK_SETOP_NOT_ELEMENT_OF = 4,
  BK_SETOP_COMPLEMENT = 5, /* --complement; thr != 0 selects -L (Bedops.cpp:475-488, :891-943) */
  BK_SETOP_DIFFERENCE = 6, /* --difference: files[0] minus the others (:501-525, :948-1018) */
  BK_SETOP_SYMMDIFF = 7,   /* --symmdiff: bases covered by exactly one file (:698-745, :1341-1463) */
  BK_SETOP_EVERYTHING = 8, /* --everything: all rows of all files in sort-bed order, every file needs BK_COL_LINE (:761-786, :1468-1516) */
  BK_SETOP_PARTITION = 9   /* --partition: the covered pieces between consecutive break points of all files (:615-653, :1249-1335) */
};
/* thr / thr_is_pct: -e/-n threshold as Input::Threshold()/UsePercentage() deliver it (bedops/src/Input.hpp:344-382):
 * a fraction in (0,1] when thr_is_pct, else a base count.  files[0] is the reference file for -e/-n. */
int bk_setop(bk_ctx* ctx, int op, const bk_bed* const* files, int n_files, double thr, int thr_is_pct,
             const char* chrom, int out_on_device, bk_text* out);

/* --chop: the merged union of the files cut into pieces of `chunk` bases starting every `stagger` bases (0 = every
 * chunk bases); a piece passing its segment's end is clipped, or with exclude_short (-x) ends the segment's pieces
 * (doChop, Bedops.cpp:438-467; option grammar Input.hpp:221-258). */
int bk_chop(bk_ctx* ctx, const bk_bed* const* files, int n_files, uint64_t chunk, uint64_t stagger, int exclude_short,
            const char* chrom, int out_on_device, bk_text* out);

/* --range L:R: the padded view of a parsed file as the set operators read it (BedPadReader.hpp:116-277): start += lpad,
 * end += rpad; rows that stop being intervals vaporise; with lpad < 0 the starts that would pass zero are clamped to 0 and
 * those rows re-ordered by their end (input order on ties).  The result borrows the text of `src` (which must outlive it)
 * and owns its columns.  BK_ERR_COORD_RANGE when a padded coordinate leaves the 32-bit layout (the reference wraps an
 * end below |rpad| to 2^64 - x there). */
int bk_bed_pad(bk_ctx* ctx, const bk_bed* src, long long lpad, long long rpad, bk_bed** out);

/* ---- closest-features ----------------------------------------------------------------------- */
typedef struct bk_cfspec {
  int         dist;        /* --dist */
  int         closest;     /* --closest */
  int         no_overlaps; /* --no-overlaps */
  int         no_ref;      /* --no-ref */
  int         no_query;    /* --no-query */
  int         center;      /* --center */
  const char* delim;       /* --delim, default "|" */
  const char* chrom;       /* --chrom */
  int         out_on_device;
} bk_cfspec;
void bk_cfspec_default(bk_cfspec* spec);
int  bk_closest(bk_ctx* ctx, const bk_bed* ref, const bk_bed* query, const bk_cfspec* spec, bk_text* out);

/* ---- Starch v2 archives as input ---------------------------------------------------------------- */
/* Replaces the reading side of interfaces/src/data/starch/unstarchHelpers.c (UNSTARCH_extractDataWithBzip2 :265-355,
 * UNSTARCH_extractDataWithGzip :57-263, UNSTARCH_reverseTransformHeaderlessInput :1161-1238) behind the archive
 * detection of allocate_iterator_starch_bed (AllocateIterator_BED_starch.hpp:100-112).  The host inflates the
 * per-chromosome bzip2 / gzip streams (libbz2.so.1.0 at run time, zlib), the device undoes the delta coding (two scans)
 * and writes plain BED text: what `unstarch archive` prints, byte for byte.  chrom: NULL / "all" or one chromosome
 * (`unstarch chrN archive`).  Not read (BK_ERR_STARCH with a message): v1 archives, archives made with --header. */
int bk_is_starch(const char* bytes, size_t nbytes); /* magic ca 5c ad e5 */
int bk_unstarch(bk_ctx* ctx, const char* archive, size_t nbytes, const char* chrom, int out_on_device, bk_text* out);
/* host stage alone (no device needed; tests): the inflated, still delta-coded streams as ">chrom\n<stream>" blocks in
 * archive order, in a malloc'ed buffer the caller releases with bk_host_free */
int  bk_starch_inflate_host(const char* archive, size_t nbytes, const char* chrom, char** text, size_t* len);
void bk_host_free(void* p);

/* ---- sort-bed ------------------------------------------------------------------------------ */
/* Replaces processData / lexSortBedData / printBed (applications/bed/sort-bed/src/SortDetails.cpp:530-1208): BED rows in
 * any order -> rows ordered by chromosome (strcmp), start, end, rest of the line (strcmp; a row without a rest first,
 * Structures.hpp:47-76), printed "chrom\tstart\tend[\trest]\n".  The text is the concatenation of the input files
 * with each file's leading header lines removed (the tool does that, SortDetails.cpp:645-653) and must end with '\n'.
 * A row sort-bed rejects (SortDetails.cpp:638-779, :833-853) gives BK_ERR_PARSE, a coordinate >= 2^32-1
 * BK_ERR_COORD_RANGE; *bad_offset (may be NULL) receives the byte offset of that row's line so that the caller can
 * name the file, the line and the reference's message for it. */
int bk_sort_bed(bk_ctx* ctx, const char* host_text, size_t nbytes, int out_on_device, bk_text* out, uint64_t* bad_offset);
int bk_sort_bed_device(bk_ctx* ctx, const char* dev_text, size_t nbytes, int out_on_device, bk_text* out,
                       uint64_t* bad_offset);
/* the sorter's engine by itself: stable least-significant-digit radix sort of n (u64 key, u32 value) pairs in device
 * memory by key bits [0, 8 * ceil(nbits / 8)) -- whole 8-bit passes --, in place from the caller's point of view (d_vals may be NULL) */
int bk_radix_sort_pairs(bk_ctx* ctx, uint64_t* d_keys, uint32_t* d_vals, uint64_t n, int nbits);

/* ---- BED writer used by the synthetic-input generator and tests -------------------------------------------- */
/* format n rows "chrom\tstart\tend[\tid<k>\tscore]\n" from device SoA arrays (one chromosome name per call);
 * id_base < 0 writes BED3.  Result text in HBM (out->on_device = 1). */
int bk_format_bed_device(bk_ctx* ctx, const char* chrom, const uint32_t* d_start, const uint32_t* d_end,
                         const uint32_t* d_score, uint64_t n, int64_t id_base, bk_text* out);

void bk_free_text(bk_ctx* ctx, bk_text* text);

/* ---- multi-GPU planning (host only, no device needed) ------------------------------------------------ */
/* The path shards by genomic range with no data-path collective: cut the sorted inputs at chromosome boundaries,
 * give every GPU (one bk_ctx each) a contiguous group of chromosomes, concatenate the outputs in shard order.
 * Replaces the reference's per-chromosome seek (AllocateIterator_BED_starch.hpp:113-160 -> FindBedRange.hpp:68). */
typedef struct bk_chrom_span {
  char     name[128];
  uint64_t begin; /* byte offset of the chromosome's first line */
  uint64_t end;   /* byte offset one past its last line */
} bk_chrom_span;
/* chromosome byte ranges of a sorted BED text by galloping + bisection on line-aligned probes (O(#chrom log n)).
 * Returns BK_ERR_NOMEM (and the needed count in *n_out) when cap is too small. */
int bk_chrom_index(const char* host_text, size_t nbytes, bk_chrom_span* out, int cap, int* n_out);
/* contiguous partition of n_items loads into n_shards groups minimising the largest group;
 * first_item[0..n_shards] receives the group boundaries */
int bk_plan_shards(const uint64_t* load, int n_items, int n_shards, int* first_item);


/* ---- cuts INSIDE chromosomes: one dataset over N GPUs by balanced genomic ranges with boundary halos --------------
 * A cut is a genomic position (chromosome, start coordinate); shard k owns the records whose (chromosome, start) lies in
 * [cut k-1, cut k).  Reference rows go to the shard that owns their start.  A shard's map rows are its own records plus
 *   a right halo: records after its right cut that start before the largest reference end of the shard
 *                 (bk_bed_chrom_max_end of the parsed reference slice, then bk_find_start on the host text), and
 *   a left halo:  records before its left cut whose end passes the cut.  Where it begins is known only to the shards
 *                 that parsed those records: every shard reports, for each later cut in a chromosome it holds,
 *                 bk_bed_reach_start = the smallest start among its records that end beyond the cut (the prefix-max-end
 *                 index), the shards exchange these few numbers (the path's only collective: an allgather of
 *                 n_shards^2 u64), and the minimum over the earlier shards, bisected in the host text with bk_find_start,
 *                 is the halo's first byte.  bk_bed_concat puts halo and own records into one bk_bed.
 * Replaces the reference's seek to a chromosome by bisection over byte offsets (AllocateIterator_BED_starch.hpp:113-160
 * -> FindBedRange.hpp:68) with a seek to any genomic position. */
typedef struct bk_cut {
  char     chrom[128];
  uint64_t coord;  /* records with start >= coord (in chrom) belong to the right of the cut; 0 = the chromosome's first record */
  int      at_end; /* the cut lies behind the last record (empty shards to its right) */
} bk_cut;
uint64_t bk_find_start(const char* host_text, uint64_t begin, uint64_t end, uint64_t coord);
int      bk_plan_cuts(const char* host_text, size_t nbytes, const bk_chrom_span* idx, int n_idx, int n_shards, bk_cut* cuts);
uint64_t bk_cut_offset(const char* host_text, size_t nbytes, const bk_chrom_span* idx, int n_idx, const bk_cut* cut);
/* smallest start among the records of `chrom` whose end > pos; UINT64_MAX if none (or the chromosome is absent) */
int bk_bed_reach_start(bk_ctx* ctx, const bk_bed* bed, const char* chrom, uint64_t pos, uint64_t* start_out);
/* largest end among the records of `chrom`; 0 if none */
int bk_bed_chrom_max_end(bk_ctx* ctx, const bk_bed* bed, const char* chrom, uint64_t* end_out);
/* head ++ tail as one bk_bed (every record of head sorts before every record of tail; same min_fields and columns).
 * The result owns copies of the columns (and of the text when lines are kept); head and tail stay valid. */
int bk_bed_concat(bk_ctx* ctx, const bk_bed* head, const bk_bed* tail, bk_bed** out);

/* ---- range-sharded bedmap: ONE dataset over N GPUs (one rank = one bk_ctx = one GPU) ----------------------------------
 * bk_shard_plan_make (host only, deterministic: every rank computes the same plan from the same text) cuts the larger
 * file into n_shards byte-balanced genomic ranges and locates the cuts in the other file.
 * bk_bedmap_shard_begin uploads and parses this rank's reference slice and its map slice (own records + right halo)
 * and fills reach[j] (j > rank; UINT64_MAX = none) for the exchange; the caller allgathers the n_shards vectors
 * (row i = rank i's) -- NCCL across processes, shared memory across threads -- and bk_bedmap_shard_finish adds the left
 * halo, maps, and returns this rank's part of the output.  The parts in rank order are byte-identical to the unsharded
 * call.  *_src: where the bytes are copied from -- NULL (the host text itself) or a mirror with identical offsets
 * (pinned host or device memory); the host text is always needed for the bisections. */
#define BK_MAX_SHARDS 64
typedef struct bk_shard_plan {
  int      n_shards;
  bk_cut   cuts[BK_MAX_SHARDS];            /* cuts[k] separates shard k from shard k+1 (n_shards - 1 used) */
  uint64_t ref_off[BK_MAX_SHARDS + 1];     /* shard k owns ref bytes [ref_off[k], ref_off[k+1]) */
  uint64_t map_off[BK_MAX_SHARDS + 1];     /* ... and map bytes [map_off[k], map_off[k+1]) */
  uint64_t map_chrom_begin[BK_MAX_SHARDS]; /* byte span, in the map file, of the chromosome cut k falls into */
  uint64_t map_chrom_end[BK_MAX_SHARDS];
} bk_shard_plan;
typedef struct bk_shard bk_shard;
/* BK_ERR_UNSORTED when the chromosomes of a file are not in strictly ascending strcmp order (run unsharded then) */
int bk_shard_plan_make(const char* ref_text, size_t ref_len, const char* map_text, size_t map_len, int n_shards,
                       bk_shard_plan* plan);
int bk_bedmap_shard_begin(bk_ctx* ctx, const bk_shard_plan* plan, int rank, const char* ref_text, size_t ref_len,
                          int ref_fields, unsigned ref_cols, const char* map_text, size_t map_len, int map_fields,
                          unsigned map_cols, const char* ref_src, const char* map_src, const bk_mapspec* spec,
                          bk_shard** out, uint64_t* reach);
int bk_bedmap_shard_finish(bk_ctx* ctx, bk_shard* shard, const uint64_t* all_reach, bk_text* out);
void bk_shard_free(bk_ctx* ctx, bk_shard* shard); /* after finish, or to abandon a call after begin */
/* bytes this rank copied from *_src so far (own slices + halos): bench.py's h2d_bytes_per_step */
uint64_t bk_shard_bytes_in(const bk_shard* shard);

#ifdef __cplusplus
}
#endif
#endif /* BEDKIT_H */
