// bedmap -- drop-in command line for the B200 engine.  Mirrors the option grammar, output text and exit codes of
// the reference tool (applications/bed/bedmap/src/Input.hpp:75-367, Bedmap.cpp:95-187) and replaces its
// WindowSweep::sweep call (Bedmap.cpp:288) with bk_bedmap().
#include <sstream>
#include "cli_common.hpp"
#include "help_text.hpp"

namespace {

using cli::UserError;

struct OpInfo {
  const char* name;
  int         op;          // BK_OP_* or 0 if the operation is outside the device hot path
  int         map_fields;  // Input.hpp:401-418 (MapFields::num)
  int         nargs;       // fixed extra arguments (only the out-of-scope ops take any)
};
const OpInfo kOps[] = {
    {"bases", BK_OP_BASES, 3, 0},          {"bases-uniq", BK_OP_BASES_UNIQ, 3, 0},           {"bases-uniq-f", BK_OP_BASES_UNIQ_F, 3, 0},
    {"echo", BK_OP_ECHO, 3, 0},            {"echo-ref-size", BK_OP_ECHO_REF_SIZE, 3, 0},
    {"echo-ref-name", BK_OP_ECHO_REF_NAME, 3, 0},                             {"echo-ref-row-id", BK_OP_ECHO_REF_ROW_ID, 3, 0},
    {"echo-map", BK_OP_ECHO_MAP, 3, 0},                 {"echo-map-id", BK_OP_ECHO_MAP_ID, 4, 0},
    {"echo-map-id-uniq", BK_OP_ECHO_MAP_ID_UNIQ, 4, 0},         {"echo-map-size", BK_OP_ECHO_MAP_SIZE, 3, 0},        {"echo-overlap-size", BK_OP_ECHO_OVERLAP_SIZE, 3, 0},
    {"echo-map-range", BK_OP_ECHO_MAP_RANGE, 3, 0},           {"echo-map-score", BK_OP_ECHO_MAP_SCORE, 5, 0},       {"count", BK_OP_COUNT, 3, 0},
    {"indicator", BK_OP_INDICATOR, 3, 0},  {"max", BK_OP_MAX, 5, 0},          {"max-element-rand", 0, 5, 0},
    {"max-element", BK_OP_MAX_ELEMENT, 5, 0},              {"min", BK_OP_MIN, 5, 0},          {"min-element-rand", 0, 5, 0},
    {"min-element", BK_OP_MIN_ELEMENT, 5, 0},              {"mean", BK_OP_MEAN, 5, 0},        {"variance", BK_OP_VARIANCE, 5, 0},
    {"stdev", BK_OP_STDEV, 5, 0},                    {"cv", BK_OP_CV, 5, 0},                   {"sum", BK_OP_SUM, 5, 0},
    {"wmean", BK_OP_WMEAN, 5, 0},                    {"median", BK_OP_MEDIAN, 5, 0},               {"mad", BK_OP_MAD, 5, -1},
    {"kth", BK_OP_KTH, 5, 1},                      {"tmean", BK_OP_TMEAN, 5, 2},
};

struct Options {
  std::string ref, map;
  std::vector<int> ops;
  std::vector<double> op_args;  // parallel to ops (--kth <val>, --mad <mult>, --tmean <low>)
  std::vector<double> op_args2; // --tmean <hi>
  std::string unsupported_op;
  int         overlap_kind = -1;
  long        range_bp = 0, overlap_bp = 0;
  double      frac = 0;
  bool        is_range = false, is_bp = false, range_alias = false;
  bool        pm = false, pr = false, pe = false, pb = false, exact = false;
  int         precision = 6;
  bool        sci = false, set_prec = false, ec = false, sweep_all = false, fast = false, skip_unmapped = false;
  std::string delim = "|", multidelim = ";", chrom = "all";
  int         min_ref_fields = 0, min_map_fields = 0;
  int         num_files = 0;
};

void require(bool ok, const std::string& msg) {
  if (!ok) throw UserError(msg);
}

double parse_frac(const std::string& sval) {
  std::stringstream conv(sval);
  double            v = 0;
  conv >> v;
  return v;
}

struct Help {};
struct Version {};
struct NoInput {};

Options parse_args(int argc, char** argv) {
  Options o;
  if (argc <= 1) throw NoInput();
  const char* pos_ints = "0123456789";
  const char* reals = ".-0123456789";
  int         i = 1;
  bool        has_op = false;
  while (i < argc) {
    std::string next = argv[i++];
    if (next.find("--") == std::string::npos && argc - i < 2) break;  // file inputs
    require(next.find("--") == 0, "Option " + next + " does not start with '--'");
    next = next.substr(2);
    auto frac_opt = [&](bool& flag, const char* nm) {
      require(!flag, std::string("multiple --") + nm + "'s detected");
      require(i < argc, std::string("No arg for --") + nm);
      std::string sval = argv[i++];
      require(cli::only_chars(sval, reals), "Non-numeric argument: " + sval + " for --" + nm);
      o.frac = parse_frac(sval);
      require(o.frac > 0 && o.frac <= 1, std::string("--") + nm + " value must be: >0-1.0");
      flag = true;
    };
    if (next == "help") throw Help();
    else if (next == "version") throw Version();
    else if (next == "ec" || next == "header") o.ec = true;
    else if (next == "faster") o.fast = true;
    else if (next == "sweep-all") o.sweep_all = true;
    else if (next == "delim") {
      require(o.delim == "|", "--delim specified multiple times");
      require(i < argc, "No output delimiter given");
      o.delim = argv[i++];
      require(o.delim.find("--") != 0, "Apparent option: " + o.delim + " where output delimiter expected.");
    } else if (next == "chrom") {
      require(o.chrom == "all", "--chrom specified multiple times");
      require(i < argc, "No chromosome name given");
      o.chrom = argv[i++];
      require(o.chrom.find("--") != 0, "Apparent option: " + o.chrom + " where chromosome expected.");
    } else if (next == "multidelim") {
      require(o.multidelim == ";", "--multidelim specified multiple times");
      require(i < argc, "No multi-value column delimmiter given");
      o.multidelim = argv[i++];
      require(o.multidelim.find("--") != 0, "Apparent option: " + o.multidelim + " where output delimiter expected.");
    } else if (next == "skip-unmapped") o.skip_unmapped = true;
    else if (next == "sci") o.sci = true;
    else if (next == "prec") {
      require(i < argc, "No precision value given");
      require(!o.set_prec, "--prec specified multiple times.");
      std::string sval = argv[i++];
      require(cli::only_chars(sval, pos_ints), "Non-positive-integer argument: " + sval + " for --prec");
      o.precision = std::atoi(sval.c_str());
      require(o.precision >= 0, "--prec value must be >= 0");
      o.set_prec = true;
    } else if (next == "bp-ovr") {
      require(!o.range_alias, "--range and --bp-ovr detected.  Choose one.");
      require(!o.is_bp, "multiple --bp-ovr's detected");
      require(i < argc, "No arg for --bp-ovr");
      std::string sval = argv[i++];
      require(cli::only_chars(sval, pos_ints), "Non-positive-integer argument: " + sval + " for --bp-ovr");
      o.overlap_bp = std::atol(sval.c_str());
      require(o.overlap_bp > 0, "--bp-ovr value must be > 0");
      o.is_bp = true;
    } else if (next == "range") {
      require(!o.is_range && !o.range_alias, "multiple --range's detected");
      require(i < argc, "No arg for --range");
      std::string sval = argv[i++];
      require(cli::only_chars(sval, pos_ints), "Non-positive-integer argument: " + sval + " for --range");
      o.range_bp = std::atol(sval.c_str());
      require(o.range_bp >= 0, "--range value must be >= 0");
      o.is_range = true;
      if (o.range_bp == 0) {  // alias for --bp-ovr 1
        require(!o.is_bp, "--bp-ovr and --range detected.  Choose one.");
        o.is_range = false;
        o.is_bp = true;
        o.range_alias = true;
        o.overlap_bp = 1;
      }
    } else if (next == "fraction-ref") frac_opt(o.pr, "fraction-ref");
    else if (next == "fraction-map") frac_opt(o.pm, "fraction-map");
    else if (next == "fraction-either") frac_opt(o.pe, "fraction-either");
    else if (next == "fraction-both") frac_opt(o.pb, "fraction-both");
    else if (next == "exact") {
      require(!o.exact, "multiple --exact's detected - use one");
      o.exact = true;
    } else {
      const OpInfo* info = nullptr;
      for (const OpInfo& k : kOps)
        if (next == k.name) info = &k;
      if (!info) throw UserError("Unknown option: --" + next);
      double op_arg = 0, op_arg2 = 0;
      int    op = info->op;
      if (op == BK_OP_TMEAN) {  // Input.hpp:303-326
        require(i < argc, "No <low> arg given for --" + next);
        const std::string lo = argv[i++];
        require(cli::only_chars(lo, reals), "Non-numeric argument: " + lo + " for --" + next);
        require(i < argc, "No <hi> arg given for --" + next);
        const std::string hi = argv[i++];
        require(cli::only_chars(hi, reals), "Non-numeric argument: " + hi + " for --" + next);
        op_arg = op_arg2 = 100;
        std::stringstream cl(lo), ch(hi);
        cl >> op_arg;
        ch >> op_arg2;
        require(op_arg >= 0 && op_arg <= 1, "--" + next + " Expect 0 <= low < hi <= 1");
        require(op_arg2 >= 0 && op_arg2 <= 1, "--" + next + " Expect 0 <= low < hi <= 1");
        require(op_arg + op_arg2 <= 1, "--" + next + " Expect (low + hi) <= 1.");
      } else if (info->nargs > 0) {
        require(i + info->nargs <= argc, "No arg for --" + next);
        if (info->op == BK_OP_KTH) {  // Input.hpp:290-302
          const std::string sval = argv[i];
          require(cli::only_chars(sval, reals), "Non-numeric argument: " + sval + " for --" + next);
          op_arg = -1;
          std::stringstream conv(sval);
          conv >> op_arg;
          require(op_arg >= 0 && op_arg <= 1, "--" + next + " Expect 0 <= val <= 1");
          if (op_arg == 0) op = BK_OP_MIN;       // "min faster" / "max faster", Bedmap.cpp:495-498
          else if (op_arg == 1) op = BK_OP_MAX;
        }
        i += info->nargs;
      } else if (info->nargs < 0 && i < argc && cli::only_chars(argv[i], reals)) {  // optional multiplier of --mad
        op_arg = -1;
        std::stringstream conv(std::string(argv[i]));
        conv >> op_arg;
        require(op_arg > 0, "--" + next + " Expect 0 < val");  // Input.hpp:275-288
        i++;
      }
      if (!op && o.unsupported_op.empty()) o.unsupported_op = next;
      o.ops.push_back(op);
      o.op_args.push_back(op_arg);
      o.op_args2.push_back(op_arg2);
      o.min_map_fields = std::max(o.min_map_fields, info->map_fields);
      o.min_ref_fields = std::max(o.min_ref_fields, 3);
      has_op = true;
    }
  }
  if (!(o.pm || o.pr || o.pe || o.pb || o.is_range || o.is_bp || o.exact)) {
    o.is_bp = true;
    o.overlap_bp = 1;
  }
  int count = o.pm + o.pr + o.pe + o.pb + o.is_range + o.is_bp + o.exact;
  require(count == 1, "More than one overlap specification used.");
  require(has_op, "No processing option specified (ie; --max).");
  require(!o.fast || o.is_bp || o.is_range || o.pb || o.exact,
          "--faster compatible with --range, --bp-ovr, --fraction-both, and --exact only");
  require(argc - i <= 2, "Need [one or] two input files");
  o.num_files = argc - i + 1;
  require(o.num_files >= 1 && o.num_files <= 2, "Need [one or] two input files");
  if (o.num_files == 2) {
    o.ref = argv[argc - 2];
    o.map = argv[argc - 1];
  } else {
    o.ref = argv[argc - 1];
    o.min_ref_fields = o.min_map_fields;
    o.min_map_fields = 0;
  }
  require(o.ref != "-" || o.map != "-", "Cannot have stdin set for two files");
  o.overlap_kind = o.pm ? BK_OVR_FRAC_MAP : o.pr ? BK_OVR_FRAC_REF : o.pe ? BK_OVR_FRAC_EITHER : o.pb ? BK_OVR_FRAC_BOTH
                   : o.exact ? BK_OVR_EXACT : o.is_range ? BK_OVR_RANGE : BK_OVR_BP;
  return o;
}

std::string unescape_delim(const std::string& d) {  // PrintDelim, ProcessVisitorRow.hpp:107-126
  if (d == "\t" || d == "\\t" || d == "'\t'") return "\t";
  if (d == "\n" || d == "\\n" || d == "'\n'") return "\n";
  return d;
}

void usage(FILE* f) { std::fputs(kUsageBedmap, f); }  // byte for byte the reference's text (help_text.hpp)

}  // namespace

int main(int argc, char** argv) {
  try {
    cli::trace_lap("start");
    Options o = parse_args(argc, argv);
    if (!o.unsupported_op.empty())
      throw UserError("--" + o.unsupported_op + " is not on the B200 hot path of this build (see DESIGN.md, out of scope)");
    cli::Input rtext, mtext;
    if (!rtext.open(o.ref)) throw UserError("Unable to find file: " + o.ref);
    if (o.num_files == 2 && !mtext.open(o.map)) throw UserError("Unable to find file: " + o.map);
    if (cli::any_archive({&rtext, &mtext})) {
      cli::Engine eng;
      cli::unstarch_if_archive(eng, rtext);
      cli::unstarch_if_archive(eng, mtext);
    }

    bk_mapspec spec;
    bk_mapspec_default(&spec);
    bool need_line = false, need_score = false, need_id = false, need_mapline = false, element_ops = false;
    if (o.ops.size() > (size_t)BK_MAX_OPS)  // bk_mapspec carries a fixed table (the reference chains any number of visitors)
      throw UserError("More than " + std::to_string(BK_MAX_OPS) + " operations given; this build prints at most that many columns.");
    for (size_t k = 0; k < o.ops.size(); k++) {
      const int op = o.ops[k];
      spec.op_arg[spec.n_ops] = o.op_args[k];
      spec.op_arg2[spec.n_ops] = o.op_args2[k];
      element_ops |= op == BK_OP_MAX_ELEMENT || op == BK_OP_MIN_ELEMENT;
      need_mapline |= op == BK_OP_MAX_ELEMENT || op == BK_OP_MIN_ELEMENT;
      need_score |= op == BK_OP_WMEAN || op == BK_OP_TMEAN || op == BK_OP_MAX_ELEMENT || op == BK_OP_MIN_ELEMENT;
      spec.ops[spec.n_ops++] = op;
      need_line |= op == BK_OP_ECHO || op == BK_OP_ECHO_REF_NAME || op == BK_OP_ECHO_MAP_RANGE;
      need_score |= op == BK_OP_SUM || op == BK_OP_MEAN || op == BK_OP_MAX || op == BK_OP_MIN || op == BK_OP_ECHO_MAP_SCORE ||
                    op == BK_OP_VARIANCE || op == BK_OP_STDEV || op == BK_OP_CV || op == BK_OP_MEDIAN || op == BK_OP_KTH ||
                    op == BK_OP_MAD;
      need_id |= op == BK_OP_ECHO_MAP_ID || op == BK_OP_ECHO_MAP_ID_UNIQ;
      need_mapline |= op == BK_OP_ECHO_MAP;
    }
    spec.overlap_kind = o.overlap_kind;
    spec.overlap_bp = o.overlap_kind == BK_OVR_RANGE ? (uint64_t)o.range_bp : (uint64_t)o.overlap_bp;
    spec.overlap_frac = o.frac;
    spec.precision = o.precision;
    spec.sci = o.sci;
    spec.skip_unmapped = o.skip_unmapped;
    std::string delim = unescape_delim(o.delim), mdelim = unescape_delim(o.multidelim);
    spec.delim = delim.c_str();
    spec.multidelim = mdelim.c_str();
    spec.chrom = o.chrom.c_str();

    const unsigned hdr = o.ec ? BK_LOAD_HEADERS : 0;
    const unsigned map_cols = (need_score ? BK_COL_SCORE : 0) | (need_id ? (BK_COL_ID | BK_COL_LINE) : 0) |
                              (need_mapline ? BK_COL_LINE : 0) | hdr;
    if (o.ec) {  // validate first (Bedmap.cpp:229-253, :292-329 use bed_check_iterator under --ec)
      cli::ec_prepare(rtext);
      cli::ec_prepare(mtext);
      cli::Engine eng;
      if (o.num_files == 2) {
        cli::ec_check(eng, rtext, o.ref, 3, true, o.fast);
        cli::ec_check(eng, mtext, o.map, o.min_map_fields, true, o.fast);
      } else {
        cli::ec_check(eng, rtext, o.ref, o.min_ref_fields, true, o.fast);
      }
    }
    const unsigned ref_cols = (need_line ? BK_COL_LINE : 0) | hdr;
    bool row_ids = false;
    for (int op : o.ops) row_ids |= op == BK_OP_ECHO_REF_ROW_ID;
    const int gpus = cli::gpus_requested();
    // N GPUs: ONE dataset cut into genomic ranges with boundary halos (include/bedkit.h); the printed-row counter of
    // --echo-ref-row-id runs across shards, so that operation stays on one GPU
    if (gpus > 1 && o.num_files == 2 && o.chrom == "all" && !row_ids && !element_ops &&
        cli::run_range_sharded_bedmap(rtext, mtext, 3, ref_cols, o.min_map_fields, map_cols, spec, gpus))
      return EXIT_SUCCESS;
    cli::trace_lap("inputs mapped");
    cli::Engine eng;
    cli::trace_lap("bk_init done");
    bk_text     out;
    int         rc;
    if (o.num_files == 2) {
      // the whole call over the (mapped) host text: chromosome groups uploaded, parsed, mapped and downloaded in a pipeline
      rc = bk_bedmap_host(eng.ctx, rtext.data, rtext.size, 3, ref_cols, mtext.data, mtext.size, o.min_map_fields, map_cols, &spec, &out);
    } else {
      bk_bed* ref = eng.load(rtext, o.min_ref_fields, map_cols | (need_line ? BK_COL_LINE : 0));
      rc = bk_bedmap(eng.ctx, ref, nullptr, &spec, &out);
      bk_free_bed(eng.ctx, ref);
    }
    if (rc == BK_ERR_NAN_ELEMENT) {  // the rows the reference had printed before it threw, then its message
      cli::write_all(out.ptr, out.len);
      bk_free_text(eng.ctx, &out);
      eng.raise(rc);
    }
    if (rc != BK_OK) eng.raise(rc);
    cli::trace_lap("result on the host");
    cli::write_all(out.ptr, out.len);  // pinned result buffer -> stdout, no intermediate copy
    cli::trace_lap("written");
    rtext.settle();
    mtext.settle();
    cli::finish_now(EXIT_SUCCESS);
  } catch (const Help&) {
    cli::banner(stdout, "bedmap");
    usage(stdout);
    return EXIT_SUCCESS;
  } catch (const Version&) {
    cli::banner(stdout, "bedmap");  // the reference falls through to EXIT_FAILURE here (Bedmap.cpp:166-170, :186)
  } catch (const NoInput&) {
    cli::banner(stderr, "bedmap");
    usage(stderr);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "May use bedmap --help for more help.\n\nError: %s\n", e.what());
  }
  return EXIT_FAILURE;
}
