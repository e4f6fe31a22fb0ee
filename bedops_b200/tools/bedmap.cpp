// bedmap -- drop-in command line for the B200 engine.  Mirrors the option grammar, output text and exit codes of
// the reference tool (applications/bed/bedmap/src/Input.hpp:75-367, Bedmap.cpp:95-187) and replaces its
// WindowSweep::sweep call (Bedmap.cpp:288) with bk_bedmap().
#include <sstream>
#include "cli_common.hpp"
#include "help_text.hpp"

namespace {

using cli::UserError;

// What follows an operation's name on the command line (Input.hpp:275-326)
enum class OpArg { None, Multiplier /* --mad [mult] */, Quantile /* --kth <val> */, TrimPair /* --tmean <low> <hi> */ };

struct OpInfo {
  const char* name;
  int         op;          // BK_OP_* or 0 if the operation is outside the device hot path
  int         map_fields;  // Input.hpp:401-418 (MapFields::num)
  OpArg       arg;
};
const OpInfo kOps[] = {
    {"bases", BK_OP_BASES, 3, OpArg::None},
    {"bases-uniq", BK_OP_BASES_UNIQ, 3, OpArg::None},
    {"bases-uniq-f", BK_OP_BASES_UNIQ_F, 3, OpArg::None},
    {"echo", BK_OP_ECHO, 3, OpArg::None},
    {"echo-ref-size", BK_OP_ECHO_REF_SIZE, 3, OpArg::None},
    {"echo-ref-name", BK_OP_ECHO_REF_NAME, 3, OpArg::None},
    {"echo-ref-row-id", BK_OP_ECHO_REF_ROW_ID, 3, OpArg::None},
    {"echo-map", BK_OP_ECHO_MAP, 3, OpArg::None},
    {"echo-map-id", BK_OP_ECHO_MAP_ID, 4, OpArg::None},
    {"echo-map-id-uniq", BK_OP_ECHO_MAP_ID_UNIQ, 4, OpArg::None},
    {"echo-map-size", BK_OP_ECHO_MAP_SIZE, 3, OpArg::None},
    {"echo-overlap-size", BK_OP_ECHO_OVERLAP_SIZE, 3, OpArg::None},
    {"echo-map-range", BK_OP_ECHO_MAP_RANGE, 3, OpArg::None},
    {"echo-map-score", BK_OP_ECHO_MAP_SCORE, 5, OpArg::None},
    {"count", BK_OP_COUNT, 3, OpArg::None},
    {"indicator", BK_OP_INDICATOR, 3, OpArg::None},
    {"max", BK_OP_MAX, 5, OpArg::None},
    {"max-element-rand", 0, 5, OpArg::None},
    {"max-element", BK_OP_MAX_ELEMENT, 5, OpArg::None},
    {"min", BK_OP_MIN, 5, OpArg::None},
    {"min-element-rand", 0, 5, OpArg::None},
    {"min-element", BK_OP_MIN_ELEMENT, 5, OpArg::None},
    {"mean", BK_OP_MEAN, 5, OpArg::None},
    {"variance", BK_OP_VARIANCE, 5, OpArg::None},
    {"stdev", BK_OP_STDEV, 5, OpArg::None},
    {"cv", BK_OP_CV, 5, OpArg::None},
    {"sum", BK_OP_SUM, 5, OpArg::None},
    {"wmean", BK_OP_WMEAN, 5, OpArg::None},
    {"median", BK_OP_MEDIAN, 5, OpArg::None},
    {"mad", BK_OP_MAD, 5, OpArg::Multiplier},
    {"kth", BK_OP_KTH, 5, OpArg::Quantile},
    {"tmean", BK_OP_TMEAN, 5, OpArg::TrimPair},
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

struct Help {};
struct Version {};
struct NoInput {};

// ---- the settings (everything that is not an operation) as data ---------------------------------------------------
// The reference's Input constructor (Input.hpp:92-330) is one long if-chain; what it accepts and what it says when it
// does not is kept here as a table that one interpreter walks.  A row names where the value goes, the complaints in the
// reference's words, and -- because the chain does not always complain in the same order -- whether the missing value
// or the repetition is noticed first.
enum class Takes { Nothing, Text, Count, Fraction };
constexpr const char* kDigits = "0123456789";
constexpr const char* kReal = ".-0123456789";

struct Setting {
  const char* name;
  Takes       takes;
  bool Options::*        seen;        // Nothing: the switch itself; Count / Fraction: "was given" (repetition test)
  std::string Options::* text;        // Text: destination; "given before" = differs from its default
  const char* text_default;
  const char* role;                   // Text: "Apparent option: X where <role> expected."
  const char* repeated;               // complaint on repetition (nullptr: repeating is harmless)
  const char* missing;                // complaint when argv ends here
  bool        missing_first;          // --prec looks for its value before it notices the repetition
  void (*store)(Options&, const std::string& value);  // Count / Fraction: convert, range-check, store
  void (*before)(Options&);           // conflicts the reference tests before anything else
};

// The reference converts with `stringstream >> member`: an empty word leaves the member as it was (a failed sentry does not
// touch the target), an overflowing one saturates it.  Extract into the very field, as it does.
template <class T>
void extract(const std::string& word, T& into) {
  std::stringstream conv(word);
  conv >> into;
}
// argv[argc]: the reference builds a std::string from it in two places without looking (Input.hpp:116, :124, :132, :275)
[[noreturn]] void null_word() { throw UserError("basic_string: construction from null is not valid"); }

const Setting kSettings[] = {
    {"ec", Takes::Nothing, &Options::ec},
    {"header", Takes::Nothing, &Options::ec},
    {"faster", Takes::Nothing, &Options::fast},
    {"sweep-all", Takes::Nothing, &Options::sweep_all},
    {"skip-unmapped", Takes::Nothing, &Options::skip_unmapped},
    {"sci", Takes::Nothing, &Options::sci},
    {"exact", Takes::Nothing, &Options::exact, nullptr, nullptr, nullptr, "multiple --exact's detected - use one"},
    {"delim", Takes::Text, nullptr, &Options::delim, "|", "output delimiter", "--delim specified multiple times", "No output delimiter given"},
    {"multidelim", Takes::Text, nullptr, &Options::multidelim, ";", "output delimiter", "--multidelim specified multiple times",
     "No multi-value column delimmiter given"},
    {"chrom", Takes::Text, nullptr, &Options::chrom, "all", "chromosome", "--chrom specified multiple times", "No chromosome name given"},
    {"prec", Takes::Count, &Options::set_prec, nullptr, nullptr, nullptr, "--prec specified multiple times.", "No precision value given", true,
     [](Options& o, const std::string& v) {
       extract(v, o.precision);
       require(o.precision >= 0, "--prec value must be >= 0");
     }},
    {"bp-ovr", Takes::Count, &Options::is_bp, nullptr, nullptr, nullptr, "multiple --bp-ovr's detected", "No arg for --bp-ovr", false,
     [](Options& o, const std::string& v) {
       extract(v, o.overlap_bp);
       require(o.overlap_bp > 0, "--bp-ovr value must be > 0");
     },
     [](Options& o) { require(!o.range_alias, "--range and --bp-ovr detected.  Choose one."); }},
    {"range", Takes::Count, &Options::is_range, nullptr, nullptr, nullptr, "multiple --range's detected", "No arg for --range", false,
     [](Options& o, const std::string& v) {
       extract(v, o.range_bp);
       require(o.range_bp >= 0, "--range value must be >= 0");
     },
     [](Options& o) { require(!o.range_alias, "multiple --range's detected"); }},
    {"fraction-ref", Takes::Fraction, &Options::pr},
    {"fraction-map", Takes::Fraction, &Options::pm},
    {"fraction-either", Takes::Fraction, &Options::pe},
    {"fraction-both", Takes::Fraction, &Options::pb},
};

// argv cursor: the interpreter and the operation arguments pull their values from it
struct Args {
  int    argc;
  char** argv;
  int    i = 1;
  bool        more() const { return i < argc; }
  int         left() const { return argc - i; }
  std::string take() { return argv[i++]; }
  const char* peek() const { return argv[i]; }
};

void apply_setting(const Setting& st, Options& o, Args& a) {
  const std::string dashed = std::string("--") + st.name;
  if (st.before) st.before(o);
  const bool again = st.takes == Takes::Text ? o.*st.text != st.text_default : (st.takes != Takes::Nothing || st.repeated) && o.*st.seen;
  const std::string repeated = st.repeated ? st.repeated : "multiple " + dashed + "'s detected";
  const std::string missing = st.missing ? st.missing : "No arg for " + dashed;
  if (st.takes == Takes::Nothing) {
    require(!again, repeated);
    o.*st.seen = true;
    return;
  }
  if (st.missing_first) require(a.more(), missing);
  require(!again, repeated);
  require(a.more(), missing);
  const std::string value = a.take();
  switch (st.takes) {
    case Takes::Text:
      o.*st.text = value;
      // the complaint is put together before the test, and from the word AFTER the value: with nothing behind the value
      // that alone fails, and a value that looks like an option is reported under its successor's name
      if (!a.more()) null_word();
      require(value.find("--") != 0, std::string("Apparent option: ") + a.peek() + " where " + st.role + " expected.");
      return;
    case Takes::Count:
      require(cli::only_chars(value, kDigits), "Non-positive-integer argument: " + value + " for " + dashed);
      st.store(o, value);
      break;
    case Takes::Fraction:
      require(cli::only_chars(value, kReal), "Non-numeric argument: " + value + " for " + dashed);
      extract(value, o.frac);
      require(o.frac > 0 && o.frac <= 1, dashed + " value must be: >0-1.0");
      break;
    default: break;
  }
  o.*st.seen = true;
}

// the numbers that may follow an operation's name; returns the operation actually run (--kth 0 / 1 are --min / --max)
int read_op_args(const OpInfo& info, const std::string& dashed, Args& a, double* arg1, double* arg2) {
  auto number = [&](const std::string& text, double otherwise) {
    require(cli::only_chars(text, kReal), "Non-numeric argument: " + text + " for " + dashed);
    double            v = otherwise;
    std::stringstream conv(text);
    conv >> v;
    return v;
  };
  switch (info.arg) {
    case OpArg::None: break;
    case OpArg::Multiplier:  // optional: only a numeric-looking word is taken (Input.hpp:275-288); the word is read unseen
      if (!a.more()) null_word();
      if (cli::only_chars(a.peek(), kReal)) {
        *arg1 = number(a.take(), -1);
        require(*arg1 > 0, dashed + " Expect 0 < val");
      }
      break;
    case OpArg::Quantile:  // Input.hpp:290-302
      require(a.more(), "No arg for " + dashed);
      *arg1 = number(a.take(), -1);
      require(*arg1 >= 0 && *arg1 <= 1, dashed + " Expect 0 <= val <= 1");
      if (*arg1 == 0) return BK_OP_MIN;  // "min faster" / "max faster", Bedmap.cpp:495-498
      if (*arg1 == 1) return BK_OP_MAX;
      break;
    case OpArg::TrimPair: {  // Input.hpp:303-326: both words are looked at before either number is judged
      require(a.more(), "No <low> arg given for " + dashed);
      const std::string lo = a.take();
      require(cli::only_chars(lo, kReal), "Non-numeric argument: " + lo + " for " + dashed);
      require(a.more(), "No <hi> arg given for " + dashed);
      const std::string hi = a.take();
      *arg1 = number(lo, 100);
      *arg2 = number(hi, 100);
      require(*arg1 >= 0 && *arg1 <= 1, dashed + " Expect 0 <= low < hi <= 1");
      require(*arg2 >= 0 && *arg2 <= 1, dashed + " Expect 0 <= low < hi <= 1");
      require(*arg1 + *arg2 <= 1, dashed + " Expect (low + hi) <= 1.");
      break;
    }
  }
  return info.op;
}

template <class Row, size_t N>
const Row* lookup(const Row (&table)[N], const std::string& name) {
  const Row* hit = nullptr;
  for (const Row& r : table)
    if (name == r.name) hit = &r;
  return hit;
}

Options parse_args(int argc, char** argv) {
  Options o;
  if (argc <= 1) throw NoInput();
  Args a{argc, argv};
  bool has_op = false;
  while (a.more()) {
    const std::string word = a.take();
    if (word.find("--") == std::string::npos && a.left() < 2) break;  // file inputs
    require(word.find("--") == 0, "Option " + word + " does not start with '--'");
    const std::string name = word.substr(2);
    if (name == "help") throw Help();
    if (name == "version") throw Version();
    if (const Setting* st = lookup(kSettings, name)) {
      apply_setting(*st, o, a);
      if (st->store && o.is_range && o.range_bp == 0 && name == "range") {  // --range 0 is an alias for --bp-ovr 1
        require(!o.is_bp, "--bp-ovr and --range detected.  Choose one.");
        o.is_range = false;
        o.is_bp = o.range_alias = true;
        o.overlap_bp = 1;
      }
      continue;
    }
    const OpInfo* info = lookup(kOps, name);
    if (!info) throw UserError("Unknown option: --" + name);
    double    arg1 = 0, arg2 = 0;
    const int op = read_op_args(*info, "--" + name, a, &arg1, &arg2);
    if (!op && o.unsupported_op.empty()) o.unsupported_op = name;
    o.ops.push_back(op);
    o.op_args.push_back(arg1);
    o.op_args2.push_back(arg2);
    o.min_map_fields = std::max(o.min_map_fields, info->map_fields);
    o.min_ref_fields = std::max(o.min_ref_fields, 3);
    has_op = true;
  }
  const int i = a.i;
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

void dump_options(const Options& o) {
  std::printf("ref=%s map=%s files=%d overlap=%d range_bp=%ld overlap_bp=%ld frac=%.17g prec=%d sci=%d ec=%d sweep_all=%d fast=%d "
              "skip_unmapped=%d delim=[%s] multidelim=[%s] chrom=[%s] ref_fields=%d map_fields=%d unsupported=[%s] ops=",
              o.ref.c_str(), o.map.c_str(), o.num_files, o.overlap_kind, o.range_bp, o.overlap_bp, o.frac, o.precision, (int)o.sci, (int)o.ec,
              (int)o.sweep_all, (int)o.fast, (int)o.skip_unmapped, o.delim.c_str(), o.multidelim.c_str(), o.chrom.c_str(), o.min_ref_fields,
              o.min_map_fields, o.unsupported_op.c_str());
  for (size_t k = 0; k < o.ops.size(); k++) std::printf("%s%d(%.17g,%.17g)", k ? "," : "", o.ops[k], o.op_args[k], o.op_args2[k]);
  std::printf("\n");
}

void usage(FILE* f) { std::fputs(kUsageBedmap, f); }  // byte for byte the reference's text (help_text.hpp)

}  // namespace

int main(int argc, char** argv) {
  try {
    cli::trace_lap("start");
    Options o = parse_args(argc, argv);
    if (std::getenv("BEDKIT_DUMP_OPTIONS")) {  // what the command line was understood as, then stop (host-logic tests, no GPU)
      dump_options(o);
      return EXIT_SUCCESS;
    }
    if (!o.unsupported_op.empty())
      throw UserError("--" + o.unsupported_op + " is not on the B200 hot path of this build (see DESIGN.md, out of scope)");
    cli::Input rtext, mtext;
    // FPWrap.hpp:42 says "Unable to find file: X"; the --ec branch opens ifstreams itself and says "Unable to find: X" (Bedmap.cpp:231, :300-302)
    const std::string not_found = o.ec ? "Unable to find: " : "Unable to find file: ";
    if (!rtext.open(o.ref)) throw UserError(not_found + o.ref);
    if (o.num_files == 2 && !mtext.open(o.map)) throw UserError(not_found + o.map);
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
