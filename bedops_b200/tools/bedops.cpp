// bedops -- drop-in command line for the B200 engine.  Mirrors the option grammar, output and exit codes of
// applications/bed/bedops/src/Input.hpp:57-298 / Bedops.cpp:81-127 for the operators on the hot path
// (--merge, --intersect, --element-of, --not-element-of) and replaces selectWork() (Bedops.cpp:1523-1577)
// with bk_setop().
#include <fstream>
#include <map>
#include <sstream>
#include "cli_common.hpp"
#include "help_text.hpp"

namespace {

using cli::UserError;

// `stringstream >> int` as Input.hpp:237, :248 use it: an empty word leaves the previous value, an overflowing one saturates
int as_the_reference_reads_an_int(const std::string& word, int previous) {
  std::stringstream s;
  s << word;
  s >> previous;
  return previous;
}

enum Mode { MERGE, INTERSECTION, ELEMENTOF, NOTELEMENTOF, COMPLEMENT, DIFFERENCE, SYMMDIFF, UNIONALL, PARTITION, CHOP };

struct Options {
  Mode   mode = MERGE;
  int    min_files = 1000;
  double subset = 1;  // Input::subsetPerc_
  bool   use_pct = true;
  bool   ec = false, has_range = false, full_left = false;
  long   chop_bp = 1, chop_stagger = 0;  // Input::chopBP_, chopStaggerBP_
  bool   chop_cut_short = false;
  int    lpad = 0, rpad = 0;  // Input::lpad_, rpad_
  std::string chrom = "all";
  std::vector<std::string> files;
};

void require(bool ok, const std::string& msg) {
  if (!ok) throw UserError(msg);
}

int stream_int(const std::string& t) {  // what "std::stringstream(t) >> int" leaves in an int initialised to 0
  std::stringstream conv(t);
  int               v = 0;
  conv >> v;
  return v;
}

struct Help {
  std::string which;  // "--help" or "--help-<operation>"
};
struct Version {};
struct NoInput {};

void set_mode(Options& o, char t) {  // Input::setModeType, Input.hpp:390-420
  int min = 1;
  switch (t) {
    case 'c': case 'C': o.mode = COMPLEMENT; break;
    case 'd': case 'D': o.mode = DIFFERENCE; ++min; break;
    case 'e': case 'E': o.mode = ELEMENTOF; ++min; break;
    case 'i': case 'I': o.mode = INTERSECTION; ++min; break;
    case 'm': case 'M': o.mode = MERGE; break;
    case 'n': case 'N': o.mode = NOTELEMENTOF; ++min; break;
    case 'p': case 'P': o.mode = PARTITION; break;
    case 's': case 'S': o.mode = SYMMDIFF; ++min; break;
    case 'u': case 'U': o.mode = UNIONALL; break;
    case 'w': case 'W': o.mode = CHOP; break;
    default: throw UserError(std::string("Unknown operation: -") + t);
  }
  o.min_files = min;
}

void set_subset(Options& o, const std::string& str) {  // Input::setSubsetOption, Input.hpp:344-382
  const char* nums = ".1234567890";
  const char* ints = "1234567890";
  std::string l = str.substr(1);
  std::string::size_type pos = str.find("%");
  if (pos != std::string::npos) {
    require(pos + 1 == str.size(), "Bad placement of %");
    std::string value = str.substr(0, pos);
    require(!value.empty(), "Bad % value");
    if (value[0] == '-') {
      value = value.substr(1);
      require(!value.empty(), "Bad % value");
    }
    require(cli::only_chars(value, nums), "Bad: % value");
    o.subset = std::strtod(value.c_str(), nullptr);
    o.subset /= 100.0;
    if (o.subset > 1) throw UserError("Expect percentage less than or equal to 100%");
    o.use_pct = true;
    if (o.subset == 0) {  // 0% can match everything: convert to 1bp
      o.subset = 1;
      o.use_pct = false;
    }
  } else if (cli::only_chars(str, ints)) {
    o.subset = std::atoi(str.c_str());
    o.use_pct = false;
  } else if (cli::only_chars(l, ints)) {
    o.subset = std::atoi(l.c_str());
    o.use_pct = false;
  } else if (cli::only_chars(str, nums)) {
    throw UserError("Fractional amounts require a '%' symbol (e.g.; 5.4% not 5.4 base-pair)");
  } else {
    throw UserError("Unknown arg: " + str);
  }
}

Options parse_args(int argc, char** argv) {
  Options o;
  const std::map<std::string, std::string> longopts = {
      {"--complement", "-c"}, {"--difference", "-d"}, {"--element-of", "-e"},     {"--intersect", "-i"},  {"--merge", "-m"},
      {"--not-element-of", "-n"}, {"--partition", "-p"}, {"--symmdiff", "-s"}, {"--everything", "-u"}, {"--chop", "-w"}};
  try {
    if (argc <= 1) throw NoInput();
    bool has_option = false;
    int  i = 1;
    bool chr_specific = false;
    const char* plusints = "0123456789";
    while (i < argc) {
      std::string next = argv[i];
      if (next == "--ec" || next == "--header") {
        o.ec = true;
      } else if (next == "--chrom") {
        require(!chr_specific, "--chrom specified multiple times.");
        require(++i < argc, "No value for --chrom given.");
        o.chrom = argv[i];
        chr_specific = o.chrom != "all";
      } else if (next == "--range") {
        require(!o.has_range, "--range specified multiple times.");
        require(++i < argc, "No value for --range given.");
        const std::string v = argv[i];  // Input.hpp:89-126
        const std::string ok = std::string("-") + plusints;
        auto one_minus = [](const std::string& t) { return t.find_first_of("-") == t.find_last_of("-"); };
        if (v.find(":") != std::string::npos) {
          const std::string l = v.substr(0, v.find(":")), r = v.substr(v.find(":") + 1);
          require(!l.empty(), "integer expected for the 'L' value of --range L:R.");
          require(!r.empty(), "integer expected for the 'R' value of --range L:R.");
          require(l.find_first_not_of(ok) == std::string::npos, "integer expected for the 'L' value of --range L:R.");
          require(r.find_first_not_of(ok) == std::string::npos, "integer expected for the 'R' value of --range L:R.");
          require(one_minus(l), "multiple '-' signs detected for 'L' value of --range option");
          require(one_minus(r), "multiple '-' signs detected for 'R' value of --range option");
          o.lpad = stream_int(l);
          o.rpad = stream_int(r);
        } else {
          require(v.find_first_not_of(ok) == std::string::npos, "integer value expected for --range");
          require(one_minus(v), "multiple '-' signs detected in <val> for --range option");
          const int range = stream_int(v);
          o.lpad = -range;
          o.rpad = range;
        }
        o.has_range = true;
      } else if (next == "--help") {
        throw Help{next};
      } else if (next.find("--help-") == 0) {
        throw Help{next};
      } else if (next == "--version") {
        throw Version();
      } else if (next.find("-") != 0) {
        break;
      } else if (next.size() > 1) {
        require(next.find_first_not_of("-") != std::string::npos, "Bad option: " + next);
        require(!has_option, "More than one operation specified: " + next);
        has_option = true;
        if (next.find("--") == 0) {
          auto it = longopts.find(next);
          require(it != longopts.end(), "Unknown operation: " + next);
          next = it->second;
        }
        require(next.size() == 2, "Unknown operation: " + next);
        set_mode(o, next[1]);
        if (o.mode == ELEMENTOF || o.mode == NOTELEMENTOF) {  // optional overlap spec, Input.hpp:171-206
          const char* ints = "1234567890";
          if (i + 1 < argc) {
            std::string a = argv[i + 1];
            if ((a[0] == '-' && a.size() > 1) || cli::only_chars(a, ints) || a.find("%") != std::string::npos) {
              if (a.find("--") == std::string::npos) {
                std::ifstream tmpfile(a.c_str());
                if (!tmpfile) {
                  set_subset(o, a);
                  ++i;
                } else if (cli::only_chars(a, plusints)) {
                  std::fprintf(stderr,
                               "Warning: interpreting argument '%s' as a file input and not as an overlap spec,\n"
                               "         since the file exists.\n"
                               "You can use the legacy syntax '-%s' if you want to use it as an overlap criterion.\n",
                               a.c_str(), a.c_str());
                }
              }
            }
          }
        } else if (o.mode == CHOP) {  // [chunk] [--stagger nt] [-x], Input.hpp:221-258
          const char* ints = "1234567890";
          bool value_set = false, aux_set = false, stagger_set = false;
          int  cntr = 0;
          while (i + 1 < argc) {
            std::string a = argv[i + 1];
            if (a == "--stagger") {
              require(!stagger_set, "chop's --stagger suboption specified multiple times.");
              require(i + 2 < argc, "No #nt value found for --stagger suboption in --chop");
              std::string v = argv[i + 2];
              require(cli::only_chars(v, ints), "Invalid --stagger suboption #nt value in --chop.  Expect a +integer.");
              o.chop_stagger = as_the_reference_reads_an_int(v, (int)o.chop_stagger);
              require(o.chop_stagger > 0, "nt setting for chop's --stagger suboption must be > 0");
              stagger_set = aux_set = true;
              i += 2;
            } else if (a == "-x") {
              require(!o.chop_cut_short, "chop's -x suboption specified multiple times.");
              o.chop_cut_short = aux_set = true;
              i += 1;
            } else if (cli::only_chars(a, ints)) {  // an empty word passes this test too, and leaves the default in place
              require(!value_set, "Stray integer found (invalid argument for --chop?)");
              require(!aux_set, "Stray integer value found: not valid for --chop");
              o.chop_bp = as_the_reference_reads_an_int(a, (int)o.chop_bp);
              require(o.chop_bp > 0, "bp setting for chop must be > 0");
              value_set = true;
              i += 1;
            } else {
              break;
            }
            ++cntr;
          }
          require(cntr <= 4, "Too many arguments for a --chop operation");
        } else if (o.mode == COMPLEMENT) {
          while (i + 1 < argc && std::string(argv[i + 1]) == "-L") {
            o.full_left = true;
            ++i;
          }
        }
      } else {
        break;
      }
      ++i;
    }
    require(i < argc, "No input file given.");
    require(has_option, "No operation argument given.");
    bool only_one = true;
    int  nfiles = 0;
    for (; i < argc; ++i) {
      std::string a = argv[i];
      if (a == "-") {
        require(only_one, "Too many '-'");
        only_one = false;
      } else {
        require(a[0] != '-', "Bad option: " + a);
        std::ifstream check(a.c_str());
        require(static_cast<bool>(check), "Cannot find " + a);
      }
      o.files.push_back(a);
      ++nfiles;
    }
    require(nfiles >= o.min_files, "Not enough files");
  } catch (const UserError& e) {
    throw UserError(std::string("Bad Input\n") + e.what());  // Input.hpp:292-297
  }
  return o;
}

void usage(FILE* f) { std::fputs(kUsageBedops, f); }  // byte for byte the reference's text (help_text.hpp)

}  // namespace

int main(int argc, char** argv) {
  try {
    Options o = parse_args(argc, argv);
    int     op = 0;
    switch (o.mode) {
      case MERGE: op = BK_SETOP_MERGE; break;
      case INTERSECTION: op = BK_SETOP_INTERSECT; break;
      case ELEMENTOF: op = BK_SETOP_ELEMENT_OF; break;
      case NOTELEMENTOF: op = BK_SETOP_NOT_ELEMENT_OF; break;
      case COMPLEMENT: op = BK_SETOP_COMPLEMENT; break;
      case DIFFERENCE: op = BK_SETOP_DIFFERENCE; break;
      case SYMMDIFF: op = BK_SETOP_SYMMDIFF; break;
      case UNIONALL: op = BK_SETOP_EVERYTHING; break;
      case PARTITION: op = BK_SETOP_PARTITION; break;
      case CHOP: op = -1; break;  // bk_chop
    }
    std::vector<cli::Input> texts(o.files.size());
    for (size_t f = 0; f < o.files.size(); f++)
      if (!texts[f].open(o.files[f])) throw UserError("Cannot find " + o.files[f]);
    {
      bool archive = false;
      for (auto& t : texts) archive |= bk_is_starch(t.data, t.size) != 0;
      if (archive) {
        cli::Engine eng;
        for (auto& t : texts) cli::unstarch_if_archive(eng, t);
      }
    }
    const bool has_ref = op == BK_SETOP_ELEMENT_OF || op == BK_SETOP_NOT_ELEMENT_OF;
    const bool all_lines = op == BK_SETOP_EVERYTHING;  // every row of every file is echoed
    const unsigned hdr = o.ec ? BK_LOAD_HEADERS : 0;
    if (o.ec) {  // Bedops.cpp:259-286: BedPadReader over bed_check_iterator; B3Rest for the -e/-n reference, B3NoRest otherwise
      cli::Engine eng;
      for (size_t f = 0; f < texts.size(); f++) {
        cli::ec_prepare(texts[f]);
        cli::ec_check(eng, texts[f], o.files[f], 3, (has_ref && f == 0) || all_lines, false);
      }
    }
    auto run_one = [&](cli::Engine& eng, const std::vector<cli::Slice>& sl) {
      std::vector<bk_bed*> beds;
      for (size_t f = 0; f < sl.size(); f++) beds.push_back(eng.load(sl[f].ptr, sl[f].len, 3, (((has_ref && f == 0) || all_lines) ? BK_COL_LINE : 0) | hdr));
      std::vector<bk_bed*> plain;  // --range: the operators see the padded view; -e/-n leave the reference file alone (Bedops.cpp:220-236)
      if (o.has_range && (o.lpad != 0 || o.rpad != 0))
        for (size_t f = 0; f < beds.size(); f++) {
          if (has_ref && f == 0) continue;
          bk_bed* padded = nullptr;
          int     prc = bk_bed_pad(eng.ctx, beds[f], o.lpad, o.rpad, &padded);
          if (prc != BK_OK) eng.raise(prc);
          plain.push_back(beds[f]);
          beds[f] = padded;
        }
      bk_text out;
      const double thr = op == BK_SETOP_COMPLEMENT ? (o.full_left ? 1.0 : 0.0) : o.subset;
      int     rc = op < 0 ? bk_chop(eng.ctx, beds.data(), (int)beds.size(), (uint64_t)o.chop_bp, (uint64_t)o.chop_stagger,
                                    o.chop_cut_short ? 1 : 0, o.chrom.c_str(), 0, &out)
                          : bk_setop(eng.ctx, op, beds.data(), (int)beds.size(), thr, o.use_pct ? 1 : 0, o.chrom.c_str(), 0, &out);
      if (rc != BK_OK) eng.raise(rc);
      std::string text(out.ptr ? out.ptr : "", out.len);
      bk_free_text(eng.ctx, &out);
      for (bk_bed* b : beds) bk_free_bed(eng.ctx, b);
      for (bk_bed* b : plain) bk_free_bed(eng.ctx, b);
      return text;
    };
    const int gpus = cli::gpus_requested();
    // "-e 0" / "-n 0" look at later chromosomes (Bedops.cpp:1044-1049): keep those runs on one GPU
    std::vector<std::vector<cli::Slice>> slices;
    // --range: what the reader does around zero depends on the position in the FILE (BedPadReader's constructor): one GPU
    if (gpus > 1 && o.chrom == "all" && !o.has_range && !(has_ref && !o.use_pct && o.subset <= 0)) {
      std::vector<const cli::Input*> files;
      for (auto& t : texts) files.push_back(&t);
      slices = cli::plan_slices(files, gpus * 4);
    }
    if (!slices.empty()) {
      cli::run_sharded(slices, run_one, gpus);
    } else {
      cli::Engine eng;
      std::vector<cli::Slice> sl;
      for (auto& t : texts) sl.push_back(cli::Slice{t.data, t.size});
      std::string text = run_one(eng, sl);
      cli::write_all(text.data(), text.size());
      for (auto& t : texts) t.settle();
      cli::finish_now(EXIT_SUCCESS);
    }
    return EXIT_SUCCESS;
  } catch (const Help& h) {
    cli::banner(stdout, "bedops");
    const char* text = kUsageBedops;
    for (const BedopsHelp* e = kBedopsHelp; e->op; e++)
      if (h.which == e->op) text = e->text;
    std::fputs(text, stdout);
    return EXIT_SUCCESS;
  } catch (const Version&) {
    cli::banner(stdout, "bedops");
    return EXIT_SUCCESS;
  } catch (const NoInput&) {
    cli::banner(stdout, "bedops");  // the reference sends the banner to stdout and the usage text to stderr
    std::fflush(stdout);
    usage(stderr);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "May use bedops --help for more help.\n\nError: %s\n", e.what());
  }
  return EXIT_FAILURE;
}
