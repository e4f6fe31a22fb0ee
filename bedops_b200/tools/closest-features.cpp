// closest-features -- drop-in command line for the B200 engine.  Mirrors the option grammar of
// applications/bed/closestfeats/src/Input.hpp:58-102 and replaces findDistances()
// (ClosestFeature.cpp:260-413) with bk_closest().
#include "cli_common.hpp"
#include "help_text.hpp"

namespace {
using cli::UserError;
struct Help {};
struct Version {};
struct NoInput {};

struct Options {
  bool ec = false, overlaps = true, closest = false, dist = false, no_ref = false;
  std::string delim = "|", chrom = "all", ref, query;
};

void require(bool ok, const std::string& msg) {
  if (!ok) throw UserError(msg);
}

Options parse_args(int argc, char** argv) {
  Options o;
  if (argc <= 1) throw NoInput();
  int  i = 1;
  bool outoption = false;
  while (i < argc) {
    std::string next = argv[i];
    if (next == "--help") throw Help();
    else if (next == "--version") throw Version();
    else if (next == "--ec" || next == "--header") o.ec = true;
    else if (next == "--no-overlaps") o.overlaps = false;
    else if (next == "--delim") {
      require(++i < argc, "No value given for --delim.");
      o.delim = argv[i];
    } else if (next == "--chrom") {
      require(++i < argc, "No value given for --chrome.");
      o.chrom = argv[i];
    } else if (next == "--closest" || next == "--shortest") {
      require(!outoption, "Multiple output options not allowed.");
      o.closest = true;
      outoption = true;
    } else if (next == "--dist") o.dist = true;
    else if (next == "--no-ref") o.no_ref = true;
    else {
      require(i + 2 == argc, "Unknown option: " + next + ".");
      break;
    }
    ++i;
  }
  require(i + 2 == argc, "Not enough input files given.");
  o.ref = argv[i++];
  o.query = argv[i];
  require(o.ref.find("--") != 0, "Option given where file expected: " + o.ref + ".");
  require(o.query.find("--") != 0, "Option given where file expected: " + o.query + ".");
  return o;
}

void usage(FILE* f) { std::fputs(kUsageClosest, f); }  // byte for byte the reference's text (help_text.hpp)
}  // namespace

int main(int argc, char** argv) {
  try {
    Options           o = parse_args(argc, argv);
    cli::Input rtext, qtext;
    if (!rtext.open(o.ref)) throw UserError("Unable to find file: " + o.ref);
    if (!qtext.open(o.query)) throw UserError("Unable to find file: " + o.query);
    if (cli::any_archive({&rtext, &qtext})) {
      cli::Engine eng;
      cli::unstarch_if_archive(eng, rtext);
      cli::unstarch_if_archive(eng, qtext);
    }
    bk_cfspec spec;
    bk_cfspec_default(&spec);
    spec.dist = o.dist;
    spec.closest = o.closest;
    spec.no_overlaps = !o.overlaps;
    spec.no_ref = o.no_ref;
    std::string delim = o.delim;
    if (delim == "\t" || delim == "\\t" || delim == "'\t'") delim = "\t";
    spec.delim = delim.c_str();
    spec.chrom = o.chrom.c_str();
    const unsigned hdr = o.ec ? BK_LOAD_HEADERS : 0;
    if (o.ec) {
      cli::ec_prepare(rtext);
      cli::ec_prepare(qtext);
      cli::Engine eng;
      cli::ec_check(eng, rtext, o.ref, 3, true, false);
      cli::ec_check(eng, qtext, o.query, 3, true, false);
    }
    auto run_one = [&](cli::Engine& eng, const std::vector<cli::Slice>& sl) {
      bk_bed* ref = eng.load(sl[0].ptr, sl[0].len, 3, BK_COL_LINE | hdr);
      bk_bed* qry = eng.load(sl[1].ptr, sl[1].len, 3, BK_COL_LINE | hdr);
      bk_text out;
      int     rc = bk_closest(eng.ctx, ref, qry, &spec, &out);
      if (rc != BK_OK) eng.raise(rc);
      std::string text(out.ptr ? out.ptr : "", out.len);
      bk_free_text(eng.ctx, &out);
      bk_free_bed(eng.ctx, ref);
      bk_free_bed(eng.ctx, qry);
      return text;
    };
    const int gpus = cli::gpus_requested();
    std::vector<std::vector<cli::Slice>> slices;
    if (gpus > 1 && o.chrom == "all") slices = cli::plan_slices({&rtext, &qtext}, gpus * 4);
    if (!slices.empty()) {
      cli::run_sharded(slices, run_one, gpus);
    } else {
      cli::Engine eng;
      std::string text = run_one(eng, {cli::Slice{rtext.data, rtext.size}, cli::Slice{qtext.data, qtext.size}});
      cli::write_all(text.data(), text.size());
      rtext.settle();
      qtext.settle();
      cli::finish_now(EXIT_SUCCESS);
    }
    return EXIT_SUCCESS;
  } catch (const Help&) {
    cli::banner(stdout, "closest-features");
    usage(stdout);
    return EXIT_SUCCESS;
  } catch (const Version&) {
    cli::banner(stdout, "closest-features");
    return EXIT_SUCCESS;
  } catch (const NoInput&) {
    cli::banner(stderr, "closest-features");
    usage(stderr);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "May use closest-features --help for more help.\n\nError: %s\n", e.what());
  }
  return EXIT_FAILURE;
}
