// sort-bed -- drop-in command line for the B200 engine.  Mirrors applications/bed/sort-bed/src/Sort.cpp:44-234 (argv
// grammar, banners, messages) and replaces processData / lexSortBedData / printBed (SortDetails.cpp:530-1208) with
// bk_sort_bed(), checkSort (CheckSort.cpp:35-56) with bk_check_text().  --max-mem / --tmpdir select the reference's
// external merge sort; the device sorts in HBM (180 GB), so both are accepted and have no effect.
#include "cli_common.hpp"
#include <cctype>

namespace {
const char* kName = "sort-bed";
const char* kSortAuthors = "Scott Kuehn";
const char* kUsage =
    "\nUSAGE: sort-bed [--help] [--version] [--check-sort] [--max-mem <val>] [--tmpdir <path>] <file1.bed> <file2.bed> <...>\n"
    "        Sort BED file(s).\n        May use '-' to indicate stdin.\n        Results are sent to stdout.\n\n"
    "        <val> for --max-mem may be 8G, 8000M, or 8000000000 to specify 8 GB of memory.\n"
    "        --tmpdir is useful only with --max-mem.\n";

void banner(FILE* f) {
  std::fprintf(f, "%s\n  citation: %s\n  version:  %s\n  authors:  %s\n", kName, cli::kCitation, cli::kVersion, kSortAuthors);
}
[[noreturn]] void die(const char* msg) {
  std::fputs(msg, stderr);
  std::exit(EXIT_FAILURE);
}

struct Options {
  std::vector<std::string> files;
  bool                     check = false;
};

// Sort.cpp:44-190
Options get_args(int argc, char** argv) {
  Options o;
  int     num_files = argc - 1;
  if (num_files < 1) {
    banner(stderr);
    std::fprintf(stderr, "%s\n", kUsage);
    std::exit(EXIT_FAILURE);
  }
  if (num_files > 10000) {
    banner(stderr);
    std::fprintf(stderr, "%s\nToo Many Files\n", kUsage);
    std::exit(EXIT_FAILURE);
  }
  int  stdincnt = 0;
  bool change_mem = false, change_tdir = false;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--help") {
      banner(stdout);
      std::fprintf(stdout, "%s\n", kUsage);
      std::exit(EXIT_SUCCESS);
    } else if (a == "--version") {
      banner(stdout);
      std::exit(EXIT_SUCCESS);
    } else if (a == "--max-mem") {
      if (change_mem) die("Specify --max-mem at most one time!\n");
      change_mem = true;
      if (++i == argc) die("No value given for --max-mem.\n");
      const std::string v = argv[i];
      size_t            lng = v.size();
      double            factor = 1, max_mem = 0;
      bool              units = false;
      for (size_t k = 0; k < lng; ++k) {
        if (!std::isdigit((unsigned char)v[k])) {
          if (k == 0 || k != lng - 1)
            die("Bad number for --max-mem.  Expect value to be like 10G (for 10 gigabytes) or 1000M (for 1000 megabytes) or just "
                "1000000000 (for 1 gigabyte).\n");
          if (v[k] == 'G') factor = 1000000000, --lng;
          else if (v[k] == 'M') factor = 1000000, --lng;
          else
            die("Unrecognized units for --max-mem.  Expect value to be like 10G (for 10 gigabytes) or 1000M (for 1000 megabytes) or "
                "just 1000000000 (for 1 gigabyte).\n");
          units = true;
          max_mem = factor * std::strtod(v.substr(0, lng).c_str(), nullptr);
        }
      }
      if (!units) max_mem = std::strtod(v.c_str(), nullptr);
      if (max_mem > 128000000000.0)
        std::fputs("\nSetting memory > 128 GB probably isn't practical.\nIf you remove --max-mem, the program will use up to all "
                   "available system memory.\nContinuing.\n\n",
                   stderr);
      if (max_mem < 500000000.0)
        die("While theoretically possible to sort with less memory, we expect at least 500 megabytes for --max-mem\n");
      num_files -= 2;
      continue;
    } else if (a == "--tmpdir") {
      if (change_tdir) die("Specify --tmpdir at most one time!\n");
      change_tdir = true;
      if (++i == argc) die("No value given for --tmpdir.\n");
      num_files -= 2;
      continue;
    } else if (a == "--check-sort") {
      o.check = true;
      num_files -= 1;
      continue;
    } else if (a == "-") {
      stdincnt++;
    }
    o.files.push_back(a);
  }
  if (stdincnt > 1) die("Cannot specify '-' more than once\n");
  if (num_files < 1) {
    banner(stderr);
    std::fprintf(stderr, "%s\n%s\n", kUsage, "No file given.");
    std::exit(EXIT_FAILURE);
  }
  return o;
}

bool sep(char c) { return c == '\t' || c == ' '; }
bool starts_with(const char* p, const char* e, const char* w) {
  const size_t n = std::strlen(w);
  return (size_t)(e - p) >= n && std::memcmp(p, w, n) == 0;
}

// The reference's message for a line it rejects (SortDetails.cpp:631-779, :833-853), "" if it takes the line.
// [p, e) is the line without its NL; has_nl tells whether the file had one there.
std::string explain(const char* p, const char* e, bool has_nl, unsigned long long line, const std::string& file) {
  auto at = [&](const char* fmt) {
    char buf[512];
    std::snprintf(buf, sizeof buf, fmt, line, file.c_str());
    return std::string(buf);
  };
  if (p < e && sep(*p)) return at("Row begins with a tab or space at line %llu in %s.\n");
  const char* c = p;
  while (c < e && !sep(*c)) c++;
  if (c == e) return at("No tabs/spaces found at line %llu in %s.\n");
  if (c - p > 127)
    return at("Chromosome name too long at line %llu in %s.\n") +
           "Check that you have unix newlines (cat -A) or increase TOKEN_CHR_MAX_LENGTH in BEDOPS.Constants.hpp and recompile BEDOPS.\n";
  const char* s0 = c + 1;
  const char* d = s0;
  while (d < e && !sep(*d)) d++;
  if (d == e)
    return at("No tabs/spaces found after the start coordinate (or no start coordinate at all) at line %llu in %s.\n");
  if (d - s0 > 12) return at("Start coordinate is too large.  Max decimal digits allowed is 12 in BEDOPS.Constants.hpp.  See line %llu in %s.\n");
  if (d == s0) return at("Consecutive tabs and/or spaces between chromosome and start coordinate.  See line %llu in %s.\n");
  for (const char* k = s0; k < d; k++)
    if (!std::isdigit((unsigned char)*k))
      return at("Non-numeric start coordinate.  See line %llu in %s.\n(remember that chromosome names should not contain spaces.)\n");
  const char* e0 = d + 1;
  const char* f = e0;
  while (f < e && !sep(*f)) f++;
  if (f == e && !has_nl)
    return at("No end of line found at %llu in %s.\nMay need to increase BED_LINE_LEN and recompile.\nFirst check that you have unix newlines (cat -A).");
  if (f - e0 > 12) return at("End coordinate is too large.  Max decimal digits allowed is 12 in BEDOPS.Constants.hpp.  See line %llu in %s.\n");
  if (f == e0) return at("Extra tab and/or space found in between start and end coordinates.  See line %llu in %s.\n");
  for (const char* k = e0; k < f; k++)
    if (!std::isdigit((unsigned char)*k)) return at("Non-numeric end coordinate.  See line %llu in %s.\n");
  const unsigned long long st = std::strtoull(std::string(s0, d).c_str(), nullptr, 10), en = std::strtoull(std::string(e0, f).c_str(), nullptr, 10);
  if (en <= st) return at("Error on line %llu in %s. Genomic end coordinate is less than (or equal to) start coordinate.\n");
  const char* q = f;
  while (q < e && std::isspace((unsigned char)*q)) q++;
  if (q < e) {
    const char* k = q;
    while (k < e && !sep(*k)) k++;
    if (k - q > 16383)
      return at("ID field too long at line %llu in %s.\n") +
             "Check that you have unix newlines (cat -A) or increase TOKEN_ID_MAX_LENGTH in BEDOPS.Constants.hpp and recompile BEDOPS.\n"
             "You may instead choose to put a dummy id column (like 'id') in as the 4th field to fix this.\n";
  }
  return "";
}

struct Piece {  // a stretch of one input file inside the concatenated text
  size_t             off, len;
  int                file;
  unsigned long long first_line;  // line number (1-based) of the stretch's first line in its file
  bool               added_nl;    // the file's last line had no NL: one was appended
};
}  // namespace

int main(int argc, char** argv) {
  const Options o = get_args(argc, argv);
  try {
    if (o.check) {  // CheckSort.cpp:35-56
      try {
        std::unique_ptr<cli::Engine> eng;  // made when the first file is there: a missing file is reported without touching the GPU
        for (const std::string& name : o.files) {
          cli::Input in;
          if (!in.open(name)) throw std::runtime_error("Unable to find: " + name);
          if (!eng) eng.reset(new cli::Engine());
          cli::ec_prepare(in);
          cli::ec_check(*eng, in, name, 3, true, false);
        }
        return EXIT_SUCCESS;
      } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return EXIT_FAILURE;
      }
    }
    // checkfiles, SortDetails.cpp:359-387: a pass of its own over all the names before anything is read -- whatever fopen
    // opens passes (a directory does), stdin may be named once
    {
      int stdin_count = 0;
      for (const std::string& name : o.files) {
        if (name != "-") {
          const int fd = ::open(name.c_str(), O_RDONLY);
          if (fd < 0) {
            std::fprintf(stderr, "Unable to access %s\n", name.c_str());
            return EXIT_FAILURE;
          }
          ::close(fd);
        } else if (++stdin_count > 1) {
          std::fprintf(stderr, "stdin specified multiple times\n");
          return EXIT_FAILURE;
        }
      }
    }
    std::vector<std::unique_ptr<cli::Input>> inputs;
    for (const std::string& name : o.files) {
      inputs.emplace_back(new cli::Input());
      if (!inputs.back()->open(name)) {
        std::fprintf(stderr, "Unable to access %s\n", name.c_str());
        return EXIT_FAILURE;
      }
    }
    // header lines at the top of every file are dropped (SortDetails.cpp:645-653; empty lines do not end the header
    // zone, :625-629); what follows travels to the device as one text
    std::vector<Piece> pieces;
    std::vector<char>  joined;
    const char*        text = nullptr;
    size_t             size = 0;
    {
      struct Span {
        const char* p;
        size_t      n;
      };
      std::vector<Span> spans;
      for (size_t f = 0; f < inputs.size(); f++) {
        const char*        p = inputs[f]->data;
        const char* const  e = p + inputs[f]->size;
        unsigned long long line = 1;
        while (p < e) {
          const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(e - p)));
          const char* le = nl ? nl : e;
          const bool  header = le > p && !sep(*p) &&
                              (starts_with(p, le, "browser") || starts_with(p, le, "track") || *p == '#' || *p == '@');
          if (le != p && !header) break;
          p = nl ? nl + 1 : e;
          line++;
        }
        if (p == e) continue;
        const bool add_nl = e[-1] != '\n';
        pieces.push_back(Piece{0, (size_t)(e - p) + (add_nl ? 1 : 0), (int)f, line, add_nl});
        spans.push_back(Span{p, (size_t)(e - p)});
      }
      if (spans.size() == 1 && !pieces[0].added_nl) {
        text = spans[0].p;
        size = spans[0].n;
      } else {
        size_t total = 0;
        for (const Piece& pc : pieces) total += pc.len;
        joined.resize(total);
        size_t at = 0;
        for (size_t k = 0; k < spans.size(); k++) {
          pieces[k].off = at;
          std::memcpy(joined.data() + at, spans[k].p, spans[k].n);
          at += spans[k].n;
          if (pieces[k].added_nl) joined[at++] = '\n';
        }
        text = joined.data();
        size = total;
      }
    }
    // a last line without NL and without a rest is an error for the reference ("No end of line found"); it comes
    // after every earlier error, so the text before it is validated first
    auto report = [&](uint64_t off) -> bool {  // true if the reference rejects the line at `off`
      size_t k = 0;
      while (k + 1 < pieces.size() && pieces[k + 1].off <= off) k++;
      const Piece&       pc = pieces[k];
      unsigned long long line = pc.first_line;
      for (const char* q = text + pc.off; q < text + off;) {
        const char* nl = static_cast<const char*>(std::memchr(q, '\n', (size_t)(text + off - q)));
        if (!nl) break;
        line++;
        q = nl + 1;
      }
      const char* le = static_cast<const char*>(std::memchr(text + off, '\n', size - off));
      const bool  has_nl = !(pc.added_nl && le + 1 == text + pc.off + pc.len);
      const std::string msg = explain(text + off, le, has_nl, line, o.files[pc.file]);
      if (msg.empty()) return false;
      std::fputs(msg.c_str(), stderr);
      return true;
    };
    size_t   limit = size;  // sort [0, limit)
    uint64_t tail_err = ~0ull;
    for (const Piece& pc : pieces) {
      if (!pc.added_nl) continue;
      size_t ls = pc.off + pc.len - 1;  // the appended NL
      while (ls > pc.off && text[ls - 1] != '\n') ls--;
      const char* le = text + pc.off + pc.len - 1;
      if (!explain(text + ls, le, false, 0, "").empty() && explain(text + ls, le, true, 0, "").empty()) {
        tail_err = ls;
        limit = ls;
        break;
      }
    }
    cli::Engine eng;
    bk_text     out{};
    uint64_t    bad = ~0ull;
    int         rc = bk_sort_bed(eng.ctx, text, limit, 0, &out, &bad);
    if ((rc == BK_ERR_PARSE || rc == BK_ERR_COORD_RANGE) && bad != ~0ull) {
      if (report(bad)) return EXIT_FAILURE;
      eng.raise(rc);  // a row the reference takes and this build does not (coordinate beyond 32 bits): say so
    }
    if (rc != BK_OK) eng.raise(rc);
    if (tail_err != ~0ull) {
      bk_free_text(eng.ctx, &out);
      report(tail_err);
      return EXIT_FAILURE;
    }
    cli::write_all(out.ptr ? out.ptr : "", out.len);
    for (auto& in : inputs) in->settle();
    cli::finish_now(EXIT_SUCCESS);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
  }
  return EXIT_FAILURE;
}
