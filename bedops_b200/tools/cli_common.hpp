// cli_common.hpp -- shared host-side plumbing of the drop-in command-line tools (bedmap, bedops,
// closest-features).  The tools keep the reference's argv grammar, stdout/stderr text and exit codes
//  and hand all compute to libbedkit.so through include/bedkit.h.  There is no CPU compute path:
// when the library cannot reach a B200 the tool fails with the library's message.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include "bedkit.h"

namespace cli {

static const char* kVersion = "2.4.26";  // interfaces/general-headers/suite/BEDOPS.Version.hpp:36-42
static const char* kCitation = "http://bioinformatics.oxfordjournals.org/content/28/14/1919.abstract";
static const char* kAuthors = "Shane Neph & Scott Kuehn";

struct UserError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void banner(FILE* f, const char* prog) {
  std::fprintf(f, "%s\n  citation: %s\n  version:  %s\n  authors:  %s\n", prog, kCitation, kVersion, kAuthors);
}

inline bool only_chars(const std::string& s, const char* set) { return s.find_first_not_of(set) == std::string::npos; }

// whole-file read ("-" = stdin).  Returns false if the file cannot be opened.
inline bool slurp(const std::string& name, std::vector<char>& buf) {
  int fd = name == "-" ? 0 : ::open(name.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  size_t      hint = 0;
  if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) hint = (size_t)st.st_size;
  buf.clear();
  buf.resize(hint ? hint : (1u << 20));
  size_t n = 0;
  while (true) {
    if (n == buf.size()) buf.resize(buf.size() * 2);
    ssize_t r = ::read(fd, buf.data() + n, buf.size() - n);
    if (r < 0) {
      if (fd) ::close(fd);
      return false;
    }
    if (r == 0) break;
    n += (size_t)r;
  }
  buf.resize(n);
  if (fd) ::close(fd);
  return true;
}

inline void write_all(const char* p, size_t n) {
  while (n) {
    size_t w = std::fwrite(p, 1, n, stdout);
    if (w == 0) break;
    p += w;
    n -= w;
  }
  std::fflush(stdout);
}

struct Engine {
  bk_ctx* ctx = nullptr;
  Engine() {
    const char* dev = std::getenv("BEDKIT_DEVICE");
    int         rc = bk_init(&ctx, dev ? std::atoi(dev) : 0);
    if (rc != BK_OK) throw std::runtime_error(bk_strerror(rc));
  }
  ~Engine() { bk_destroy(ctx); }
  [[noreturn]] void raise(int rc) const {
    const char* detail = bk_last_error(ctx);
    throw std::runtime_error(detail && *detail ? std::string(detail) : std::string(bk_strerror(rc)));
  }
  bk_bed* load(const std::vector<char>& text, int min_fields, unsigned cols) const {
    bk_bed* b = nullptr;
    int     rc = bk_load_bed(ctx, text.data(), text.size(), min_fields, cols, &b);
    if (rc != BK_OK) raise(rc);
    return b;
  }
};

}  // namespace cli
