// synth-bed -- host generator of the synthetic sorted BED the benchmark is quoted on: hg38 chromosome
// sizes in strcmp order, n_c = round(N * size_c / sum), start ~ U[0, size_c - 1), len = max(1, floor(LogNormal(mu, sigma))),
// end = min(start + len, size_c) forced > start, rows sorted by (start, end), columns chrom start end id<k> score with an
// integer score U[0,1000).  Same distributions as bedops_b200/synth.py and bench.py's device generator (not the same
// random stream: SURVEY 8d asks for the distributions, the checker is re-run on whatever is generated).  Used by
// bench.py's CPU arm, which must not load the CUDA library: plain C++, one thread per chromosome.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

static const struct { const char* name; uint64_t size; } kChrom[] = {
    {"chr1", 248956422},  {"chr10", 133797422}, {"chr11", 135086622}, {"chr12", 133275309}, {"chr13", 114364328},
    {"chr14", 107043718}, {"chr15", 101991189}, {"chr16", 90338345},  {"chr17", 83257441},  {"chr18", 80373285},
    {"chr19", 58617616},  {"chr2", 242193529},  {"chr20", 64444167},  {"chr21", 46709983},  {"chr22", 50818468},
    {"chr3", 198295559},  {"chr4", 190214555},  {"chr5", 181538259},  {"chr6", 170805979},  {"chr7", 159345973},
    {"chr8", 145138636},  {"chr9", 138394717},  {"chrX", 156040895},  {"chrY", 57227415}};
static const int kN = sizeof(kChrom) / sizeof(kChrom[0]);

static char* put_u64(char* p, uint64_t v) {
  char tmp[24];
  int  n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

int main(int argc, char** argv) {
  if (argc < 7) {
    std::fprintf(stderr, "usage: synth-bed <rows> <seed> <mu> <sigma> <fields 3|5> <out-file> [chrom ...]\n");
    return 2;
  }
  const uint64_t rows = std::strtoull(argv[1], nullptr, 10);
  const uint64_t seed = std::strtoull(argv[2], nullptr, 10);
  const double   mu = std::atof(argv[3]), sigma = std::atof(argv[4]);
  const int      fields = std::atoi(argv[5]);
  double         total = 0;
  for (int c = 0; c < kN; c++) total += (double)kChrom[c].size;
  std::vector<uint64_t> count(kN), base(kN + 1, 0);
  for (int c = 0; c < kN; c++) {
    bool wanted = argc == 7;
    for (int a = 7; a < argc; a++) wanted |= std::strcmp(argv[a], kChrom[c].name) == 0;
    count[c] = wanted ? (uint64_t)std::llround((double)rows * (double)kChrom[c].size / total) : 0;
    base[c + 1] = base[c] + count[c];
  }
  std::vector<std::string> text(kN);
  auto work = [&](int c) {
    const uint64_t m = count[c], size = kChrom[c].size;
    if (!m) return;
    std::mt19937_64                        rng(seed * 1000003ull + (uint64_t)c);
    std::lognormal_distribution<double>    len(mu, sigma);
    std::uniform_int_distribution<uint64_t> st(0, size - 2), sc(0, 999);
    std::vector<uint64_t> key(m);
    for (uint64_t i = 0; i < m; i++) {
      const uint64_t s = st(rng);
      uint64_t       l = (uint64_t)std::floor(len(rng));
      if (l < 1) l = 1;
      uint64_t e = std::min(s + l, size);
      if (e <= s) e = s + 1;
      key[i] = (s << 32) | e;
    }
    std::sort(key.begin(), key.end());
    std::string& out = text[c];
    out.resize(m * 64);
    char*        p = &out[0];
    const size_t nl = std::strlen(kChrom[c].name);
    for (uint64_t i = 0; i < m; i++) {
      std::memcpy(p, kChrom[c].name, nl);
      p += nl;
      *p++ = '\t';
      p = put_u64(p, key[i] >> 32);
      *p++ = '\t';
      p = put_u64(p, key[i] & 0xFFFFFFFFull);
      if (fields >= 5) {
        *p++ = '\t'; *p++ = 'i'; *p++ = 'd';
        p = put_u64(p, base[c] + i);
        *p++ = '\t';
        p = put_u64(p, sc(rng));
      }
      *p++ = '\n';
    }
    out.resize((size_t)(p - out.data()));
  };
  unsigned nt = std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  std::vector<std::thread> th;
  std::vector<int>         order(kN);
  for (int c = 0; c < kN; c++) order[c] = c;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return count[a] > count[b]; });  // big chromosomes first
  std::vector<int> next(1, 0);
  std::vector<std::vector<int>> mine(nt);
  for (int k = 0; k < kN; k++) mine[k % nt].push_back(order[k]);
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t]() {
      for (int c : mine[t]) work(c);
    });
  for (auto& t : th) t.join();
  FILE* f = std::fopen(argv[6], "wb");
  if (!f) {
    std::perror(argv[6]);
    return 1;
  }
  uint64_t written = 0;
  for (int c = 0; c < kN; c++) {
    if (text[c].size() && std::fwrite(text[c].data(), 1, text[c].size(), f) != text[c].size()) return 1;
    written += count[c];
  }
  std::fclose(f);
  std::printf("%llu\n", (unsigned long long)written);
  return 0;
}
