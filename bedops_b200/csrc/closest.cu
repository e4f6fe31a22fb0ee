// closest.cu -- closest-features.  Placeholder entry points until the kernel lands.
#include "common.cuh"
using namespace bk;

extern "C" void bk_cfspec_default(bk_cfspec* spec) {
  memset(spec, 0, sizeof(*spec));
  spec->delim = "|";
}

extern "C" int bk_closest(bk_ctx* ctx, const bk_bed* ref, const bk_bed* query, const bk_cfspec* spec, bk_text* out) {
  if (!ctx || !ref || !query || !spec || !out) return BK_ERR_ARG;
  return fail(ctx, BK_ERR_UNSUPPORTED, "closest-features is not implemented yet");
}
