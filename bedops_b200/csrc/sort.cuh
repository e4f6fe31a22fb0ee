// sort.cuh -- device radix sort shared by sort-bed and the bedops operators that need an order the input does not have
#pragma once
#include "common.cuh"

namespace bk {

// Sorts n (u64 key, u32 value) pairs by key bits [0, nbits), least significant digit first, 8 bits per pass, stable.
// The two buffer pairs are swapped as needed: on return *keys / *vals point at the sorted data (vals may be NULL for a
// keys-only sort).  n < 2^32 - 1.
int radix_sort_pairs(bk_ctx* ctx, uint64_t** keys, uint32_t** vals, uint64_t** keys_alt, uint32_t** vals_alt, uint64_t n,
                     int nbits);
int finish_text(bk_ctx* ctx, char* d_out, uint64_t bytes, uint64_t rows, int on_device, bk_text* out);

}  // namespace bk
