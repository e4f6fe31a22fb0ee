#!/usr/bin/env bash
# Builds tuning variants of libbedkit.so next to the product library (bedops_b200/lib/variants/<name>.so); bench.py /
# the tests pick one with BEDKIT_LIB=<path>.  usage: build_variants.sh name "extra nvcc flags" [name flags ...]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/../.." && pwd)"
SRC="$HERE/bedops_b200/csrc"
OUT="$HERE/bedops_b200/lib/variants"
mkdir -p "$OUT"
while [ $# -ge 2 ]; do
  name="$1"; flags="$2"; shift 2
  obj="$HERE/bedops_b200/build/var_$name"
  mkdir -p "$obj"
  for f in api parse bedmap setops closest hostplan check pipeline shard sort pad starch; do
    ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
        $flags -c "$SRC/$f.cu" -o "$obj/$f.o" ) &
  done
  wait
  /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/$name.so" "$obj"/*.o -cudart static -lz -ldl
  echo "built $OUT/$name.so"
done
