#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference tools (BEDOPS v2.4.26) as the
# parity oracle and the CPU baseline.  Nothing under bedops_b200/ may call what this produces.
#
# Recipe (our own; the reference's make/configure system is NOT run):
#   * sources are compiled where they lie under $REF (default /root/reference); nothing is copied
#     into the repo and nothing is written outside oracle/_ref/;
#   * the three vendored dependency tarballs ($REF/third-party/*.tar.bz2: jansson-2.6, bzip2-1.0.6,
#     zlib-1.2.7) are unpacked into oracle/_ref/third-party/ and their library .c files compiled
#     directly with gcc (they are linked only for Starch archive I/O, which is off the hot path);
#   * the eight starch support objects and the three tools use the reference's own flags
#     (-O3 -std=c++11 -static; applications/bed/bedmap/src/Makefile:26-31).
# Output: oracle/_ref/bin/{bedops,bedmap,closest-features,sort-bed,bedextract,starch,unstarch}  (static binaries, so they
# run unchanged on the GPU box, where /root/reference does not exist).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
TP="$OUT/third-party"
OBJ="$OUT/obj"
BIN="$OUT/bin"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$REF/applications/bed" ]; then
  echo "build_ref.sh: reference tree not found at $REF (expected on the GPU box); keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$TP" "$OBJ" "$BIN"

# ---- third-party static libraries, compiled file by file -----------------------------------------
if [ ! -f "$TP/libthird.a" ]; then
  for t in jansson-2.6 bzip2-1.0.6 zlib-1.2.7; do
    [ -d "$TP/$t" ] || tar -xjf "$REF/third-party/$t.tar.bz2" -C "$TP"
  done
  mkdir -p "$OBJ/tp"
  (
    cd "$OBJ/tp"
    J="$TP/jansson-2.6/src"; B="$TP/bzip2-1.0.6"; Z="$TP/zlib-1.2.7"
    srcs=()
    for f in dump error hashtable hashtable_seed load memory pack_unpack strbuffer strconv utf value; do srcs+=("$J/$f.c"); done
    for f in blocksort huffman crctable randtable compress decompress bzlib; do srcs+=("$B/$f.c"); done
    for f in adler32 crc32 deflate infback inffast inflate inftrees trees zutil compress uncompr gzclose gzlib gzread gzwrite; do srcs+=("$Z/$f.c"); done
    printf '%s\n' "${srcs[@]}" | xargs -P "$JOBS" -I{} sh -c \
      'o=$(echo {} | tr "/." "__").o; gcc -O2 -w -D_LARGEFILE64_SOURCE=1 -DHAVE_STDINT_H=1 -I"'"$J"'" -I"'"$B"'" -I"'"$Z"'" -c {} -o $o'
    ar rcs "$TP/libthird.a" ./*.o
  )
fi
INC="-iquote$REF/interfaces/general-headers -I$TP -I$TP/jansson-2.6/src -I$TP/bzip2-1.0.6 -I$TP/zlib-1.2.7"
# the reference includes <jansson.h>, <bzlib.h>, <zlib.h> through third-party-relative paths as well
ln -sfn jansson-2.6 "$TP/jansson"; ln -sfn bzip2-1.0.6 "$TP/bzip2"; ln -sfn zlib-1.2.7 "$TP/zlib"
mkdir -p "$TP/jansson-2.6/include"; cp -f "$TP/jansson-2.6/src/jansson.h" "$TP/jansson-2.6/src/jansson_config.h" "$TP/jansson-2.6/include/"
INC="$INC -I$TP/jansson/include"

CXXFLAGS="-O3 -std=c++11 -w"
# ---- starch support objects (needed by every tool's input iterator) -----------------------------
deps=()
for n in starchConstants starchFileHelpers starchHelpers starchMetadataHelpers unstarchHelpers starchSha1Digest starchBase64Coding; do
  deps+=("$OBJ/$n.o")
  [ -f "$OBJ/$n.o" ] || echo "$REF/interfaces/src/data/starch/$n.c $OBJ/$n.o"
done > "$OBJ/todo.txt"
deps+=("$OBJ/NaN.o")
[ -f "$OBJ/NaN.o" ] || echo "$REF/interfaces/src/data/measurement/NaN.cpp $OBJ/NaN.o" >> "$OBJ/todo.txt"
if [ -s "$OBJ/todo.txt" ]; then
  xargs -P "$JOBS" -L1 sh -c 'g++ '"$CXXFLAGS $INC"' -c "$0" -o "$1"' < "$OBJ/todo.txt"
fi

build_tool() { # name, source dir, main source(s)
  local name="$1" dir="$2"; shift 2
  if [ ! -x "$BIN/$name" ]; then
    ( cd "$dir" && g++ -static -s $CXXFLAGS $INC -o "$BIN/$name" "$@" "${deps[@]}" "$TP/libthird.a" )
    echo "built $BIN/$name"
  fi
}
A="$REF/applications/bed"
build_tool bedops           "$A/bedops/src"       Bedops.cpp &
build_tool closest-features "$A/closestfeats/src" ClosestFeature.cpp &
build_tool bedextract       "$A/bedextract/src"   ExtractRows.cpp &
build_tool sort-bed         "$A/sort-bed/src"     Sort.cpp SortDetails.cpp CheckSort.cpp &
build_tool bedmap           "$A/bedmap/src"       Bedmap.cpp &
# starch / unstarch: C sources the reference compiles with its C++ compiler (applications/bed/starch/src/Makefile:67);
# they make and read the archives the Starch-input tests use
build_c_tool() { # name, source
  if [ ! -x "$BIN/$1" ]; then
    g++ -static -s $CXXFLAGS -D__STDC_CONSTANT_MACROS -D_FILE_OFFSET_BITS=64 -D_LARGEFILE64_SOURCE=1 -DUSE_ZLIB -DUSE_BZLIB \
        -x c++ $INC -o "$BIN/$1" "$2" -x none "${deps[@]}" "$TP/libthird.a"
    echo "built $BIN/$1"
  fi
}
build_c_tool starch   "$A/starch/src/starch.c" &
build_c_tool unstarch "$A/starch/src/unstarch.c" &
wait
ls -la "$BIN"
