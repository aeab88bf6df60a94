#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE (not product code).
#
# Compiles the UNMODIFIED reference (isinaltinkaya/vcfgl @ da6a334 + its
# vendored htslib 1.15.1) from the sources where they lie under $REF
# (default /root/reference, read-only) into oracle/_ref/:
#
#   oracle/_ref/vcfgl_ref        the reference CLI, unmodified
#   oracle/_ref/vcfgl_ref_dump   same + replay-capture hooks (ref_dump_hooks.h)
#   oracle/_ref/vcfgl_ref_vgl    same + the hot path answered by libvgl.so (ref_vgl_binding.h; INTEGRATION.md compiled)
#   oracle/_ref/libref_errmod.so htslib errmod.c alone (errmod_init/errmod_cal)
#
# The reference's own build systems are NOT run (its Makefile does `git
# submodule update`; htslib's default config.h assumes bz2/lzma/curl which are
# absent here).  We call gcc/g++ on the source files directly with three
# hand-written generated headers: config.h, version.h
# (HTS_VERSION_TEXT + HTSCODECS_VERSION_TEXT + VCFGL_VERSION) and build.h.
# No reference source is copied into the repository: generated headers and
# objects live under oracle/_ref/ (git-ignored), the instrumented copies of two
# .cpp files live in a mktemp dir that is removed after the build.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/htslib" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt oracle/_ref (if any)" >&2
  exit 0
fi
mkdir -p "$OUT/gen" "$OUT/obj/hts"
GEN="$OUT/gen"
cat > "$GEN/config.h" <<'H'
/* hand-written: no bz2 / lzma / curl in this image */
#ifndef _XOPEN_SOURCE
#define _XOPEN_SOURCE 600
#endif
#define HAVE_DRAND48 1
H
cat > "$GEN/version.h" <<'H'
#define HTS_VERSION_TEXT "1.15.1"
#define HTSCODECS_VERSION_TEXT "1.2.2"
#define VCFGL_VERSION "v1.3.0-da6a334"
H
cat > "$GEN/config_vars.h" <<'H'
#define HTS_CC "gcc"
#define HTS_CPPFLAGS ""
#define HTS_CFLAGS "-O2"
#define HTS_LDFLAGS ""
#define HTS_LIBS "-lz -lm -lpthread"
H
cat > "$GEN/build.h" <<'H'
#define VCFGL_MAKE_CXX ("g++")
#define VCFGL_MAKE_LIBS ("-lz -lm -lpthread")
#define VCFGL_MAKE_FLAGS ("")
#define VCFGL_MAKE_HTSSRC ("bundled")
#define VCFGL_MAKE_CXXFLAGS ("-O3")
#define VCFGL_MAKE_CPPFLAGS ("")
H

H="$REF/htslib"
HTS_SRCS="kfunc kstring bcf_sr_sort bgzf errmod faidx header hfile hts hts_expr hts_os md5 multipart probaln realn regidx region sam synced_bcf_reader vcf_sweep tbx textutils thread_pool vcf vcfutils
cram/cram_codecs cram/cram_decode cram/cram_encode cram/cram_external cram/cram_index cram/cram_io cram/cram_stats cram/mFILE cram/open_trace_file cram/pooled_alloc cram/string_alloc
htscodecs/htscodecs/arith_dynamic htscodecs/htscodecs/fqzcomp_qual htscodecs/htscodecs/htscodecs htscodecs/htscodecs/pack htscodecs/htscodecs/rANS_static4x16pr htscodecs/htscodecs/rANS_static htscodecs/htscodecs/rle htscodecs/htscodecs/tokenise_name3"
JOBS="${JOBS:-$(nproc)}"
objs=""
pids=""
n=0
for s in $HTS_SRCS; do
  o="$OUT/obj/hts/$(echo "$s" | tr '/' '_').o"
  objs="$objs $o"
  if [ ! -f "$o" ]; then
    gcc -g0 -O2 -fno-strict-aliasing -fPIC -w -I"$GEN" -I"$H" -c "$H/$s.c" -o "$o" &
    n=$((n+1))
    if [ "$n" -ge "$JOBS" ]; then wait; n=0; fi
  fi
done
wait
rm -f "$OUT/libhts_ref.a"
ar rcs "$OUT/libhts_ref.a" $objs

CXXFLAGS="-O3 -w -I$GEN -I$H -I$REF"
LIBS="$OUT/libhts_ref.a -lz -lm -lpthread"
# (1) unmodified reference
for f in vcfgl io bcf_utils gl_methods shared; do
  g++ $CXXFLAGS -c "$REF/$f.cpp" -o "$OUT/obj/$f.o" &
done
wait
g++ -o "$OUT/vcfgl_ref" "$OUT/obj/vcfgl.o" "$OUT/obj/io.o" "$OUT/obj/bcf_utils.o" "$OUT/obj/gl_methods.o" "$OUT/obj/shared.o" $LIBS

# (2) instrumented copy (temporary patched sources)
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
python3 "$HERE/patch_ref_for_dump.py" "$REF" "$TMP"
g++ $CXXFLAGS -I"$HERE" -c "$TMP/vcfgl.cpp" -o "$OUT/obj/vcfgl_dump.o" &
g++ $CXXFLAGS -I"$HERE" -c "$TMP/gl_methods.cpp" -o "$OUT/obj/gl_methods_dump.o" &
wait
g++ -o "$OUT/vcfgl_ref_dump" "$OUT/obj/vcfgl_dump.o" "$OUT/obj/io.o" "$OUT/obj/bcf_utils.o" "$OUT/obj/gl_methods_dump.o" "$OUT/obj/shared.o" $LIBS

# (2b) the INTEGRATION.md binding compiled for real: the same instrumented copy with the hot path answered by libvgl.so
#      (replay of the reference's own draws, oracle/ref_vgl_binding.h); needs the product library built first
LIBVGL_DIR="$HERE/../vcfgl_b200"
if [ -f "$LIBVGL_DIR/libvgl.so" ]; then
  TMP2="$(mktemp -d)"
  python3 "$HERE/patch_ref_for_dump.py" "$REF" "$TMP2" --vgl > /dev/null
  g++ $CXXFLAGS -I"$HERE" -I"$HERE/../include" -c "$TMP2/vcfgl.cpp" -o "$OUT/obj/vcfgl_vgl.o"
  g++ -o "$OUT/vcfgl_ref_vgl" "$OUT/obj/vcfgl_vgl.o" "$OUT/obj/io.o" "$OUT/obj/bcf_utils.o" "$OUT/obj/gl_methods_dump.o" "$OUT/obj/shared.o" $LIBS \
      -L"$LIBVGL_DIR" -lvgl -Wl,-rpath,'$ORIGIN/../../vcfgl_b200'
  rm -rf "$TMP2"
else
  echo "build_ref.sh: vcfgl_b200/libvgl.so not built yet; skipping vcfgl_ref_vgl" >&2
fi

# (2c) a reader on the reference's htslib (oracle/hts_read_bcf.c, our code): which records does the library see in a BCF file
gcc -O2 -w -I"$GEN" -I"$H" -I"$H/.." "$HERE/hts_read_bcf.c" -o "$OUT/hts_read_bcf" $LIBS

# (3) errmod alone, for table-level cross-checks of the oracle restatement
gcc -O2 -fPIC -shared -w -I"$GEN" -I"$H" "$H/errmod.c" "$H/hts_os.c" -o "$OUT/libref_errmod.so" -lm

# (4) the reference's LUTs (shared.cpp) as a tiny shared object, for table checks
g++ -O2 -fPIC -shared -w -I"$REF" "$REF/shared.cpp" -o "$OUT/libref_shared.so"
echo "oracle/_ref built: $(ls "$OUT" | tr '\n' ' ')"
