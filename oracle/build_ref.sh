#!/usr/bin/env bash
# Build the UNMODIFIED reference (genome/breakdancer @ /root/reference) into oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the
# product path (breakdancer_b200/); only tests/, __graft_entry__.smoke() and bench.py's
# cpu_baseline / --impl reference legs may use it, and only as the checker / baseline.
#
# The reference's own build system (CMake 2.8 + ExternalProject + b2) is NOT run; the
# sources are compiled where they lie under $REF with plain g++, following SURVEY.md §8(c):
#   1. vendor/samtools-0.1.19.tar.gz  -> libbam.a (+ samtools CLI, no curses)
#   2. vendor/boost-1.54-breakdancer.tar.gz -> libboost_bd.a (serialization, regex, system, chrono)
#   3. version/version.h.in -> version.h
#   4. the TUs listed in src/lib/{common,io,breakdancer}/CMakeLists.txt + BreakDancerMax.cpp
# Outputs (binaries only, no reference sources) go to oracle/_ref/:
#   breakdancer-max   the reference executable (release flags: -O2 -DNDEBUG, double score type)
#   samtools          vendored samtools 0.1.19 CLI (view/index for fixtures)
#   score_ref         harness: boost-1.54 Poisson complement-cdf log-probabilities (see score_ref.cpp)
#   breakdancer-max-trace   the reference with ComputeProbScore's Poisson calls traced to stderr (ref_trace.cpp)
#   breakdancer-max-nodecode   the reference's algorithm on records decoded into memory beforehand (ref_nodecode.cpp)
# Intermediates live in oracle/_ref/build (gpurun-ignored).
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
B=$OUT/build
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt oracle/_ref as is" >&2
  exit 0
fi
if [ -x "$OUT/breakdancer-max" ] && [ -x "$OUT/samtools" ] && [ -x "$OUT/score_ref" ] && [ -x "$OUT/breakdancer-max-trace" ] && [ -x "$OUT/breakdancer-max-nodecode" ] && [ "${FORCE:-0}" != 1 ]; then
  echo "build_ref.sh: oracle/_ref already built (FORCE=1 to rebuild)"
  exit 0
fi
mkdir -p "$B"
cd "$B"
# 1. samtools
if [ ! -f samtools-0.1.19/libbam.a ]; then
  tar xzf "$REF/vendor/samtools-0.1.19.tar.gz"
  make -C samtools-0.1.19 -j"$JOBS" libbam.a CFLAGS="-g -Wall -O2 -fPIC -w" >/dev/null
fi
if [ ! -x "$OUT/samtools" ]; then
  make -C samtools-0.1.19 -j"$JOBS" samtools CFLAGS="-g -Wall -O2 -fPIC -w" \
      DFLAGS="-D_FILE_OFFSET_BITS=64 -D_LARGEFILE64_SOURCE -D_USE_KNETFILE -D_CURSES_LIB=0" LIBCURSES= >/dev/null
  cp samtools-0.1.19/samtools "$OUT/samtools"
fi
# 2. boost (no bootstrap / b2: compile the needed library TUs directly)
if [ ! -f libboost_bd.a ]; then
  [ -d boost-bd ] || tar xzf "$REF/vendor/boost-1.54-breakdancer.tar.gz"
  mkdir -p bobj
  ls boost-bd/libs/{serialization,regex,system,chrono}/src/*.cpp \
    | grep -v -E 'xml_w|wiarchive|woarchive|text_w|utf8_codecvt|codecvt_null|_w[io]' > boost_tus.txt
  cat boost_tus.txt | xargs -P "$JOBS" -I{} sh -c \
    'o=bobj/$(echo {} | tr / _).o; g++ -std=c++11 -O2 -w -fPIC -Iboost-bd -c {} -o $o'
  ar rcs libboost_bd.a bobj/*.o
fi
# 3. version header
mkdir -p ver
sed -e 's/@FULL_VERSION@/1.4.5-oracle/' -e 's/@COMMIT_HASH@/4e44b43/' "$REF/version/version.h.in" > ver/version.h
# 4. reference TUs
CXXF="-std=c++11 -O2 -DNDEBUG -DSCORE_FLOAT_TYPE=double -w -fPIC -I$REF/src/lib -Iboost-bd -Isamtools-0.1.19 -Iver"
mkdir -p robj
TUS="common/Options.cpp common/ReadFlags.cpp
io/Alignment.cpp io/BamConfig.cpp io/BamConfigEntry.cpp io/BamIo.cpp io/BamMerger.cpp io/BamSummary.cpp
io/BamWriter.cpp io/ConfigLoader.cpp io/FastqWriter.cpp io/IlluminaPEReadClassifier.cpp io/LibraryFlagDistribution.cpp
breakdancer/BedWriter.cpp breakdancer/BreakDancer.cpp breakdancer/ReadRegionData.cpp breakdancer/SvBuilder.cpp"
for t in $TUS; do echo "$t"; done | xargs -P "$JOBS" -I{} sh -c \
  "o=robj/\$(echo {} | tr / _).o; g++ $CXXF -c $REF/src/lib/{} -o \$o"
g++ $CXXF -c "$REF/src/exe/breakdancer-max/BreakDancerMax.cpp" -o robj/main.o
g++ -o "$OUT/breakdancer-max" robj/main.o $(ls robj/*.o | grep -v -E 'main.o|trace_|nodecode_') libboost_bd.a samtools-0.1.19/libbam.a -lz -lm -lpthread -lrt
# 5. harnesses (our own sources, compiled against the reference's vendored headers)
g++ -std=c++11 -O2 -w -Iboost-bd "$HERE/score_ref.cpp" -o "$OUT/score_ref"
g++ $CXXF -c "$HERE/ref_trace.cpp" -o robj/trace_BreakDancer.o
g++ -o "$OUT/breakdancer-max-trace" robj/main.o robj/trace_BreakDancer.o $(ls robj/*.o | grep -v -E 'main.o|breakdancer_BreakDancer.cpp.o|trace_') libboost_bd.a samtools-0.1.19/libbam.a -lz -lm -lpthread -lrt
g++ $CXXF -Dmain=reference_main -c "$REF/src/exe/breakdancer-max/BreakDancerMax.cpp" -o robj/nodecode_main.o
g++ $CXXF -c "$HERE/ref_nodecode.cpp" -o robj/nodecode_io.o
g++ -o "$OUT/breakdancer-max-nodecode" robj/nodecode_main.o robj/nodecode_io.o $(ls robj/*.o | grep -v -E 'main.o|io_BamIo.cpp.o|trace_|nodecode_') libboost_bd.a samtools-0.1.19/libbam.a -lz -lm -lpthread -lrt
echo "build_ref.sh: built $(ls "$OUT" | grep -v build | tr '\n' ' ')"
