#!/usr/bin/env bash
# Builds the UNMODIFIED reference (polymutt 0.13) from the sources where they lie under
# /root/reference into oracle/_ref/ (git-ignored).  Test/bench infrastructure only: the product
# never links or executes anything built here.  Recipe follows SURVEY.md §8c:
#   * only the translation units the linker actually pulls in are compiled
#     (src/*.cpp, 27 libcore members, base/IO.cpp, libVcf/VCFIndividual.cpp + VCFInputFile.cpp);
#   * four shim headers in oracle/shims stand in for tabix/bgzf/bzlib/pcre (never reached by
#     polymutt's LINE_MODE VCF reader);
#   * -std=gnu++14 -fpermissive (bool++ in core/Parameters.cpp:518), pow10 -> exp10 (glibc
#     dropped the pow10 alias), the reference's own -D flags, -O3 -fopenmp.
# Usage: oracle/build_ref.sh [REF_DIR]      (default /root/reference)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
if [ ! -d "$REF/src" ]; then echo "reference tree not found at $REF (nothing to do)"; exit 0; fi
mkdir -p "$OBJ"
CXX="${PM_REF_CXX:-g++}"   # deliberately not $CXX: the image exports a second toolchain there
COMMON="-std=gnu++14 -fpermissive -w -O3 -fopenmp -D__ZLIB_AVAILABLE__ -D_FILE_OFFSET_BITS=64 -D__STDC_LIMIT_MACROS -include unistd.h -include string.h -include stdio.h -include stdlib.h"
INC="-I$HERE/shims -I$REF/core -I$REF/base -I$REF/libVcf"
CORE="BaseQualityHelper Error FortranFormat GenotypeLists Hash InputFile IntArray MapFunction MathGold MathMatrix MathVector MemoryInfo Parameters Pedigree PedigreeAlleleFreq PedigreeDescription PedigreeFamily PedigreeGlobals PedigreeLoader PedigreePerson PedigreeTwin QuickIndex Sort StringArray StringBasics StringHash StringMap glfHandler"
SRC="FamilyLikelihoodES FamilyLikelihoodSeq FamilyLikelihoodSeq_VCF MutationModel NucFamGenotypeLikelihood PedVCF PedigreeGLF main"
pids=()
cc() { # cc <dir> <name> <extra flags>
  local src="$REF/$1/$2.cpp" obj="$OBJ/$1_$2.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then
    $CXX -c $COMMON $3 $INC "$src" -o "$obj" &
    pids+=($!)
  fi
}
for f in $CORE; do cc core "$f" ""; done
cc base IO ""
cc libVcf VCFIndividual ""
cc libVcf VCFInputFile ""
for f in $SRC; do cc src "$f" "-Dpow10(x)=exp10(x)"; done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
# everything except main.o also goes into a static archive, so that oracle/ref_dump.cpp (our own
# instrumented driver) can link the reference's classes without its main().
rm -f "$OUT/libpolymutt_ref.a"
ar -cr "$OUT/libpolymutt_ref.a" $(ls "$OBJ"/*.o | grep -v src_main.o)
$CXX -O3 -fopenmp -o "$OUT/polymutt" "$OBJ/src_main.o" "$OUT/libpolymutt_ref.a" -lm -lz -lgomp
echo "built $OUT/polymutt"
