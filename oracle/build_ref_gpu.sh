#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- links oracle/ref_gpu.cpp (our own TU) against the UNMODIFIED reference objects built by oracle/build_ref.sh
# (minus the reference's main) and against the product library through its C ABI only (include/hifiles_b200.h).  Output:
# oracle/_ref/ref_gpu (git-ignored, travels with gpurun).  Needs /root/reference for the reference's headers; on a box without it the
# prebuilt binary is kept.
set -euo pipefail
REF=${HIFILES_REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
LIB="$HERE/../hifiles-solver_b200/lib"
[ -d "$REF/include" ] || { echo "reference tree not present at $REF: keeping prebuilt $OUT/ref_gpu"; exit 0; }
[ -f "$LIB/libhifiles_b200.so" ] || { echo "libhifiles_b200.so not built yet"; exit 1; }
CXXFLAGS="-std=c++14 -D_CPU -O2 -fPIC -I$REF/include -I$HERE/../include -include cstdint -w"
g++ $CXXFLAGS -c "$HERE/ref_gpu.cpp" -o "$OUT/obj/ref_gpu.o"
objs=$(ls "$OUT"/obj/*.o | grep -v '/HiFiLES.o$' | grep -v '/ref_dump.o$' | grep -v '/ref_gpu.o$')
g++ "$OUT/obj/ref_gpu.o" $objs -L"$LIB" -lhifiles_b200 -Wl,-rpath,'$ORIGIN/../../hifiles-solver_b200/lib' -o "$OUT/ref_gpu"
echo "built $OUT/ref_gpu"
