#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference CPU solver (serial, -D_CPU, no BLAS/MPI/HDF5/CGNS)
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored, travels with gpurun).
# Nothing from the reference is copied into the repo; the single accommodation is an on-the-fly sed of
# src/bdy_inters.cpp (uninitialised `id` in add_les_inlet, bdy_inters.cpp:1196-1298 -> crash with no inlet BC),
# piped straight into g++.
set -euo pipefail
REF=${HIFILES_REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/src" ] || { echo "reference tree not present at $REF: keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT/obj" "$OUT/data"
CXXFLAGS="-std=c++14 -D_CPU -O3 -fPIC -I$REF/include -include cstdint -w"
SRCS="global param_reader input bc mesh_reader probe_input flux source cubature_tet cubature_hexa cubature_quad cubature_tri cubature_1d cubature_pris funcs wall_model_funcs inters bdy_inters int_inters eles eles_tris eles_quads eles_tets eles_hexas eles_pris output geometry solver mesh"
pids=()
for f in $SRCS HiFiLES; do
  if [ "$f" = bdy_inters ]; then
    ( sed 's/^\(\s*\)inlet\.nbs=count;/\1inlet.nbs=count; if(count==0){ inlet.type=0; return; }/' "$REF/src/$f.cpp" \
      | g++ $CXXFLAGS -I"$REF/src" -x c++ -c - -o "$OUT/obj/$f.o" ) &
  else
    g++ $CXXFLAGS -c "$REF/src/$f.cpp" -o "$OUT/obj/$f.o" &
  fi
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
g++ $(ls "$OUT"/obj/*.o | grep -v '/ref_dump.o$' | grep -v '/ref_gpu.o$') -o "$OUT/HiFiLES_ref"
# instrumented dumper (our own TU, oracle/ref_dump.cpp) linked against the reference objects minus its main()
if [ -f "$HERE/ref_dump.cpp" ]; then
  objs=$(ls "$OUT"/obj/*.o | grep -v '/HiFiLES.o$' | grep -v '/ref_dump.o$' | grep -v '/ref_gpu.o$')
  g++ $CXXFLAGS -c "$HERE/ref_dump.cpp" -o "$OUT/obj/ref_dump.o"
  g++ "$OUT/obj/ref_dump.o" $objs -o "$OUT/ref_dump"
fi
# run-time tables the reference binary reads from $HIFILES_HOME/data (cubature_1d.cpp:50-85): data, not source
cp -f "$REF"/data/*.bin "$OUT/data/"
echo "built $OUT/HiFiLES_ref"
