#!/bin/bash
# oracle/ref/build_ref.sh [trace] — translate the reference's Fortran WHERE IT LIES (f90cxx.py) and compile the result.
# Only shared libraries are left behind, in oracle/_ref/ (git-ignored; they travel to the GPU box like every built
# .so); the generated C++ lives in a temporary directory (KEEP_GENERATED=dir keeps a copy for debugging).
set -e
cd "$(dirname "$0")/.."
R=${NOAHMP_REFERENCE:-/root/reference}
[ -d "$R/phys" ] || { echo "no reference tree at $R"; exit 3; }
mkdir -p _ref
T=$(mktemp -d)
trap 'rm -rf "$T"' EXIT
python ref/f90cxx.py "$T/ref_gen.cpp" $R/util/module_model_constants.F \
  $R/phys/module_sf_noahmplsm.F90:skip=READ_MP_VEG_PARAMETERS,SFCDIF3,SFCDIF4 \
  $R/phys/module_sf_noahmp_glacier.F90 \
  $R/phys/module_sf_noahmp_groundwater.F90 \
  $R/phys/module_sf_noahmpdrv.F90:only=NOAHMPLSM,NOAHMP_INIT,SNOW_INIT,GROUNDWATER_INIT,EQSMOISTURE:noop=READ_MP_VEG_PARAMETERS,SOIL_VEG_GEN_PARM
[ -n "$KEEP_GENERATED" ] && mkdir -p "$KEEP_GENERATED" && cp "$T/ref_gen.cpp" "$KEEP_GENERATED/"
CXXFLAGS="-std=gnu++17 -ffp-contract=off -fno-fast-math -fPIC -Iref"
g++ $CXXFLAGS -O2 -shared "$T/ref_gen.cpp" ref/ref_shim.cpp -o _ref/libnoahmp_ref.so
if [ "$1" = "trace" ]; then
  # value-tracing / op-counting instantiation (float -> nmo_count::Real): tools/ref_trace_diff.py
  g++ $CXXFLAGS -O1 -Wno-class-memaccess -include nmo_count.h -c "$T/ref_gen.cpp" -o "$T/ref_gen_count.o"
  g++ $CXXFLAGS -O1 -include nmo_count.h -c ref/ref_shim.cpp -o "$T/ref_shim_count.o"
  g++ $CXXFLAGS -O1 -c nmo_count.cpp -o "$T/ref_count.o"
  g++ -shared "$T/ref_gen_count.o" "$T/ref_shim_count.o" "$T/ref_count.o" -o _ref/libnoahmp_ref_count.so
  # statement-coverage instantiation: tools/ref_coverage.py
  g++ $CXXFLAGS -O1 -DREF_COVERAGE -shared "$T/ref_gen.cpp" ref/ref_shim.cpp -o _ref/libnoahmp_ref_cov.so
fi
