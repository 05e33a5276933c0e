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
# CALC_DECLIN (driver/module_hrldas_noahmp_driver.F90) is an external procedure whose only non-arithmetic part is reading
# day and time out of a date string: it is extracted as it stands, the string handling (the NOWDATE dummy, GETH_IDTS and
# the three internal READs) replaced by integer dummies IDAY, IHOUR, IMINUTE, ISECOND, and wrapped into a module
python - "$R/driver/module_hrldas_noahmp_driver.F90" "$T/ref_calc_declin.F90" <<'PY'
import re, sys
src = open(sys.argv[1], errors="replace").read()
m = re.search(r"^subroutine CALC_DECLIN\(.*?^end subroutine CALC_DECLIN", src, flags=re.S | re.M | re.I)
body = m.group(0)
body = re.sub(r"subroutine CALC_DECLIN\(NOWDATE,", "subroutine CALC_DECLIN(IDAY, IHOUR, IMINUTE, ISECOND,", body, count=1, flags=re.I)
keep = []
for line in body.split("\n"):
    if re.search(r"use MODULE_DATE_UTILITIES|NOWDATE", line, flags=re.I):
        continue
    keep.append(line)
open(sys.argv[2], "w").write("MODULE REF_DRIVER_EXTRACT\nCONTAINS\n" + "\n".join(keep) + "\nEND MODULE REF_DRIVER_EXTRACT\n")
PY
python ref/f90cxx.py "$T/ref_gen.cpp" "$T/ref_calc_declin.F90" $R/util/module_model_constants.F \
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
