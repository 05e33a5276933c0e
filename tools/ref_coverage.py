"""Which statements of the reference did the pin exercise?

Runs the cases of tests/test_reference_pin.py (oracle == translated reference, bit for bit) through the
statement-coverage instantiation of the translated reference (oracle/_ref/libnoahmp_ref_cov.so: every executable
Fortran statement sets a flag) and reports, per routine of the reference, how many of its executable statements were
executed while the comparison held — and lists the ones that were not.

usage: python tools/ref_coverage.py [report.txt]        (needs /root/reference; oracle/ref/build_ref.sh trace)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from noahmp_b200 import _capi, tables  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.ref import refmodel  # noqa: E402
import test_reference_pin as P  # noqa: E402

# routines that are not on the offline path (options the HRLDAS driver rejects, or never called)
OFF_PATH = {"SFCDIF3", "SFCDIF4", "BVOCFLUX", "CI2CI"}


def main():
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    td = tables.default_tables("USGS")
    ts = _capi.tables_from_dict(td)
    O.lib()
    R = refmodel.RefModel(os.path.join(ROOT, "oracle", "_ref", "libnoahmp_ref_cov.so"))
    R.set_tables(ts)
    ran = []
    for mode in (0, 1):
        P.test_c1_24_steps(O, R, td, ts, mode)
        P.test_c3_dynamic_vegetation_and_snow(O, R, td, ts, mode)
    P.test_c2_nldas_tile(O, R, td, ts)
    P.test_c4_glacier_seaice_water(O, R, td, ts)
    for k in range(len(P.OPTS)):
        P.test_every_accepted_option_value(O, R, td, ts, k)
    for case in P.CLIMATES:
        P.test_other_climates(O, R, td, ts, case)
    P.test_long_melt_season(O, R, td, ts)
    P.test_glacier_in_summer(O, R, td, ts)
    P.test_sea_ice_points_soil_type_14_and_dry_soil(O, R, td, ts)
    for a in [("C4", 96, 64, 1), ("C3", 64, 48, 1), ("C2", 60, 44, 5), ("C3", 130, 70, 5)]:
        P.test_noahmp_init(O, R, ts, *a)
    for a in [("C4", 40, 30), ("C3", 96, 64), ("C2", 80, 60)]:
        P.test_wtable_coupled_with_the_column_physics(O, R, td, ts, *a)
    P.test_wtable_rising_and_falling_through_the_layers(O, R, td, ts)
    P.test_leaf_routines_over_wide_ranges(O, R, ts)
    P.test_stomata_and_twostream_over_wide_ranges(O, R, ts)
    P.test_snow_routines_on_random_packs(O, R)
    P.test_calc_declin_of_the_forcing_pipeline(O, R, td)
    ran.append("all cases of tests/test_reference_pin.py held (oracle == translated reference)")

    R.lib.ref_cover_map.restype = C.POINTER(C.c_ubyte)
    cover = np.ctypeslib.as_array(R.lib.ref_cover_map(), shape=(1000000,))
    n = C.c_int()
    R.lib.ref_markers.restype = C.POINTER(C.c_int)
    mk = R.lib.ref_markers(C.byref(n))
    mk = np.ctypeslib.as_array(mk, shape=(n.value, 2))
    R.lib.ref_marker_names.restype = C.c_char_p
    names = R.lib.ref_marker_names().decode().split(";")
    R.lib.ref_files.restype = C.c_char_p
    files = R.lib.ref_files().decode().split(";")
    src = [open(p, errors="replace").read().split("\n") if os.path.exists(p) else [] for p in files]
    per = {}
    for line, ni in mk:
        per.setdefault(names[ni], []).append((int(line), bool(cover[line])))
    tot = hit = 0
    rows = []
    for name, lst in per.items():
        lst = sorted(set(lst))
        h = sum(1 for _, c in lst if c)
        rows.append((name, h, len(lst), [l for l, c in lst if not c]))
        if name not in OFF_PATH:
            tot += len(lst)
            hit += h
    print(ran[0], file=out)
    print("executable statements of the translated reference reached by those cases: %d of %d (%.1f %%) "
          "[routines off the offline path left out: %s]" % (hit, tot, 100.0 * hit / tot, ", ".join(sorted(OFF_PATH))), file=out)
    print("\n%-28s %6s %6s" % ("routine", "hit", "of"), file=out)
    for name, h, t, miss in sorted(rows, key=lambda r: r[0]):
        print("%-28s %6d %6d%s" % (name, h, t, "   (off path)" if name in OFF_PATH else ""), file=out)
    print("\nstatements not reached:", file=out)
    for name, h, t, miss in sorted(rows, key=lambda r: r[0]):
        if name in OFF_PATH or not miss:
            continue
        print("  %s" % name, file=out)
        for l in miss:
            fid, ln = divmod(l, 100000)
            text = src[fid][ln - 1].strip() if (ln <= len(src[fid]) and out is sys.stdout) else ""
            print("    %s:%d  %s" % (os.path.basename(files[fid]), ln, text[:110]), file=out)


if __name__ == "__main__":
    main()
