"""Oracle against the translated reference (oracle/_ref/libnoahmp_ref.so) on a window of a synthetic configuration:
every step both start from the oracle's state, so a difference is attributed to the step it arises in.
usage: python tools/ref_compare.py CONFIG NSTEPS MATH_MODE [xs xe ys ye]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from noahmp_b200 import _capi, synthetic as S, tables  # noqa: E402
from helpers import make_case, clone_state, diff_report  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.ref import refmodel  # noqa: E402


def main():
    cfgname, nsteps, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    win = [int(x) for x in sys.argv[4:8]] if len(sys.argv) >= 8 else [1, None, 1, None]
    T = tables.default_tables("USGS")
    TS = _capi.tables_from_dict(T)
    R = refmodel.RefModel()
    R.set_tables(TS)
    cfg = S.named_config(cfgname)
    xp, st, state = make_case(cfg, T, *win)
    sa = clone_state(state)
    O.set_math_mode(mode)
    R.set_math_mode(mode)
    total = 0
    for step in range(1, nsteps + 1):
        frc = S.forcing(xp, cfg, step, st)
        sb = clone_state(sa)
        arr, sc = S.args_from(cfg, st, frc, sa, step)
        status, _ = O.noahmplsm(arr, sc, TS, nthreads=8)
        arr2, sc2 = S.args_from(cfg, st, frc, sb, step)
        try:
            R.noahmplsm(arr2, sc2)
        except RuntimeError as e:
            print("step", step, "reference stopped:", e, "| oracle status", status.code, status.j, status.i)
            break
        rep = diff_report(sa, sb)
        bad = np.zeros(sa["tsk"].shape, bool)
        for n in rep:
            x, y = sa[n], sb[n]
            d = ~((x == y) | (np.isnan(x) & np.isnan(y))) if x.dtype.kind == "f" else (x != y)
            bad |= d.any(axis=1) if d.ndim == 3 else d
        idx = np.argwhere(bad)
        total += len(idx)
        if len(idx) or status.code:
            print("step", step, "oracle status", status.code, "cells differing", len(idx), "of", bad.size, "fields", sorted(rep)[:8])
            for (a, b) in idx[:4]:
                n0 = sorted(rep, key=lambda n: -rep[n][0])[0]
                print("   cell j=%d i=%d ivgtyp %d isltyp %d isnow %d xice %.2f tsk %.3f" %
                      (a, b, st["ivgtyp"][a, b], st["isltyp"][a, b], sb["isnowxy"][a, b], st["xice"][a, b], sa["tsk"][a, b]))
    print("%s window %s: %d steps, %d cell-steps differ" % (cfgname, win, nsteps, total))


if __name__ == "__main__":
    main()
