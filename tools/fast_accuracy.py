"""Measure the FAST (production) build against the oracle with host libm on the configurations the bench and the
tolerance tests use: per variable max |d|, 99.9th percentile, fraction of columns beyond the stated tolerance, and
the ISNOWXY mismatch fraction, after 1 step and at the end.  The numbers (profiles/r02_fast_accuracy.json) are what
tests/test_fast_parity_gpu.py::FAST_TOL / HARD_CAP were set from (tolerances are measured).

usage: python tools/fast_accuracy.py [out.json]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import noahmp_b200  # noqa: E402
from noahmp_b200 import _capi, synthetic as S, tables  # noqa: E402
from helpers import clone_state, make_case, run_gpu, run_oracle  # noqa: E402

VARS = ["tsk", "tslb", "smois", "sh2o", "snow", "snowh", "hfx", "lh", "grdflx", "sfcrunoff", "udrunoff", "xlaixy",
        "qfx", "tgxy", "tvxy", "canwat", "zwtxy", "waxy", "t2mvxy", "t2mbxy", "albedo", "emiss", "lfmassxy", "gppxy"]


def stats(cfg, st, a, b, tol):
    nonwater = st["xland"] < 1.5
    glac = nonwater & (st["ivgtyp"] == S.ISICE)
    out = {}
    for cls, mask2 in (("land", nonwater & ~glac), ("glacier", glac)):
        if not mask2.any():
            continue
        r = {}
        for n in VARS:
            x, y = a[n].astype(np.float64), b[n].astype(np.float64)
            m = mask2 if x.ndim == 2 else np.broadcast_to(mask2[:, None, :], x.shape)
            d = np.abs(x - y)[m]
            d = d[np.isfinite(d)]
            if d.size == 0:
                continue
            t = tol.get(n, (None, None))[0]
            r[n] = {"max": float(d.max()), "p999": float(np.quantile(d, 0.999)), "mean": float(d.mean()),
                    "frac_gt_tol": float((d > t).mean()) if t else None}
        r["isnow_mismatch"] = float((a["isnowxy"] != b["isnowxy"])[mask2].mean())
        r["columns"] = int(mask2.sum())
        out[cls] = r
    return out


def case(name, cfg, nsteps, td, ts, tol, checkpoints):
    _, st, state0 = make_case(cfg, td)
    s_cpu, s_gpu = clone_state(state0), clone_state(state0)
    m = noahmp_b200.NoahMP(td, cfg.ni, cfg.nj, device=0, math=noahmp_b200.MATH_FAST)
    res, done = {}, 0
    for upto in checkpoints:
        e1 = run_oracle(cfg, ts, st, s_cpu, upto - done, math_mode=0, first_step=done + 1)
        e2 = run_gpu(m, cfg, st, s_gpu, upto - done, first_step=done + 1)
        done = upto
        res[f"after_{upto}"] = stats(cfg, st, s_cpu, s_gpu, tol)
        res[f"after_{upto}"]["status"] = {"oracle": e1, "gpu": e2}
    m.close()
    print(name, json.dumps({k: {c: {"tsk_max": v[c]["tsk"]["max"], "snowh_max": v[c]["snowh"]["max"],
                                    "isnow": v[c]["isnow_mismatch"]} for c in v if c != "status"} for k, v in res.items()}),
          flush=True)
    return res


def main():
    from test_fast_parity_gpu import FAST_TOL
    td = tables.default_tables("USGS")
    ts = _capi.tables_from_dict(td)
    out = {}
    c3 = S.named_config("C3"); c3.ni, c3.nj = 232, 160
    out["C3_232x160_48steps"] = case("C3", c3, 48, td, ts, FAST_TOL, (1, 24, 48))
    c4 = S.named_config("C4"); c4.ni, c4.nj = 240, 180; c4.glacier_frac = 0.3; c4.snow_frac = 0.3
    out["C4_240x180_glacier30_48steps"] = case("C4", c4, 48, td, ts, FAST_TOL, (1, 24, 48))
    c3b = S.named_config("C3"); c3b.ni, c3b.nj = 116, 112
    out["C3_116x112_240steps"] = case("C3-240", c3b, 240, td, ts, FAST_TOL, (1, 120, 240))
    c2 = S.named_config("C2"); c2.ni, c2.nj = 232, 112
    out["C2_232x112_240steps"] = case("C2-240", c2, 240, td, ts, FAST_TOL, (24, 240))
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_fast_accuracy.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
