"""Golden vectors produced by THE REFERENCE ITSELF — its Fortran text machine-translated to C++ (oracle/ref/f90cxx.py) and
compiled into oracle/_ref/libnoahmp_ref.so — for small synthetic cases whose inputs the test-suite regenerates from
the same seeds: every INOUT / OUT array of `noahmplsm` after the listed steps, in the portable-math mode (the
transcendentals of noahmp_b200/csrc/nmp_math.h, so that the numbers do not depend on the host libm and the CUDA PARITY
build can be held to them bit for bit).  Needs /root/reference (build container only); the vectors travel.
usage: python tests/golden/gen_reference_vectors.py     (rewrites tests/golden/reference_vectors.npz)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from noahmp_b200 import _capi, synthetic as S, tables  # noqa: E402
from helpers import make_case  # noqa: E402

# name -> (base configuration, ni, nj, option overrides, extra Config attributes, steps at which the state is kept)
CASES = {
    "c1_default": ("C1", 10, 10, {}, {}, (1, 12, 24)),
    "c3_dynveg_snow": ("C3", 24, 16, {}, {}, (1, 6, 12)),
    "c4_glacier_water": ("C4", 24, 16, {}, {"snow_frac": 0.4, "t_base": 268.0, "glacier_frac": 0.2}, (1, 6)),
    "opts_a": ("C3", 16, 12, dict(idveg=1, iopt_crs=2, iopt_btr=2, iopt_run=2, iopt_sfc=2, iopt_frz=2, iopt_inf=2,
                                  iopt_rad=1, iopt_alb=1, iopt_snf=2, iopt_tbot=1, iopt_stc=2), {"glacier_frac": 0.1}, (4,)),
    "opts_b": ("C3", 16, 12, dict(idveg=3, iopt_crs=1, iopt_btr=3, iopt_run=3, iopt_sfc=1, iopt_frz=1, iopt_inf=1,
                                  iopt_rad=2, iopt_alb=2, iopt_snf=3, iopt_tbot=2, iopt_stc=1), {"glacier_frac": 0.1}, (4,)),
    "opts_c": ("C3", 16, 12, dict(idveg=5, iopt_crs=2, iopt_btr=1, iopt_run=4, iopt_sfc=2, iopt_frz=2, iopt_inf=1,
                                  iopt_rad=3, iopt_alb=1, iopt_snf=1, iopt_tbot=2, iopt_stc=2), {"glacier_frac": 0.1}, (4,)),
}


def make(name, td):
    base, ni, nj, opts, extra, steps = CASES[name]
    cfg = S.named_config(base)
    cfg.ni, cfg.nj = ni, nj
    cfg.opts.update(opts)
    for k, v in extra.items():
        setattr(cfg, k, v)
    xp, st, state = make_case(cfg, td)
    return cfg, xp, st, state, steps


def run(name, td, step_fn):
    """{'<name>/<step>/<array>': array} with step_fn(arrays, scalars) advancing the state one call"""
    cfg, xp, st, state, steps = make(name, td)
    out = {}
    for step in range(1, max(steps) + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        step_fn(arr, sc)
        if step in steps:
            for n in _capi.INOUT_NAMES + _capi.OUT_NAMES:
                out["%s/%d/%s" % (name, step, n)] = state[n].copy()
    return out


if __name__ == "__main__":
    from oracle.ref import refmodel
    td = tables.default_tables("USGS")
    R = refmodel.RefModel(refmodel.build())
    R.set_tables(_capi.tables_from_dict(td))
    R.set_math_mode(1)
    vec = {}
    for name in CASES:
        vec.update(run(name, td, R.noahmplsm))
    p = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(p, **vec)
    print("written", p, len(vec), "arrays", os.path.getsize(p), "bytes")
