"""Shared helpers of the test-suite: run N steps of a synthetic case through the CPU oracle or the CUDA path."""
import copy

import numpy as np

from noahmp_b200 import _capi, synthetic as S


def make_case(cfg, tables, xs=1, xe=None, ys=1, ye=None):
    xp = S.backend()
    st = S.static_fields(xp, cfg, xs, xe, ys, ye)
    frc1 = S.forcing(xp, cfg, 1, st)
    state = S.cold_start(cfg, st, frc1, tables)
    return xp, st, state


def clone_state(state):
    return {k: v.copy() for k, v in state.items()}


def run_oracle(cfg, tables_struct, st, state, nsteps, math_mode=1, nthreads=8, first_step=1):
    from oracle import oracle as O
    xp = S.backend()
    O.set_math_mode(math_mode)
    worst = None
    for step in range(first_step, first_step + nsteps):
        frc = S.forcing(xp, cfg, step, st)
        arr, sc = S.args_from(cfg, st, frc, state, step)
        status, _ = O.noahmplsm(arr, sc, tables_struct, nthreads=nthreads)
        if status.code and worst is None:
            worst = (step, status.code, status.i, status.j, status.count, status.value)
    return worst


def run_gpu(model, cfg, st, state, nsteps, first_step=1):
    xp = S.backend()
    worst = None
    for step in range(first_step, first_step + nsteps):
        frc = S.forcing(xp, cfg, step, st)
        arr, sc = S.args_from(cfg, st, frc, state, step)
        status = model.noahmplsm(arr, sc)
        if status.code and worst is None:
            worst = (step, status.code, status.i, status.j, status.count, status.value)
    return worst


def diff_report(a, b, names=None):
    """Per-field max abs difference and count of differing words between two state dicts."""
    rep = {}
    for n in names or (_capi.INOUT_NAMES + _capi.OUT_NAMES):
        x, y = a[n], b[n]
        if x.dtype.kind == "f":
            same = (x == y) | (np.isnan(x) & np.isnan(y))
            nbad = int((~same).sum())
            if nbad:
                with np.errstate(all="ignore"):
                    rep[n] = (nbad, float(np.nanmax(np.abs(x.astype(np.float64) - y.astype(np.float64))[~same])))
        else:
            nbad = int((x != y).sum())
            if nbad:
                rep[n] = (nbad, float(np.abs(x.astype(np.int64) - y).max()))
    return rep
