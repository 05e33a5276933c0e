"""Algorithmic work per column-step of the path: runs the OP-COUNTING instantiation of the oracle
(oracle/nmo_count.h: `float` replaced by a wrapper that counts every add / multiply / divide / compare and every
transcendental call by class; results bit-identical to the ordinary oracle) on a row sample of each workload over one
diurnal cycle and writes profiles/r02_opcount.json, the numerator of bench.py's compute roofline.

usage: python tools/opcount.py [out.json]
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noahmp_b200 import _capi, synthetic as S, tables  # noqa: E402

OPS = ["ADD", "MUL", "DIV", "CMP", "EXP", "LOG", "LOG10", "POW", "DPOW", "SQRT", "ATAN", "TAN", "COS", "SIN", "ASIN",
       "ACOS", "TANH"]
# How the production build issues each class (nmp_common.cuh, --use_fast_math): (FP32-pipe instructions, MUFU instructions)
EXPANSION = {"DIV": (1, 1),     # MUFU.RCP + FMUL
             "EXP": (1, 1),     # FMUL by log2(e) + MUFU.EX2
             "LOG": (1, 1), "LOG10": (1, 1),  # MUFU.LG2 + FMUL
             "POW": (1, 2),     # MUFU.LG2, FMUL, MUFU.EX2
             "SQRT": (0, 1),    # MUFU.SQRT
             "TANH": (0, 1),    # MUFU.TANH
             "COS": (1, 1), "SIN": (1, 1),  # range FMUL + MUFU.COS / SIN
             "ATAN": (16, 1), "TAN": (20, 1), "ACOS": (16, 1), "ASIN": (16, 1),  # libdevice polynomials
             "DPOW": (0, 0)}    # one fp64 pow per column-step (GROUNDWATER's S_NODE): FP64 pipe, listed apart


def lib():
    so = os.path.join(ROOT, "oracle", "libnmo_opcount.so")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "libnmo_opcount.so"])
    L = C.CDLL(so)
    L.nmo_noahmplsm.argtypes = [C.POINTER(_capi.NoahmpLsmArgs), C.POINTER(_capi.NoahmpTables),
                                C.POINTER(_capi.NoahmpStatus), C.c_int, C.POINTER(C.c_int32)]
    L.nmo_opcount_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    return L


def read(L, reset=True):
    buf = (C.c_ulonglong * len(OPS))()
    L.nmo_opcount_read(buf, int(reset))
    return np.array(list(buf), dtype=np.float64)


def derive(c):
    """c: dict class -> count per column-step.  FMA-fused lower bound and unfused upper bound of the FP32-pipe
    instructions, and the MUFU instructions, the production build needs for this work."""
    extra_f = sum(EXPANSION[k][0] * c[k] for k in EXPANSION)
    mufu = sum(EXPANSION[k][1] * c[k] for k in EXPANSION)
    fused = max(c["ADD"], c["MUL"]) + c["CMP"] + extra_f
    unfused = c["ADD"] + c["MUL"] + c["CMP"] + extra_f
    return fused, unfused, mufu


def run(L, name, jstride, warm=3, hours=24, **over):
    cfg = S.named_config(name)
    for k, v in over.items():
        setattr(cfg, k, v)
    td = tables.default_tables("USGS")
    ts = _capi.tables_from_dict(td)
    xp = S.backend()
    st = S.static_fields(xp, cfg, jstride=jstride)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), td)
    ncol = int((st["xland"] < 1.5).sum())
    L.nmo_set_math_mode(0)
    per_hour, sun = [], []
    for k in range(warm + hours):
        frc = S.forcing(xp, cfg, 1 + (k % 24), st)
        arr, sc = S.args_from(cfg, st, frc, state, 1 + k)
        a = _capi.make_args(arr, sc)
        status = _capi.NoahmpStatus()
        read(L)
        L.nmo_noahmplsm(C.byref(a), C.byref(ts), C.byref(status), 1, None)
        assert status.code == 0, status.code
        cnt = read(L) / ncol
        if k >= warm:
            per_hour.append(cnt)
            sun.append(float((frc["coszin"][st["xland"] < 1.5] > 0).mean()))
    per_hour = np.array(per_hour)
    hours_utc = [(cfg.start[3] + k) % 24 for k in range(warm, warm + hours)]
    mean = dict(zip(OPS, per_hour.mean(axis=0).tolist()))
    fused, unfused, mufu = derive(mean)
    by_hour = {}
    for h, row, s in zip(hours_utc, per_hour, sun):
        f, u, m = derive(dict(zip(OPS, row.tolist())))
        by_hour[str(h)] = {"fp32_instr": round(f, 1), "fp32_instr_unfused": round(u, 1), "mufu": round(m, 1), "sunlit": round(s, 3)}
    nj = st["xland"].shape[0]
    res = {"sample": f"every {jstride}th row of {name} {cfg.ni}x{cfg.nj} = {cfg.ni}x{nj} cells, {ncol} columns, 24 hourly steps "
                     f"after {warm} warm-up steps",
           "ops_per_column_step": {k: round(v, 2) for k, v in mean.items()},
           "fp32_instr_per_column_step": round(fused, 1), "fp32_instr_unfused_per_column_step": round(unfused, 1),
           "mufu_per_column_step": round(mufu, 1), "fp64_pow_per_column_step": round(mean["DPOW"], 3),
           "sunlit_fraction_24h": round(float(np.mean(sun)), 3), "by_hour_utc": by_hour,
           "expansion": {k: list(v) for k, v in EXPANSION.items()},
           "definition": "fp32_instr = max(ADD, MUL) [every add fused with a multiply where one exists] + CMP + the FP32 part "
                         "of each transcendental / division as the production build issues it; mufu = the MUFU part"}
    print(name, json.dumps({k: res[k] for k in ("fp32_instr_per_column_step", "fp32_instr_unfused_per_column_step",
                                                "mufu_per_column_step", "sunlit_fraction_24h")}), flush=True)
    return res


def main():
    L = lib()
    out = {"C3": run(L, "C3", 180), "C2": run(L, "C2", 2), "C4": run(L, "C4", 60)}
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_opcount.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
