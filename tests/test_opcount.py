"""The op-counting instantiation of the oracle (oracle/nmo_count.h): the physics sources compiled with
`float` replaced by a counting wrapper.  It must (i) give the bits the ordinary oracle gives and (ii) count what is
written: known answers for routines whose operation count can be read off the reference source."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from noahmp_b200 import _capi, synthetic as S

from helpers import clone_state, diff_report, make_case, run_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["ADD", "MUL", "DIV", "CMP", "EXP", "LOG", "LOG10", "POW", "DPOW", "SQRT", "ATAN", "TAN", "COS", "SIN", "ASIN",
       "ACOS", "TANH"]


@pytest.fixture(scope="module")
def opc():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "libnmo_opcount.so"])
    L = C.CDLL(os.path.join(ROOT, "oracle", "libnmo_opcount.so"))
    L.nmo_noahmplsm.argtypes = [C.POINTER(_capi.NoahmpLsmArgs), C.POINTER(_capi.NoahmpTables),
                                C.POINTER(_capi.NoahmpStatus), C.c_int, C.POINTER(C.c_int32)]
    L.nmo_opcount_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    L.nmo_esat.argtypes = [C.c_float, C.POINTER(C.c_float)]
    L.nmo_rosr12.argtypes = [C.c_int] + [C.c_void_p] * 5
    return L


def counts(L, reset=True):
    buf = (C.c_ulonglong * len(OPS))()
    L.nmo_opcount_read(buf, int(reset))
    return dict(zip(OPS, list(buf)))


def test_esat_counts(opc):
    """ESAT (noahmplsm.F90:5272-5321): four degree-6 Horner polynomials, each times 100."""
    out = (C.c_float * 4)()
    counts(opc)
    opc.nmo_esat(C.c_float(-3.5), out)
    c = counts(opc)
    assert (c["ADD"], c["MUL"], c["DIV"], c["CMP"]) == (24, 28, 0, 0)
    assert all(c[k] == 0 for k in OPS[4:])
    assert 400.0 < out[0] < 500.0  # es over water at -3.5 C, Pa


def test_rosr12_counts(opc):
    """ROSR12 (noahmplsm.F90:5979-6036), n unknowns, read off the source: the top row costs 2 divisions; every further
    row evaluates 1/(B + A*P) twice (2 divisions, 2 multiplies, 2 adds) plus -C*(..) and (D - A*DELTA)*(..) (3
    multiplies, 1 add); the back substitution is one multiply-add per row."""
    n = 4
    a = np.array([0, -1, -1, -1], np.float32); b = np.array([4, 4, 4, 4], np.float32)
    c = np.array([-1, -1, -1, 0], np.float32); d = np.array([1, 2, 3, 4], np.float32); x = np.zeros(4, np.float32)
    counts(opc)
    opc.nmo_rosr12(n, *[v.ctypes.data for v in (a, b, c, d, x)])
    k = counts(opc)
    dense = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
    assert np.allclose(dense @ x, d, atol=1e-5)
    assert k["DIV"] == 2 + 2 * (n - 1) and k["MUL"] == 5 * (n - 1) + (n - 1) and k["ADD"] == 3 * (n - 1) + (n - 1)


def test_counting_build_is_bit_identical_and_deterministic(built, tables_usgs, opc):
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = 40, 30
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state0 = make_case(cfg, tables_usgs)
    ref, cnt = clone_state(state0), clone_state(state0)
    assert run_oracle(cfg, ts, st, ref, 3, math_mode=0, nthreads=1) is None
    xp = S.backend()
    opc.nmo_set_math_mode(0)
    tot = []
    for rep in range(2):
        s = clone_state(state0)
        counts(opc)
        for step in (1, 2, 3):
            arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), s, step)
            a = _capi.make_args(arr, sc)
            status = _capi.NoahmpStatus()
            opc.nmo_noahmplsm(C.byref(a), C.byref(ts), C.byref(status), 1, None)
            assert status.code == 0
        tot.append(counts(opc))
        cnt = s
    assert not diff_report(ref, cnt)
    assert tot[0] == tot[1]
    ncs = 3 * int((st["xland"] < 1.5).sum())
    per = {k: v / ncs for k, v in tot[0].items()}
    # a land / glacier column-step is a few thousand flops and a few hundred divisions and transcendentals (SURVEY.md §8d
    # hand estimate: 3-12 k flop, 120-600 transcendental calls)
    assert 2000 < per["ADD"] + per["MUL"] < 12000 and 200 < per["DIV"] < 1500
    assert 50 < per["EXP"] + per["LOG"] + per["POW"] + per["SQRT"] < 800
    assert 0.5 < per["DPOW"] <= 1.0  # one fp64 pow per land column-step (GROUNDWATER's S_NODE), none on glacier columns


def test_committed_counts_feed_the_bench():
    import json
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_opcount.json")))
    for name in ("C3", "C2", "C4"):
        r = d[name]
        assert len(r["by_hour_utc"]) == 24
        assert r["fp32_instr_per_column_step"] <= r["fp32_instr_unfused_per_column_step"]
        day = max(v["fp32_instr"] for v in r["by_hour_utc"].values())
        night = min(v["fp32_instr"] for v in r["by_hour_utc"].values())
        if name != "C4":  # the global domain has day somewhere at every hour
            assert day > 1.2 * night  # night columns skip ALBEDO / TWOSTREAM / STOMATA
