"""ctypes loader for the CPU ORACLE (oracle/libnmo_oracle.so) — test infrastructure.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The product path (noahmp_b200) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from noahmp_b200 import _capi  # noqa: E402

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libnmo_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.nmo_noahmplsm.argtypes = [C.POINTER(_capi.NoahmpLsmArgs), C.POINTER(_capi.NoahmpTables),
                                       C.POINTER(_capi.NoahmpStatus), C.c_int, C.POINTER(C.c_int32)]
        _LIB.nmo_noahmplsm.restype = C.c_int
        _LIB.nmo_math1.argtypes = [C.c_int, C.c_float, C.c_float]
        _LIB.nmo_math1.restype = C.c_float
        _LIB.nmo_math_array.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
        _LIB.nmo_esat.argtypes = [C.c_float, C.POINTER(C.c_float)]
        _LIB.nmo_rosr12.argtypes = [C.c_int] + [C.c_void_p] * 5
        _LIB.nmo_combo.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.nmo_wtable.argtypes = [C.POINTER(_capi.NoahmpWtableArgs), C.POINTER(_capi.NoahmpTables)]
        _LIB.nmo_wtable.restype = C.c_int
    return _LIB


def set_math_mode(mode):
    """0 = host libm (reference-like), 1 = portable nmp_math.h (bit-comparable with the GPU parity build)."""
    lib().nmo_set_math_mode(int(mode))


def noahmplsm(arrays, scalars, tables_struct, nthreads=1, want_iters=False):
    """One call of the oracle's noahmplsm on host arrays (updated in place). Returns (status, iters)."""
    a = _capi.make_args(arrays, scalars)
    st = _capi.NoahmpStatus()
    it = None
    itp = None
    if want_iters:
        it = np.zeros(arrays["tsk"].shape, np.int32)
        itp = it.ctypes.data_as(C.POINTER(C.c_int32))
    lib().nmo_noahmplsm(C.byref(a), C.byref(tables_struct), C.byref(st), int(nthreads), itp)
    return st, it


def math_array(fn, x, y=None):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    yp = None
    if y is not None:
        y = np.ascontiguousarray(y, np.float32)
        yp = y.ctypes.data
    lib().nmo_math_array(fn, x.ctypes.data, yp, out.ctypes.data, x.size)
    return out


def wtable(arrays, scalars, tables_struct):
    """One call of the oracle's WTABLE_mmf_noahmp on host arrays (updated in place)."""
    a = _capi.make_wtable_args(arrays, scalars)
    rc = lib().nmo_wtable(C.byref(a), C.byref(tables_struct))
    if rc:
        raise RuntimeError(f"nmo_wtable failed: {rc}")


def init(arrays, scalars, tables_struct):
    """The oracle's NOAHMP_INIT on host arrays (updated in place). Returns (rc, STEPWTD or None)."""
    arrays = dict(arrays)
    step = np.zeros(1, np.int32)
    if scalars.get("iopt_run") == 5:
        arrays["stepwtd"] = step
    a = _capi.make_init_args(arrays, scalars)
    rc = lib().nmo_init(C.byref(a), C.byref(tables_struct))
    return rc, (int(step[0]) if scalars.get("iopt_run") == 5 else None)


def forcing(A, B, lat2d, lon2d, fraction, iday, ihour, iminute, isecond, dt, zlvl=30.0):
    """Oracle of the driver-side forcing preparation. A, B: dicts of the 9 forcing-file fields. Returns
    (list of the 12 forcing planes in NOAHMP_NFORCING order, JULIAN)."""
    fa, fb = _capi.make_forcing_fields(A), _capi.make_forcing_fields(B)
    lat = np.ascontiguousarray(lat2d, np.float32); lon = np.ascontiguousarray(lon2d, np.float32)
    out = [np.zeros_like(lat) for _ in range(12)]
    ptrs = (C.c_void_p * 12)(*[o.ctypes.data for o in out])
    fn = lib().nmo_forcing
    fn.restype = C.c_float
    fn.argtypes = [C.POINTER(_capi.NoahmpForcingFields)] * 2 + [C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_int, C.c_int,
                                                                C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    j = fn(C.byref(fa), C.byref(fb), lat.ctypes.data, lon.ctypes.data, lat.size, fraction, iday, ihour, iminute, isecond,
           dt, zlvl, ptrs)
    return out, j
