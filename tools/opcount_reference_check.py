"""Op counts of the oracle (oracle/libnmo_opcount.so) beside those of the translated reference compiled with the same
counting `float` (oracle/_ref/libnoahmp_ref_count.so): is the numerator of bench.py's compute roofline the reference's own work?
usage: python tools/opcount_reference_check.py > profiles/r02_opcount_reference_check.txt   (needs oracle/ref/build_ref.sh trace)"""
import sys, ctypes as C, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from noahmp_b200 import _capi, synthetic as S, tables
from helpers import make_case, clone_state
from oracle.ref import refmodel
T = tables.default_tables("USGS"); TS = _capi.tables_from_dict(T)
NAMES = "ADD MUL DIV CMP EXP LOG LOG10 POW DPOW SQRT ATAN TAN COS SIN ASIN ACOS TANH".split()
oc = C.CDLL(os.path.join(ROOT, 'oracle', 'libnmo_opcount.so'))
oc.nmo_noahmplsm.argtypes = [C.POINTER(_capi.NoahmpLsmArgs), C.POINTER(_capi.NoahmpTables), C.POINTER(_capi.NoahmpStatus), C.c_int, C.c_void_p]
R = refmodel.RefModel(os.path.join(ROOT, 'oracle', '_ref', 'libnoahmp_ref_count.so')); R.set_tables(TS)
for name in ("C3", "C2", "C4"):
    cfg = S.named_config(name); cfg.ni, cfg.nj = 64, 48
    if name == "C4": cfg.glacier_frac = 0.3; cfg.water_frac = 0.0
    xp, st, state = make_case(cfg, T)
    sa, sb = clone_state(state), clone_state(state)
    buf = (C.c_ulonglong*17)()
    oc.nmo_opcount_read(buf, 1); R.lib.nmo_opcount_read(buf, 1)
    for step in range(1, 13):
        frc = S.forcing(xp, cfg, step, st)
        arr, sc = S.args_from(cfg, st, frc, sa, step); a = _capi.make_args(arr, sc); stt = _capi.NoahmpStatus()
        oc.nmo_noahmplsm(C.byref(a), C.byref(TS), C.byref(stt), 1, None)
        arr2, sc2 = S.args_from(cfg, st, frc, sb, step); R.noahmplsm(arr2, sc2)
    oc.nmo_opcount_read(buf, 1); o = list(buf)
    R.lib.nmo_opcount_read(buf, 1); r = list(buf)
    n = cfg.ni*cfg.nj*12
    print(name, "per column-step  (oracle | translated reference)")
    for k, nm in enumerate(NAMES):
        if o[k] or r[k]: print("   %-6s %9.1f | %9.1f" % (nm, o[k]/n, r[k]/n))
