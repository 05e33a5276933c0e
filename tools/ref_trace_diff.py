"""Where do the oracle and the translated reference part ways?  (debugging aid of the reference pin)

Runs ONE column of a synthetic case through the value-tracing instantiations of both (oracle/libnmo_opcount.so and
oracle/_ref/libnoahmp_ref_count.so: every fp32 add / multiply / divide with operands and result) and reports the first
arithmetic operation of the reference that the oracle never performed, with the Fortran source line the translator
attached to it and the operations that led up to it.

usage: python tools/ref_trace_diff.py CONFIG STEP J I [math_mode]      (J, I: 0-based numpy indices of the cell)
Needs /root/reference (oracle/ref/build_ref.sh trace)."""
import ctypes as C
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from noahmp_b200 import _capi, synthetic as S, tables  # noqa: E402
from helpers import make_case, clone_state  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.ref import refmodel  # noqa: E402

OPS = ["ADD", "MUL", "DIV"]


class Rec(C.Structure):
    _fields_ = [("op", C.c_int), ("a", C.c_uint), ("b", C.c_uint), ("r", C.c_uint)]


def f(u):
    return struct.unpack("f", struct.pack("I", u))[0]


def one_cell(arr, sc, j, i):
    a = {}
    for n, x in arr.items():
        if x.ndim == 2:
            a[n] = np.ascontiguousarray(x[j:j + 1, i:i + 1])
        elif x.ndim == 3:
            a[n] = np.ascontiguousarray(x[j:j + 1, :, i:i + 1])
        else:
            a[n] = x
    s = dict(sc)
    s.update(ide=1, jde=1, ime=1, jme=1, ite=1, jte=1)
    return a, s


def trace(lib, run):
    lib.nmo_trace_stop.restype = C.c_long
    lib.nmo_trace_start()
    run()
    p = C.c_void_p()
    n = lib.nmo_trace_stop(C.byref(p))
    recs = C.cast(p, C.POINTER(Rec * n)).contents if n else []
    return [(r.op, r.a, r.b, r.r) for r in recs]


def main():
    cfgname, step, j, i = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    T = tables.default_tables("USGS")
    TS = _capi.tables_from_dict(T)
    cfg = S.named_config(cfgname)
    xp, st, state = make_case(cfg, T)
    O.set_math_mode(mode)
    for k in range(1, step):
        frc = S.forcing(xp, cfg, k, st)
        arr, sc = S.args_from(cfg, st, frc, state, k)
        O.noahmplsm(arr, sc, TS, nthreads=1)
    frc = S.forcing(xp, cfg, step, st)
    arr, sc = S.args_from(cfg, st, frc, state, step)
    a1, s1 = one_cell(arr, sc, j, i)
    a2 = {n: (x.copy() if isinstance(x, np.ndarray) else x) for n, x in a1.items()}

    oc = C.CDLL(os.path.join(ROOT, "oracle", "libnmo_opcount.so"))
    oc.nmo_set_math_mode(mode)
    oc.nmo_noahmplsm.argtypes = [C.POINTER(_capi.NoahmpLsmArgs), C.POINTER(_capi.NoahmpTables),
                                 C.POINTER(_capi.NoahmpStatus), C.c_int, C.c_void_p]
    args1 = _capi.make_args(a1, s1)
    stt = _capi.NoahmpStatus()
    t_or = trace(oc, lambda: oc.nmo_noahmplsm(C.byref(args1), C.byref(TS), C.byref(stt), 1, None))

    R = refmodel.RefModel(os.path.join(ROOT, "oracle", "_ref", "libnoahmp_ref_count.so"))
    R.set_tables(TS)
    R.set_math_mode(mode)
    t_ref = trace(R.lib, lambda: R.noahmplsm(a2, s1))

    diff = [n for n in a1 if isinstance(a1[n], np.ndarray) and a1[n].dtype.kind == "f" and
            not np.array_equal(a1[n], a2[n], equal_nan=True)]
    print("oracle ops %d, reference ops %d (+%d markers); fields that differ after the step: %s" %
          (len(t_or), sum(1 for r in t_ref if r[0] >= 0), sum(1 for r in t_ref if r[0] < 0), diff))
    # commutative operations: operands in canonical order
    def key(r):
        op, a, b, res = r
        if op in (0, 1) and a > b:
            a, b = b, a
        return (op, a, b, res)
    have = set(key(r) for r in t_or)
    # values the oracle ever held (operands and results; the magnitude: negation is exact and untraced): an
    # operation of the reference whose RESULT the oracle never saw is a real divergence, whereas one that only
    # differs in form (x/1.0, (-a)/b for -(a/b), a*0.5 for a/2) is not
    vals = set()
    for r in t_or:
        vals.update((r[1] & 0x7fffffff, r[2] & 0x7fffffff, r[3] & 0x7fffffff))
    line, shown, ctx = 0, 0, []
    R.lib.ref_files.restype = C.c_char_p
    files = R.lib.ref_files().decode().split(";")
    src = [open(p, errors="replace").read().split("\n") if os.path.exists(p) else [] for p in files]
    skip = os.environ.get("SKIP", "noahmpdrv")
    last_reported = None
    for r in t_ref:
        if r[0] < 0:
            line = r[1]
            continue
        ctx.append((line, r))
        fid, ln = divmod(line, 100000)
        if (r[3] & 0x7fffffff) not in vals and key(r) not in have and skip not in files[fid] and line != last_reported:
            last_reported = line
            print("\nreference result the oracle never held, at %s:%d" % (os.path.basename(files[fid]), ln))
            if ln <= len(src[fid]):
                print("    | " + src[fid][ln - 1].strip())
            for l2, q in ctx[-8:]:
                print("    line %5d  %s  %-16.9g %-16.9g -> %.9g" % (l2 % 100000, OPS[q[0]], f(q[1]), f(q[2]), f(q[3])))
            shown += 1
            if shown >= int(os.environ.get("NDIFF", "3")):
                break
    if not shown:
        print("every arithmetic operation of the reference was also done by the oracle")


if __name__ == "__main__":
    main()
