"""ctypes loader for oracle/_ref/libnoahmp_ref.so — the reference's own Fortran text, machine-translated to C++ by
oracle/ref/f90cxx.py and compiled with g++ (TEST INFRASTRUCTURE; only tests/ and tools/ may import this).

Every module procedure of the translated files is callable by name with the reference's own positional argument list:
    ref.call("ESAT", t, esw, esi, desw, desi)        # scalars: Python numbers in, updated values in the returned list
    ref.call("ROSR12", P, A, B, C, D, DELTA, ntop, nsoil, nsnow)   # arrays: numpy, updated in place
and the module variables (parameter tables, option switches) are reachable by MODULE.NAME.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _REPO)
from noahmp_b200 import _capi  # noqa: E402

SO = os.path.join(_REPO, "oracle", "_ref", "libnoahmp_ref.so")
REFERENCE = "/root/reference"


def available():
    return os.path.exists(SO)


def build(force=False):
    """translate + compile (only where the reference tree exists); returns the path of the library or None"""
    if os.path.exists(SO) and not force:
        return SO
    if not os.path.isdir(REFERENCE):
        return None
    subprocess.check_call(["make", "-C", os.path.join(_REPO, "oracle"), "-s", "ref"])
    return SO if os.path.exists(SO) else None


class _Var(C.Structure):
    _fields_ = [("name", C.c_char_p), ("p", C.c_void_p), ("bytes", C.c_ulong), ("type", C.c_char)]


_NP = {"i": np.int32, "r": np.float32, "d": np.float64, "l": np.bool_}
_CT = {"i": C.c_int32, "r": C.c_float, "d": C.c_double, "l": C.c_bool}


class RefModel:
    def __init__(self, path=SO):
        self.lib = C.CDLL(path)
        self.lib.ref_vars.restype = C.POINTER(_Var)
        self.vars = {}
        tab = self.lib.ref_vars()
        i = 0
        while tab[i].name:
            v = tab[i]
            self.vars[v.name.decode()] = (v.p, int(v.bytes), v.type.decode())
            i += 1
        self.by_short = {}
        for k in self.vars:
            self.by_short.setdefault(k.split(".", 1)[1], []).append(k)
        self._sig = {}

    # ---- module variables -------------------------------------------------------------------------------------
    def var(self, name):
        """numpy view of a module variable (MODULE.NAME, or NAME when unique)"""
        if "." not in name:
            (name,) = self.by_short[name]
        p, nbytes, t = self.vars[name]
        dt = np.dtype(_NP[t])
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(_CT[t])), shape=(nbytes // dt.itemsize,))

    def set_math_mode(self, mode):
        self.lib.ref_set_math_mode(int(mode))

    def set_tables(self, tables):
        """copy a noahmp_tables struct (the product's table reader output) into the module variables of the same
        names; returns (fields without a module variable, names set)"""
        missing, done = [], []
        for fname, ctype in _capi.NoahmpTables._fields_:
            names = self.by_short.get(fname.upper())
            if not names:
                missing.append(fname)
                continue
            src = np.frombuffer(bytes(memoryview(C.cast(C.byref(tables, getattr(_capi.NoahmpTables, fname).offset),
                                                        C.POINTER(C.c_char * C.sizeof(ctype))).contents)), np.uint8)
            for n in names:
                dst = self.var(n).view(np.uint8)
                k = min(dst.size, src.size)
                # the reference dimensions some tables larger (NLUS = 50 entries) than the struct; the leading part is
                # what the table files fill
                dst[:k] = src[:k]
                done.append(n)
        return missing, done

    # ---- procedures ---------------------------------------------------------------------------------------------
    def signature(self, name):
        if name not in self._sig:
            f = getattr(self.lib, "ref_sig_" + name)
            f.restype = C.c_char_p
            self._sig[name] = [tuple(x.split(":")) for x in f().decode().split(",") if x]
        return self._sig[name]

    def call(self, name, *args):
        """-> list of the scalar arguments after the call (arrays are updated in place); raises on wrf_error_fatal"""
        sig = self.signature(name)
        if len(args) != len(sig):
            raise TypeError("%s takes %d arguments (%s), %d given" % (name, len(sig), ",".join(a for a, _ in sig), len(args)))
        argv = (C.c_void_p * len(sig))()
        keep, scalars = [], []
        for i, ((an, at), v) in enumerate(zip(sig, args)):
            if at[0].isupper():  # array
                if v is None:
                    argv[i] = None
                    scalars.append(None)
                    continue
                want = _NP[at[0].lower()]
                if not isinstance(v, np.ndarray) or v.dtype != want:
                    raise TypeError("%s: argument %s must be a numpy array of %s" % (name, an, want.__name__))
                if not (v.flags["F_CONTIGUOUS"] or v.flags["C_CONTIGUOUS"]):
                    raise TypeError("%s: argument %s is not contiguous" % (name, an))
                argv[i] = v.ctypes.data
                keep.append(v)
                scalars.append(None)
            elif at == "c":
                b = C.create_string_buffer(str(v).encode())
                keep.append(b)
                argv[i] = C.addressof(b)
                scalars.append(None)
            else:
                if v is None:
                    argv[i] = None
                    scalars.append(None)
                    continue
                c = _CT[at](v)
                keep.append(c)
                argv[i] = C.addressof(c)
                scalars.append(c)
        msg = C.create_string_buffer(512)
        rc = getattr(self.lib, "ref_call_" + name)(argv, msg, 512)
        if rc:
            raise RuntimeError("%s: %s" % (name, msg.value.decode(errors="replace")))
        return [None if s is None else s.value for s in scalars]

    def call_struct(self, name, struct, extra=None, strict=True):
        """call procedure `name` taking each dummy argument from the member of the same (lower-case) name of a ctypes
        struct of the C-ABI (include/noahmp_b200.h), or from `extra`; a NULL pointer member = an absent OPTIONAL
        argument.  strict: the struct's members must be the dummy list, in order (the drop-in boundary)."""
        sig = self.signature(name)
        fields = [n for n, _ in struct._fields_]
        extra = extra or {}
        if strict:
            want = [n.lower() for n, _ in sig if n.lower() not in extra]
            if fields != want:
                raise RuntimeError("%s: the struct and the reference's dummy list differ: %r" %
                                   (name, [(x, y) for x, y in zip(fields, want) if x != y][:5]))
        argv = (C.c_void_p * len(sig))()
        keep = []
        for i, (an, at) in enumerate(sig):
            n = an.lower()
            if n in extra:
                v = extra[n]
            else:
                v = getattr(struct, n)
            if at[0].isupper():
                argv[i] = C.cast(v, C.c_void_p).value if not isinstance(v, np.ndarray) else v.ctypes.data
            elif at == "c":
                b = C.create_string_buffer(str(v).encode())
                keep.append(b)
                argv[i] = C.addressof(b)
            else:
                if isinstance(v, C._Pointer):  # scalar passed by address (INTENT(OUT))
                    argv[i] = C.cast(v, C.c_void_p).value
                    continue
                c = _CT[at](bool(v) if at == "l" else v)
                keep.append(c)
                argv[i] = C.addressof(c)
        msg = C.create_string_buffer(512)
        rc = getattr(self.lib, "ref_call_" + name)(argv, msg, 512)
        if rc:
            raise RuntimeError("%s: %s" % (name, msg.value.decode(errors="replace")))

    def noahmplsm(self, arrays, scalars):
        """the reference's `noahmplsm` (module_sf_noahmpdrv) on host arrays in the product's own layout; arrays are
        updated in place"""
        self.call_struct("NOAHMPLSM", _capi.make_args(arrays, scalars))

    def init(self, arrays, scalars, mminlu="USGS"):
        """the reference's NOAHMP_INIT (its table readers left out: set_tables() does their work)"""
        arrays = dict(arrays)
        step = np.zeros(1, np.int32)
        if scalars.get("iopt_run") == 5:
            arrays["stepwtd"] = step
        self.call_struct("NOAHMP_INIT", _capi.make_init_args(arrays, scalars), extra={"mminlu": mminlu})
        return int(step[0]) if scalars.get("iopt_run") == 5 else None

    def wtable(self, arrays, scalars):
        """the reference's WTABLE_mmf_noahmp (module_sf_noahmp_groundwater)"""
        self.call_struct("WTABLE_MMF_NOAHMP", _capi.make_wtable_args(arrays, scalars))
