"""ctypes mirror of include/noahmp_b200.h (noahmp_tables, noahmp_lsm_args, noahmp_status).

The same structs are consumed by the product library (noahmp_b200/libnoahmp_b200.so) and by the CPU
oracle (oracle/libnmo_oracle.so, test infrastructure). Field order must match the header exactly;
tests/test_abi.py checks sizeof() against both libraries.
"""
import ctypes as C

import numpy as np

MVT, MBAND, NLUS, NSLTYPE, NSLOPE, NSOIL, NSNOW = 27, 2, 50, 30, 30, 4, 3

_f, _i = C.c_float, C.c_int32
_pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int32)

_MP_1D = ("ch2op dleaf z0mvt hvt hvb den rc").split()
_MP_2D_A = ("rhol rhos taul taus").split()
_MP_1D_B = ("xl cwpvt c3psn kc25 akc ko25 ako avcmx aqe ltovrc dilefc dilefw rmf25 sla fragr tmin vcmx25 "
            "tdlef bp mp qe25 rms25 rmr25 arm folnmx wdpool wrrat mrp").split()
_VEG_F = ("shdtbl rstbl rgltbl hstbl snuptbl maxalb laimintbl laimaxtbl emissmintbl emissmaxtbl albedomintbl "
          "albedomaxtbl z0mintbl z0maxtbl ztopvtbl zbotvtbl").split()
_SOIL_F = "bb drysmc f11 maxsmc refsmc satpsi satdk satdw wltsmc qtz".split()
_GEN_F = ("sbeta_data fxexp_data csoil_data salp_data refdk_data refkdt_data frzk_data zbot_data czil_data "
          "smlow_data smhigh_data lvcoef_data").split()


class NoahmpTables(C.Structure):
    _fields_ = (
        [("nveg", _i), ("isurban_mp", _i), ("iswater", _i), ("isbarren", _i), ("issnow", _i), ("eblforest", _i)]
        + [(n, _f * MVT) for n in _MP_1D]
        + [(n, (_f * MVT) * MBAND) for n in _MP_2D_A]
        + [(n, _f * MVT) for n in _MP_1D_B]
        + [("saim", (_f * MVT) * 12), ("laim", (_f * MVT) * 12), ("slarea", _f * MVT), ("eps", (_f * MVT) * 5)]
        + [("lucats", _i), ("nrotbl", _i * NLUS)]
        + [(n, _f * NLUS) for n in _VEG_F]
        + [("topt_data", _f), ("cmcmax_data", _f), ("cfactr_data", _f), ("rsmax_data", _f), ("bare", _i),
           ("natural", _i)]
        + [("slcats", _i)]
        + [(n, _f * NSLTYPE) for n in _SOIL_F]
        + [("slpcats", _i), ("slope_data", _f * NSLOPE)]
        + [(n, _f) for n in _GEN_F]
    )


# (name, kind) kind: i/f scalar, pf/pi pointer; order = noahmp_lsm_args in the header = noahmplsm dummy list
_IN2D = "coszin xlatin".split()
_INOUT21 = ("tsk hfx qfx lh grdflx smstav smstot sfcrunoff udrunoff albedo snowc smois sh2o tslb snow snowh "
            "canwat acsnom acsnow emiss qsfc").split()
_INOUT_MP = ("tvxy tgxy canicexy canliqxy eahxy tahxy cmxy chxy fwetxy sneqvoxy alboldxy qsnowxy wslakexy zwtxy "
             "waxy wtxy tsnoxy zsnsoxy snicexy snliqxy lfmassxy rtmassxy stmassxy woodxy stblcpxy fastcpxy xlaixy "
             "xsaixy taussxy smoiseq smcwtdxy deeprechxy rechxy").split()
OUT44 = ("t2mvxy t2mbxy q2mvxy q2mbxy tradxy neexy gppxy nppxy fvegxy runsfxy runsbxy ecanxy edirxy etranxy "
         "fsaxy firaxy aparxy psnxy savxy sagxy rssunxy rsshaxy bgapxy wgapxy tgvxy tgbxy chvxy chbxy shgxy shcxy "
         "shbxy evgxy evbxy ghvxy ghbxy irgxy ircxy irbxy trxy evcxy chleafxy chucxy chv2xy chb2xy").split()
_OPTS = ("idveg iopt_crs iopt_btr iopt_run iopt_sfc iopt_frz iopt_inf iopt_rad iopt_alb iopt_snf iopt_tbot "
         "iopt_stc iz0tlnd").split()
_BOUNDS = "ids ide jds jde kds kde ims ime jms jme kms kme its ite jts jte kts kte".split()

ARGS_SPEC = (
    [("itimestep", "i"), ("yr", "i"), ("julian", "f"), ("coszin", "pf"), ("xlatin", "pf"), ("dz8w", "pf"),
     ("dt", "f"), ("dzs", "pf"), ("nsoil", "i"), ("dx", "f"), ("ivgtyp", "pi"), ("isltyp", "pi"),
     ("vegfra", "pf"), ("vegmax", "pf"), ("tmn", "pf"), ("xland", "pf"), ("xice", "pf"), ("xice_thres", "f"),
     ("isice", "i"), ("isurban", "i")]
    + [(n, "i") for n in _OPTS]
    + [(n, "pf") for n in "t3d qv3d u_phy v_phy swdown glw p8w3d rainbl".split()]
    + [(n, "pf") for n in _INOUT21]
    + [("isnowxy", "pi")]
    + [(n, "pf") for n in _INOUT_MP]
    + [(n, "pf") for n in OUT44]
    + [(n, "i") for n in _BOUNDS]
)
_K = {"i": _i, "f": _f, "pf": _pf, "pi": _pi}


class NoahmpLsmArgs(C.Structure):
    _fields_ = [(n, _K[k]) for n, k in ARGS_SPEC]


class NoahmpStatus(C.Structure):
    _fields_ = [("code", _i), ("i", _i), ("j", _i), ("count", _i), ("value", _f)]


# array names grouped by layer structure (middle dimension of the Fortran (i,k,j) layout)
LAYERS = {"smois": 4, "sh2o": 4, "tslb": 4, "smoiseq": 4, "tsnoxy": 3, "snicexy": 3, "snliqxy": 3, "zsnsoxy": 7}
ATM3D = ("dz8w", "t3d", "qv3d", "u_phy", "v_phy", "p8w3d")  # (i, kms:kme, j)
INT_ARRAYS = ("ivgtyp", "isltyp", "isnowxy")
INOUT_NAMES = _INOUT21 + ["isnowxy"] + _INOUT_MP
OUT_NAMES = list(OUT44)
IN_ARRAY_NAMES = ["coszin", "xlatin", "dz8w", "ivgtyp", "isltyp", "vegfra", "vegmax", "tmn", "xland", "xice",
                  "t3d", "qv3d", "u_phy", "v_phy", "swdown", "glw", "p8w3d", "rainbl"]
ARRAY_NAMES = [n for n, k in ARGS_SPEC if k in ("pf", "pi") and n != "dzs"]


def array_shape(name, ni, nj, nk=2):
    """numpy shape (C order) of a Fortran (i[,k],j) array: (nj[,nk],ni)."""
    if name in LAYERS:
        return (nj, LAYERS[name], ni)
    if name in ATM3D:
        return (nj, nk, ni)
    return (nj, ni)


def array_dtype(name):
    return np.int32 if name in INT_ARRAYS else np.float32


ARG_POINTER_TYPE = {n: _K[k] for n, k in ARGS_SPEC if k in ("pf", "pi")}


def make_args(arrays, scalars):
    """Build a NoahmpLsmArgs from dicts. `arrays` values must be C-contiguous numpy arrays with the
    shapes of array_shape(); the struct keeps no reference, so keep `arrays` alive during the call."""
    a = NoahmpLsmArgs()
    for n, k in ARGS_SPEC:
        if k in ("i", "f"):
            setattr(a, n, scalars[n])
        else:
            arr = arrays[n]
            want = np.int32 if k == "pi" else np.float32
            if arr.dtype != want or not arr.flags["C_CONTIGUOUS"]:
                raise TypeError(f"{n}: need C-contiguous {want.__name__}, got {arr.dtype}")
            setattr(a, n, arr.ctypes.data_as(_K[k]))
    return a


def tables_from_dict(d):
    """dict of numpy arrays / scalars (see tables.py) -> NoahmpTables."""
    t = NoahmpTables()
    for name, ctype in NoahmpTables._fields_:
        v = d[name]
        if ctype in (_i, _f):
            setattr(t, name, v.item() if hasattr(v, "item") else v)
        else:
            arr = np.ascontiguousarray(v, dtype=np.int32 if name == "nrotbl" else np.float32)
            dst = np.ctypeslib.as_array(getattr(t, name))
            if dst.shape != arr.shape:
                raise ValueError(f"{name}: shape {arr.shape} != {dst.shape}")
            dst[...] = arr
    return t


def tables_to_dict(t):
    d = {}
    for name, ctype in NoahmpTables._fields_:
        v = getattr(t, name)
        d[name] = v if ctype in (_i, _f) else np.array(np.ctypeslib.as_array(v))
    return d


# ---- opt_run = 5 groundwater: noahmp_wtable_args (WTABLE_mmf_noahmp dummy list, groundwater.F90:14-22) -----------
WT_SPEC = (
    [("nsoil", "i"), ("xland", "pf"), ("xice", "pf"), ("xice_threshold", "f"), ("isice", "i"), ("isltyp", "pi"),
     ("smoiseq", "pf"), ("dzs", "pf"), ("wtddt", "f"), ("fdepth", "pf"), ("area", "pf"), ("topo", "pf"),
     ("isurban", "i"), ("ivgtyp", "pi"), ("rivercond", "pf"), ("riverbed", "pf"), ("eqwtd", "pf"), ("pexp", "pf")]
    + [(n, "pf") for n in "smois sh2oxy smcwtd wtd qrf deeprech qspring qslat qrfs qsprings rech".split()]
    + [(n, "i") for n in _BOUNDS]
)
WT_STATIC = ["fdepth", "area", "topo", "rivercond", "riverbed", "eqwtd", "pexp"]
WT_INOUT = ["smois", "sh2oxy", "smcwtd", "wtd", "deeprech", "qslat", "qrfs", "qsprings", "rech"]
WT_OUT = ["qrf", "qspring"]


class NoahmpWtableArgs(C.Structure):
    _fields_ = [(n, _K[k]) for n, k in WT_SPEC]


def make_wtable_args(arrays, scalars):
    a = NoahmpWtableArgs()
    for n, k in WT_SPEC:
        if k in ("i", "f"):
            setattr(a, n, scalars[n])
        else:
            arr = arrays[n]
            want = np.int32 if k == "pi" else np.float32
            if arr.dtype != want or not arr.flags["C_CONTIGUOUS"]:
                raise TypeError(f"{n}: need C-contiguous {want.__name__}, got {arr.dtype}")
            setattr(a, n, arr.ctypes.data_as(_K[k]))
    return a


# ---- cold start (row f1): noahmp_init_args (NOAHMP_INIT dummy list, noahmpdrv.F90:847-864) ------------------------
INIT_SPEC = (
    [("snow", "pf"), ("snowh", "pf"), ("canwat", "pf"), ("isltyp", "pi"), ("ivgtyp", "pi"), ("isurban", "i"),
     ("tslb", "pf"), ("smois", "pf"), ("sh2o", "pf"), ("dzs", "pf"), ("fndsoilw", "i"), ("fndsnowh", "i"),
     ("isice", "i"), ("iswater", "i"), ("tsk", "pf"), ("isnowxy", "pi"), ("tvxy", "pf"), ("tgxy", "pf"),
     ("canicexy", "pf"), ("tmn", "pf"), ("xice", "pf")]
    + [(n, "pf") for n in ("canliqxy eahxy tahxy cmxy chxy fwetxy sneqvoxy alboldxy qsnowxy wslakexy zwtxy waxy wtxy "
                           "tsnoxy zsnsoxy snicexy snliqxy lfmassxy rtmassxy stmassxy woodxy stblcpxy fastcpxy xsaixy "
                           "t2mvxy t2mbxy chstarxy").split()]
    + [("nsoil", "i"), ("restart", "i"), ("allowed_to_read", "i"), ("iopt_run", "i")]
    + [(n, "i") for n in _BOUNDS]
    + [(n, "pf") for n in "smoiseq smcwtdxy rechxy deeprechxy areaxy".split()]
    + [("dx", "f"), ("dy", "f"), ("msftx", "pf"), ("msfty", "pf"), ("wtddt", "f"), ("stepwtd", "pi"), ("dt", "f")]
    + [(n, "pf") for n in "qrfsxy qspringsxy qslatxy fdepthxy ht riverbedxy eqzwt rivercondxy pexpxy".split()]
)
INIT_GW = ("smoiseq smcwtdxy rechxy deeprechxy areaxy msftx msfty stepwtd qrfsxy qspringsxy qslatxy fdepthxy ht "
           "riverbedxy eqzwt rivercondxy pexpxy").split()
INIT_LAYERS = {"tslb": 4, "smois": 4, "sh2o": 4, "smoiseq": 4, "tsnoxy": 3, "snicexy": 3, "snliqxy": 3, "zsnsoxy": 7}


class NoahmpInitArgs(C.Structure):
    _fields_ = [(n, _K[k]) for n, k in INIT_SPEC]


def make_init_args(arrays, scalars):
    """arrays: numpy arrays by NOAHMP_INIT dummy name (groundwater ones may be absent -> NULL); scalars: the rest."""
    a = NoahmpInitArgs()
    for n, k in INIT_SPEC:
        if k in ("i", "f"):
            setattr(a, n, scalars.get(n, 0))
        elif n in arrays and arrays[n] is not None:
            arr = arrays[n]
            want = np.int32 if k == "pi" else np.float32
            if arr.dtype != want or not arr.flags["C_CONTIGUOUS"]:
                raise TypeError(f"{n}: need C-contiguous {want.__name__}, got {arr.dtype}")
            setattr(a, n, arr.ctypes.data_as(_K[k]))
        elif n not in INIT_GW and n != "tmn":
            raise KeyError(n)
    return a


# ---- on-device forcing pipeline (row f2): noahmp_forcing_fields ---------------------------------------------------
FORCING_FIELDS = "t q u v p lw sw pcp fpar".split()


class NoahmpForcingFields(C.Structure):
    _fields_ = [(n, _pf) for n in FORCING_FIELDS]


def make_forcing_fields(d):
    f = NoahmpForcingFields()
    for n in FORCING_FIELDS:
        arr = d[n]
        if arr.dtype != np.float32 or not arr.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{n}: need C-contiguous float32")
        setattr(f, n, arr.ctypes.data_as(_pf))
    return f
