"""Noah-MP parameter tables: readers/writers for MPTABLE.TBL, VEGPARM.TBL, SOILPARM.TBL, GENPARM.TBL.

Python restatement of the table semantics of the reference:
  read_mp_veg_parameters   phys/module_sf_noahmplsm.F90:274-404  (Fortran NAMELIST, column-major fill,
                                                                  NVEG<MVT reshape fix-up :373-402)
  SOIL_VEG_GEN_PARM        phys/module_sf_noahmpdrv.F90:1528-1821 (list-directed records)
Used by tests and synthetic-case tooling; the product library has its own C++ reader
(csrc/nmp_tables.cpp) and tests compare the two bit-for-bit.
"""
import json
import os
import re

import numpy as np

from . import _capi

MVT, NLUS, NSLTYPE, NSLOPE = _capi.MVT, _capi.NLUS, _capi.NSLTYPE, _capi.NSLOPE
UNDEF = np.float32(-1.0e36)

_MP_SCALARS = ("isurban", "iswater", "isbarren", "issnow", "eblforest")
_MP_2D = {"rhol": 2, "rhos": 2, "taul": 2, "taus": 2, "saim": 12, "laim": 12, "eps": 5}
_MP_1D = list(_capi._MP_1D) + list(_capi._MP_1D_B) + ["slarea"]
_VEG_COLS = ["shdtbl", "nrotbl", "rstbl", "rgltbl", "hstbl", "snuptbl", "maxalb", "laimintbl", "laimaxtbl",
             "emissmintbl", "emissmaxtbl", "albedomintbl", "albedomaxtbl", "z0mintbl", "z0maxtbl", "ztopvtbl",
             "zbotvtbl"]
_SOIL_COLS = list(_capi._SOIL_F)
_GEN_SCALARS = list(_capi._GEN_F)


# ----------------------------------------------------------------------------------------------
# Fortran NAMELIST (the subset MPTABLE.TBL uses: '!' comments, NAME = v, v, ... with continuation)
# ----------------------------------------------------------------------------------------------
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "\"'":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out)


def _namelist_groups(text):
    groups, cur, name = {}, None, None
    for raw in text.splitlines():
        line = _strip_comment(raw).strip()
        if not line:
            continue
        if cur is None:
            m = re.match(r"&\s*(\w+)", line)
            if m:
                name, cur = m.group(1).lower(), []
                rest = line[m.end():].strip()
                if rest:
                    cur.append(rest)
            continue
        if line.startswith("/"):
            groups[name] = " ".join(cur)
            cur = None
            continue
        cur.append(line)
    return groups


def _namelist_assignments(body):
    """'A = 1, 2 B = "x"' -> {'a': ['1','2'], 'b': ['"x"']} (values as raw tokens, r*c expanded)."""
    toks = re.findall(r"\"[^\"]*\"|'[^']*'|=|[^\s,=]+", body)
    out, cur = {}, None
    i = 0
    while i < len(toks):
        t = toks[i]
        if i + 1 < len(toks) and toks[i + 1] == "=":
            cur = t.lower()
            out[cur] = []
            i += 2
            continue
        if cur is None:
            raise ValueError(f"namelist value {t!r} before any name")
        m = re.match(r"^(\d+)\*(.+)$", t)
        if m:
            out[cur].extend([m.group(2)] * int(m.group(1)))
        else:
            out[cur].append(t)
        i += 1
    return out


def _f32(tok):
    return np.float32(float(tok.replace("d", "e").replace("D", "E")))


def read_mptable(path, dataset="USGS"):
    """read_mp_veg_parameters (noahmplsm.F90:274-404)."""
    if dataset == "USGS":
        gcat, gpar = "noah_mp_usgs_veg_categories", "noah_mp_usgs_parameters"
    elif dataset == "MODIFIED_IGBP_MODIS_NOAH":
        gcat, gpar = "noah_mp_modis_veg_categories", "noah_mp_modis_parameters"
    else:
        raise ValueError("Unrecognized DATASET_IDENTIFIER in subroutine READ_MP_VEG_PARAMETERS: " + dataset)
    with open(path) as f:
        groups = _namelist_groups(f.read())
    cat = _namelist_assignments(groups[gcat])
    par = _namelist_assignments(groups[gpar])
    nveg = int(cat["nveg"][0])
    d = {"nveg": nveg}
    for s in _MP_SCALARS:
        d["isurban_mp" if s == "isurban" else s] = int(par[s][0])
    for n in _MP_1D:
        a = np.full(MVT, UNDEF, np.float32)
        vals = par.get(n, [])
        for k, t in enumerate(vals[:MVT]):
            a[k] = _f32(t)
        d[n] = a
    for n, ncol in _MP_2D.items():
        flat = np.full(MVT * ncol, UNDEF, np.float32)  # Fortran element order of X(MVT,ncol)
        vals = par.get(n, [])
        for k, t in enumerate(vals[:MVT * ncol]):
            flat[k] = _f32(t)
        if MVT > nveg:  # reshape fix-up, noahmplsm.F90:373-402
            new = np.full((ncol, MVT), UNDEF, np.float32)
            new[:, :nveg] = flat[:nveg * ncol].reshape(ncol, nveg)
            d[n] = new
        else:
            d[n] = flat.reshape(ncol, MVT)
    return d


# ----------------------------------------------------------------------------------------------
# list-directed records (VEGPARM / SOILPARM / GENPARM)
# ----------------------------------------------------------------------------------------------
class _Records:
    def __init__(self, path):
        with open(path) as f:
            self.lines = f.read().splitlines()
        self.pos = 0

    def skip(self, n=1):
        self.pos += n
        if self.pos > len(self.lines):
            raise EOFError

    def read(self, n):
        """READ(unit,*) of n items: tokens from successive records until n are found."""
        items = []
        while len(items) < n:
            if self.pos >= len(self.lines):
                raise EOFError
            line = self.lines[self.pos]
            self.pos += 1
            items.extend(re.findall(r"'[^']*'?|\"[^\"]*\"?|[^\s,]+", line))
        return items[:n]

    def eof(self):
        return self.pos >= len(self.lines)


def read_vegparm(path, mminlu="USGS"):
    r = _Records(path)
    d = {}
    while True:
        try:
            r.skip()
            lutype = r.read(1)[0]
            lucats, _ = (int(x) for x in r.read(2))
        except EOFError:
            raise ValueError(f"Land Use Dataset '{mminlu}' not found in VEGPARM.TBL.")
        if lutype == mminlu:
            break
        r.skip(lucats + 12)
    if lucats > NLUS:
        raise ValueError("Table sizes too small for value of LUCATS")
    d["lucats"] = lucats
    cols = {n: np.zeros(NLUS, np.int32 if n == "nrotbl" else np.float32) for n in _VEG_COLS}
    for lc in range(lucats):
        toks = r.read(1 + len(_VEG_COLS))
        for n, t in zip(_VEG_COLS, toks[1:]):
            cols[n][lc] = int(float(t)) if n == "nrotbl" else _f32(t)
    d.update(cols)
    for n in ("topt_data", "cmcmax_data", "cfactr_data", "rsmax_data"):
        r.skip()
        d[n] = _f32(r.read(1)[0])
    for n in ("bare", "natural"):
        r.skip()
        d[n] = int(r.read(1)[0])
    return d


def read_soilparm(path, mminsl="STAS"):
    r = _Records(path)
    r.skip()
    sltype = r.lines[r.pos][:4]  # FORMAT(A4)
    r.skip()
    slcats, _ = (int(x) for x in r.read(2))
    if sltype.strip() != mminsl:
        raise ValueError("INCONSISTENT OR MISSING SOILPARM FILE")
    if slcats > NSLTYPE:
        raise ValueError("Table sizes too small for value of SLCATS")
    d = {"slcats": slcats}
    cols = {n: np.zeros(NSLTYPE, np.float32) for n in _SOIL_COLS}
    for lc in range(slcats):
        toks = r.read(1 + len(_SOIL_COLS))
        for n, t in zip(_SOIL_COLS, toks[1:]):
            cols[n][lc] = _f32(t)
    d.update(cols)
    return d


def read_genparm(path):
    r = _Records(path)
    r.skip(2)
    num_slope = int(r.read(1)[0])
    if num_slope > NSLOPE:
        raise ValueError("NUM_SLOPE too large for slope_data array")
    d = {"slpcats": num_slope, "slope_data": np.zeros(NSLOPE, np.float32)}
    for lc in range(num_slope):
        d["slope_data"][lc] = _f32(r.read(1)[0])
    for n in _GEN_SCALARS:
        r.skip()
        d[n] = _f32(r.read(1)[0])
    return d


def read_tables(directory, dataset="USGS", soil="STAS"):
    """All four tables -> dict keyed like the fields of noahmp_tables (include/noahmp_b200.h)."""
    d = {}
    d.update(read_mptable(os.path.join(directory, "MPTABLE.TBL"), dataset))
    d.update(read_vegparm(os.path.join(directory, "VEGPARM.TBL"), dataset))
    d.update(read_soilparm(os.path.join(directory, "SOILPARM.TBL"), soil))
    d.update(read_genparm(os.path.join(directory, "GENPARM.TBL")))
    return d


# ----------------------------------------------------------------------------------------------
# writers: emit files both the Fortran reference and our readers accept (used to materialise the
# committed fixture tests/golden/tables_*.json as .TBL files in a run directory)
# ----------------------------------------------------------------------------------------------
def _fmt(x):
    x = float(x)
    return repr(np.float32(x).item()) if np.isfinite(x) else str(x)


def _g(x):
    """shortest decimal that round-trips the float32"""
    return np.format_float_scientific(np.float32(x), unique=True, trim="0") if (
        abs(x) >= 1e6 or (x != 0 and abs(x) < 1e-4)) else np.format_float_positional(
        np.float32(x), unique=True, trim="0")


def write_tables(directory, d, dataset="USGS", soil="STAS"):
    os.makedirs(directory, exist_ok=True)
    nveg = int(d["nveg"])
    tag = "usgs" if dataset == "USGS" else "modis"
    with open(os.path.join(directory, "MPTABLE.TBL"), "w") as f:
        f.write(f"&noah_mp_{tag}_veg_categories\n VEG_DATASET_DESCRIPTION = \"{dataset}\"\n NVEG = {nveg}\n/\n")
        f.write(f"&noah_mp_{tag}_parameters\n")
        f.write(f" ISURBAN = {int(d['isurban_mp'])}\n ISWATER = {int(d['iswater'])}\n"
                f" ISBARREN = {int(d['isbarren'])}\n ISSNOW = {int(d['issnow'])}\n"
                f" EBLFOREST = {int(d['eblforest'])}\n")
        for n in _MP_1D:
            vals = [v for v in np.asarray(d[n])[:nveg] if v != UNDEF]
            f.write(f" {n.upper()} = " + ", ".join(_g(v) for v in vals) + ",\n")
        for n, ncol in _MP_2D.items():
            a = np.asarray(d[n])
            f.write(f" {n.upper()} = ")
            for k in range(ncol):
                f.write(("          " if k else "") + ", ".join(_g(v) for v in a[k, :nveg]) + ",\n")
        f.write("/\n")
    with open(os.path.join(directory, "VEGPARM.TBL"), "w") as f:
        f.write("Vegetation Parameters\n" + dataset + "\n")
        f.write(f"{int(d['lucats'])},1, 'SHDFAC NROOT RS RGL HS SNUP MAXALB LAIMIN LAIMAX EMISSMIN EMISSMAX "
                "ALBEDOMIN ALBEDOMAX Z0MIN Z0MAX ZTOPV ZBOTV'\n")
        for lc in range(int(d["lucats"])):
            row = [str(int(d[n][lc])) if n == "nrotbl" else _g(d[n][lc]) for n in _VEG_COLS]
            f.write(f"{lc + 1}, " + ", ".join(row) + f", 'category {lc + 1}'\n")
        for n in ("topt_data", "cmcmax_data", "cfactr_data", "rsmax_data"):
            f.write(n.upper() + "\n" + _g(d[n]) + "\n")
        f.write("BARE\n%d\nNATURAL\n%d\n" % (int(d["bare"]), int(d["natural"])))
    with open(os.path.join(directory, "SOILPARM.TBL"), "w") as f:
        f.write("Soil Parameters\n" + soil + "\n")
        f.write(f"{int(d['slcats'])},1 'BB DRYSMC F11 MAXSMC REFSMC SATPSI SATDK SATDW WLTSMC QTZ'\n")
        for lc in range(int(d["slcats"])):
            f.write(f"{lc + 1}, " + ", ".join(_g(d[n][lc]) for n in _SOIL_COLS) + f", 'soil {lc + 1}'\n")
    with open(os.path.join(directory, "GENPARM.TBL"), "w") as f:
        f.write("General Parameters\nSLOPE_DATA\n%d\n" % int(d["slpcats"]))
        for lc in range(int(d["slpcats"])):
            f.write(_g(d["slope_data"][lc]) + "\n")
        for n in _GEN_SCALARS:
            f.write(n.upper() + "\n" + _g(d[n]) + "\n")


# ----------------------------------------------------------------------------------------------
# JSON fixtures
# ----------------------------------------------------------------------------------------------
def tables_to_json(d, path):
    out = {}
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            out[k] = {"dtype": str(v.dtype), "shape": list(v.shape),
                      "data": [_g(x) if v.dtype == np.float32 else int(x) for x in v.ravel()]}
        else:
            out[k] = _g(v) if isinstance(v, (float, np.floating)) else int(v)
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))


def tables_from_json(path):
    with open(path) as f:
        raw = json.load(f)
    d = {}
    for k, v in raw.items():
        if isinstance(v, dict):
            dt = np.dtype(v["dtype"])
            vals = [np.float32(float(x)) if dt == np.float32 else int(x) for x in v["data"]]
            d[k] = np.array(vals, dtype=dt).reshape(v["shape"])
        elif isinstance(v, str):
            d[k] = np.float32(float(v))
        else:
            d[k] = int(v)
    return d


_DEFAULT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def default_tables(dataset="USGS"):
    """The parameter values of the reference's shipped run/*.TBL, from the committed fixture
    tests/golden/tables_<dataset>.json (generated by tests/golden/gen_tables.py)."""
    tag = "usgs" if dataset == "USGS" else "modis"
    return tables_from_json(os.path.join(_DEFAULT_DIR, f"tables_{tag}.json"))
