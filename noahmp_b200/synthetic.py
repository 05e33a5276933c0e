"""Synthetic grids, forcing and cold-start state for the Noah-MP column step.

Everything is a pure function of (seed, global column index, step, field id) through a counter-based
integer hash, so a tile of a larger domain gets exactly the values the whole-domain run would give
it (needed for the multi-GPU weak-scaling bench and the tiling tests).  The same code runs on numpy
(tests, CPU oracle) and on torch tensors (bench: forcing generated directly in HBM).

The cold start restates NOAHMP_INIT / SNOW_INIT (phys/module_sf_noahmpdrv.F90:989-1120, :1182-1283)
and the driver's first-step guesses (driver/module_hrldas_noahmp_driver.F90:374-384); the cosine of
the zenith angle follows CALC_DECLIN (:813-863).
"""
import math

import numpy as np

from . import _capi

SEED = 20240601
MASK32 = 0xFFFFFFFF

# field ids for the hash
(F_WATER, F_VEG, F_VEGSEL, F_SOIL, F_TMN, F_VEGFRA, F_HGT, F_SMOIS, F_SNOW, F_SNODEP,
 F_T, F_RH, F_U1, F_U2, F_V1, F_V2, F_GLW, F_SW, F_RAINP, F_RAINA, F_FDEPTH, F_EQWTD, F_RCOND, F_WTD0,
 F_SMCWTD) = range(25)


class _NP:
    name = "numpy"

    def arange(self, n):
        return np.arange(n, dtype=np.int64)

    def i64(self, x):
        return x.astype(np.int64)

    def f32(self, x):
        return x.astype(np.float32)

    def f64(self, x):
        return x.astype(np.float64)

    def where(self, c, a, b):
        return np.where(c, a, b)

    def __getattr__(self, k):
        return getattr(np, k)


class _TORCH:
    name = "torch"

    def __init__(self, device):
        import torch
        self.t = torch
        self.device = device

    def arange(self, n):
        return self.t.arange(n, dtype=self.t.int64, device=self.device)

    def i64(self, x):
        return x.to(self.t.int64)

    def f32(self, x):
        return x.to(self.t.float32)

    def f64(self, x):
        return x.to(self.t.float64)

    def where(self, c, a, b):
        if not self.t.is_tensor(a):
            a = self.t.as_tensor(a, dtype=b.dtype if self.t.is_tensor(b) else None, device=self.device)
        if not self.t.is_tensor(b):
            b = self.t.as_tensor(b, dtype=a.dtype, device=self.device)
        return self.t.where(c, a, b)

    def maximum(self, a, b):
        if not self.t.is_tensor(b):
            return self.t.clamp(a, min=b)
        return self.t.maximum(a, b)

    def minimum(self, a, b):
        if not self.t.is_tensor(b):
            return self.t.clamp(a, max=b)
        return self.t.minimum(a, b)

    def clip(self, a, lo, hi):
        return self.t.clamp(a, lo, hi)

    def __getattr__(self, k):
        return getattr(self.t, k)


def backend(device=None):
    return _NP() if device is None else _TORCH(device)


def _fmix32(xp, h):
    """murmur3 finaliser on the low 32 bits of int64 lanes (wrap-around multiply is exact mod 2^64)."""
    h = h & MASK32
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & MASK32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & MASK32
    h = h ^ (h >> 16)
    return h


def uniform(xp, g, step, field, seed=SEED):
    """U[0,1) float32 with 24 random bits; g = int64 global column indices."""
    k = (seed * 0x9E3779B1 + (step + 1) * 0x85EBCA77 + (field + 1) * 0xC2B2AE3D) & MASK32
    h = _fmix32(xp, (g & MASK32) ^ k)
    h = _fmix32(xp, h + ((g >> 32) & MASK32) * 0x27D4EB2F + 0x165667B1)
    return xp.f32(h >> 8) * np.float32(1.0 / 16777216.0)


def normal(xp, g, step, f1, f2, seed=SEED):
    u1 = xp.f64(uniform(xp, g, step, f1, seed)) + 2.0 ** -25
    u2 = xp.f64(uniform(xp, g, step, f2, seed))
    return xp.f32(xp.sqrt(-2.0 * xp.log(u1)) * xp.cos(2.0 * math.pi * u2))


class Config:
    """One of the named workloads (BASELINE.json configs / SURVEY.md §8d C1..C4)."""

    def __init__(self, name, ni, nj, dveg=4, opt_run=1, water_frac=0.0, glacier_frac=0.0, snow_frac=0.0,
                 t_base=283.0, start=(2017, 5, 1, 0), lat=(25.0, 50.0), lon=(-125.0, -67.0), dt=3600.0,
                 opts=None):
        self.name, self.ni, self.nj = name, ni, nj
        self.water_frac, self.glacier_frac, self.snow_frac = water_frac, glacier_frac, snow_frac
        self.t_base, self.start, self.lat, self.lon, self.dt = t_base, start, lat, lon, dt
        o = dict(idveg=dveg, iopt_crs=1, iopt_btr=1, iopt_run=opt_run, iopt_sfc=1, iopt_frz=1, iopt_inf=1,
                 iopt_rad=3, iopt_alb=2, iopt_snf=1, iopt_tbot=2, iopt_stc=1, iz0tlnd=0)
        if opts:
            o.update(opts)
        self.opts = o


def named_config(name):
    if name == "C1":
        return Config("C1", 10, 10, dveg=4)
    if name == "C2":
        return Config("C2", 464, 224, dveg=4, water_frac=0.10)
    if name == "C3":
        return Config("C3", 4608, 3840, dveg=2, snow_frac=0.40, t_base=263.0, start=(2017, 1, 15, 0))
    if name == "C5":  # the C3 grid with the Miguez-Macho & Fan groundwater scheme
        return Config("C5", 4608, 3840, dveg=2, opt_run=5, snow_frac=0.40, t_base=263.0, start=(2017, 1, 15, 0))
    if name == "C4":
        return Config("C4", 7200, 3600, dveg=4, water_frac=0.69, glacier_frac=0.10, lat=(-60.0, 75.0),
                      lon=(-180.0, 180.0))
    raise KeyError(name)


DZS = np.array([0.1, 0.3, 0.6, 1.0], np.float32)  # run/noahmp.namelist soil_layer_thickness
ISICE, ISURBAN, ISWATER = 24, 1, 16                # USGS (const-file global attributes)
# 20 vegetated USGS classes (2-15, 17-18, 20-23); urban (1), barren (19) and ice (24) are drawn apart
_VEG_POOL = np.array([2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 17, 18, 20, 21, 22, 23], np.int64)


def _days_before(y, m, d):
    mdays = [31, 29 if (y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)) else 28, 31, 30, 31, 30, 31, 31, 30, 31,
             30, 31]
    return sum(mdays[:m - 1]) + (d - 1)


def clock(cfg, step):
    """(yr, julian, hour) of model step `step` (1-based; forcing valid at start + (step-1)*dt)."""
    y, m, d, h = cfg.start
    hours = h + (step - 1) * cfg.dt / 3600.0
    day = _days_before(y, m, d) + int(hours // 24)
    hour = hours - 24.0 * (hours // 24)
    julian = np.float32(np.float32(day) + np.float32(int(hour)) / np.float32(24.0))
    return y, julian, hour


def tile_index(xp, cfg, xs, xe, ys, ye, jstride=1):
    """int64 global column index g (0-based, i fastest) of the tile [xs..xe]x[ys..ye] (1-based incl); with
    jstride > 1 only every jstride-th row of it (a row sample of the domain for the CPU arms of bench.py)."""
    ii = xp.arange(xe - xs + 1) + (xs - 1)
    jj = xp.arange((ye - ys) // jstride + 1) * jstride + (ys - 1)
    g = jj[:, None] * cfg.ni + ii[None, :]
    return g, ii, jj


def static_fields(xp, cfg, xs=1, xe=None, ys=1, ye=None, jstride=1):
    xe = xe or cfg.ni
    ye = ye or cfg.nj
    g, ii, jj = tile_index(xp, cfg, xs, xe, ys, ye, jstride)
    nj, ni = g.shape
    u = lambda f: uniform(xp, g, -1, f)
    water = u(F_WATER) < cfg.water_frac
    uv = u(F_VEG)
    sel = xp.i64(u(F_VEGSEL) * len(_VEG_POOL))
    pool = _VEG_POOL if xp.name == "numpy" else xp.as_tensor(_VEG_POOL, device=g.device)
    veg = pool[xp.clip(sel, 0, len(_VEG_POOL) - 1)]
    veg = xp.where(uv < 0.03, 1, veg)
    veg = xp.where((uv >= 0.03) & (uv < 0.05), 19, veg)
    veg = xp.where((uv >= 0.05) & (uv < 0.05 + cfg.glacier_frac), ISICE, veg)
    soil = 1 + xp.clip(xp.i64(u(F_SOIL) * 12), 0, 11)
    soil = xp.where(veg == ISICE, 16, soil)
    veg = xp.where(water, ISWATER, veg)
    soil = xp.where(water, 14, soil)
    tmn = 275.0 + 20.0 * u(F_TMN)
    tmn = xp.where(veg == ISICE, 260.0, xp.f64(tmn))
    lat = cfg.lat[0] + (cfg.lat[1] - cfg.lat[0]) * xp.f64(jj) / max(cfg.nj - 1, 1)
    lon = cfg.lon[0] + (cfg.lon[1] - cfg.lon[0]) * xp.f64(ii) / max(cfg.ni - 1, 1)
    lat2 = lat[:, None] + 0.0 * xp.f64(g)
    lon2 = lon[None, :] + 0.0 * xp.f64(g)
    vegfra = 5.0 + 90.0 * u(F_VEGFRA)
    hgt = 2500.0 * u(F_HGT)
    i32 = (lambda x: x.astype(np.int32)) if xp.name == "numpy" else (lambda x: x.to(xp.t.int32))
    s = {
        "ivgtyp": i32(veg), "isltyp": i32(soil), "xland": xp.f32(xp.where(water, 2.0, 1.0 + 0.0 * xp.f64(g))),
        "xice": xp.f32(0.0 * xp.f64(g)), "tmn": xp.f32(tmn), "xlatin": xp.f32(lat2), "xlong": xp.f32(lon2),
        "vegfra": xp.f32(vegfra), "vegmax": xp.f32(xp.maximum(xp.f64(vegfra), 50.0)), "hgt": xp.f32(hgt),
    }
    s["_g"] = g
    return s


def cosz_julian(xp, cfg, step, lat, lon):
    """CALC_DECLIN (driver/module_hrldas_noahmp_driver.F90:813-863), fp32 like the reference."""
    yr, julian, hour = clock(cfg, step)
    f = np.float32
    DEGRAD = f(f(3.14159265) / f(180.0))
    DPD = f(f(360.0) / f(365.0))
    OBECL = f(f(23.5) * DEGRAD)
    SINOB = f(math.sin(OBECL))
    if julian >= 80.0:
        SXLONG = f(f(DPD * f(julian - f(80.0))) * DEGRAD)
    else:
        SXLONG = f(f(DPD * f(julian + f(285.0))) * DEGRAD)
    ARG = f(SINOB * f(math.sin(SXLONG)))
    DECLIN = f(math.asin(ARG))
    ihour = int(hour)
    tloc = xp.f32(f(ihour) + lon / f(15.0))
    tloc = xp.f32(xp.fmod(tloc + f(24.0), f(24.0)))
    hrang = xp.f32(f(15.0) * (tloc - f(12.0)) * DEGRAD)
    latr = xp.f32(lat * DEGRAD)
    cosz = xp.f32(xp.sin(latr) * f(math.sin(DECLIN)) + xp.cos(latr) * f(math.cos(DECLIN)) * xp.cos(hrang))
    return cosz, yr, julian


def forcing(xp, cfg, step, st):
    """Forcing for model step `step` (1-based). Returns dict in 2-D (nj,ni) float32 plus scalars."""
    g = st["_g"]
    u = lambda f: uniform(xp, g, step, f)
    cosz, yr, julian = cosz_julian(xp, cfg, step, st["xlatin"], st["xlong"])
    _, _, hour = clock(cfg, step)
    latr = xp.f64(st["xlatin"]) * (math.pi / 180.0)
    tloc = hour + xp.f64(st["xlong"]) / 15.0
    diurnal = xp.cos((tloc - 15.0) * (2.0 * math.pi / 24.0))
    T = cfg.t_base + 12.0 * (xp.cos(latr) - 0.75) + 6.0 * diurnal + (-2.0 + 4.0 * xp.f64(u(F_T)))
    P = 101325.0 * xp.exp(-xp.f64(st["hgt"]) / 8400.0)
    es = 611.2 * xp.exp(17.67 * (T - 273.15) / (T - 29.65))
    qs = 0.622 * es / (P - 0.378 * es)
    rh = 0.3 + 0.6 * xp.f64(u(F_RH))
    q = rh * qs
    qv = q / (1.0 - q)  # mixing ratio, as QV3D
    U = 3.0 * normal(xp, g, step, F_U1, F_U2)
    V = 3.0 * normal(xp, g, step, F_V1, F_V2)
    glw = 0.85 * 5.67e-8 * T ** 4 * (1.0 + 0.2 * xp.f64(u(F_GLW)))
    sw = 1000.0 * xp.maximum(xp.f64(cosz), 0.0) * (0.3 + 0.7 * xp.f64(u(F_SW)))
    rainrate = xp.where(u(F_RAINP) < 0.85, 0.0 * T, -3.0e-4 * xp.log(1.0 - xp.f64(u(F_RAINA)) + 2.0 ** -26))
    return {
        "coszin": cosz, "t": xp.f32(T), "qv": xp.f32(qv), "u": xp.f32(U), "v": xp.f32(V), "swdown": xp.f32(sw),
        "glw": xp.f32(glw), "p": xp.f32(P), "rainbl": xp.f32(rainrate * cfg.dt), "yr": yr, "julian": julian,
    }


def _snow_init(swe, snodep, tg, zsoil):
    """SNOW_INIT (noahmpdrv.F90:1182-1283), numpy, vectorised over columns (nj,ni)."""
    f = np.float32
    nj, ni = swe.shape
    isnow = np.zeros((nj, ni), np.int32)
    dz = np.zeros((3, nj, ni), np.float32)  # index 0..2 <-> layers -2..0
    sd = snodep
    c1 = (sd >= f(0.025)) & (sd <= f(0.05))
    c2 = (sd > f(0.05)) & (sd <= f(0.10))
    c3 = (sd > f(0.10)) & (sd <= f(0.25))
    c4 = (sd > f(0.25)) & (sd <= f(0.45))
    c5 = sd > f(0.45)
    isnow[c1] = -1
    dz[2][c1] = sd[c1]
    isnow[c2] = -2
    dz[1][c2] = sd[c2] / f(2.0)
    dz[2][c2] = sd[c2] / f(2.0)
    isnow[c3] = -2
    dz[1][c3] = f(0.05)
    dz[2][c3] = sd[c3] - f(0.05)
    isnow[c4] = -3
    dz[0][c4] = f(0.05)
    dz[1][c4] = f(0.5) * (sd[c4] - f(0.05))
    dz[2][c4] = f(0.5) * (sd[c4] - f(0.05))
    isnow[c5] = -3
    dz[0][c5] = f(0.05)
    dz[1][c5] = f(0.20)
    dz[2][c5] = (sd[c5] - f(0.20)) - f(0.05)
    tsno = np.zeros((nj, 3, ni), np.float32)
    snice = np.zeros((nj, 3, ni), np.float32)
    snliq = np.zeros((nj, 3, ni), np.float32)
    zsnso = np.zeros((nj, 7, ni), np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        dens = swe / snodep
    for k in range(3):  # layer index k-2
        act = (k - 2) >= (isnow + 1)
        tsno[:, k, :] = np.where(act, tg, f(0.0))
        snice[:, k, :] = np.where(act, f(1.0) * dz[k] * dens, f(0.0))
    dzsnso = np.zeros((7, nj, ni), np.float32)
    for k in range(3):
        dzsnso[k] = -dz[k]
    dzsnso[3] = zsoil[0]
    for k in range(1, 4):
        dzsnso[3 + k] = zsoil[k] - zsoil[k - 1]
    # ZSNSO(ISNOW+1) = DZSNSO(ISNOW+1); ZSNSO(IZ) = ZSNSO(IZ-1) + DZSNSO(IZ)
    acc = np.zeros((nj, ni), np.float32)
    for k in range(7):
        lay = k - 2
        act = lay >= (isnow + 1)
        first = lay == (isnow + 1)
        acc = np.where(first, dzsnso[k], np.where(act, acc + dzsnso[k], acc)).astype(np.float32)
        zsnso[:, k, :] = np.where(act, acc, f(0.0))
    return isnow, tsno, snice, snliq, zsnso


def raw_initial_fields(cfg, st, frc1):
    """What a cold-start input file provides before NOAHMP_INIT runs: TSK, TSLB, SMOIS, SNOW (mm), SNOWH (m)."""
    f = np.float32
    g = st["_g"]
    nj, ni = g.shape
    xp = backend()
    tsk = frc1["t"].astype(np.float32)
    zsoil = -np.cumsum(DZS).astype(np.float32)
    R = {"tsk": tsk, "tslb": np.zeros((nj, 4, ni), f), "smois": np.zeros((nj, 4, ni), f)}
    sm = (0.15 + 0.20 * uniform(xp, g, -1, F_SMOIS)).astype(np.float32)
    for k in range(4):
        w = f(np.exp(zsoil[k] / 2.0))  # relax from skin temperature towards TMN with depth
        R["tslb"][:, k, :] = (st["tmn"] + (tsk - st["tmn"]) * w).astype(np.float32)
        R["smois"][:, k, :] = sm
    snodep = np.where(uniform(xp, g, -1, F_SNOW) < cfg.snow_frac, 0.5 + uniform(xp, g, -1, F_SNODEP), 0.0)
    R["snowh"] = snodep.astype(np.float32)
    R["snow"] = (f(250.0) * R["snowh"]).astype(np.float32)  # SNOW = 250*SNODEP mm
    return R


def cold_start(cfg, st, frc1, tables):
    """Initial state arrays (numpy, Fortran layout as (nj[,k],ni)) for a tile: restated NOAHMP_INIT.

    st = static_fields(numpy backend), frc1 = forcing at step 1. Returns dict of all INOUT/OUT arrays.
    """
    f = np.float32
    g = st["_g"]
    nj, ni = g.shape
    A = {}
    for n in _capi.INOUT_NAMES + _capi.OUT_NAMES:
        A[n] = np.zeros(_capi.array_shape(n, ni, nj), _capi.array_dtype(n))
    veg, soil = st["ivgtyp"], st["isltyp"]
    glac = (veg == ISICE) & (st["xice"] <= 0.0)
    smcmax = tables["maxsmc"][soil - 1]
    bb = tables["bb"][soil - 1]
    psisat = tables["satpsi"][soil - 1]
    zsoil = -np.cumsum(DZS).astype(np.float32)
    raw = raw_initial_fields(cfg, st, frc1)
    tsk = raw["tsk"]
    A["tslb"][...] = raw["tslb"]
    A["smois"][...] = raw["smois"]
    snodep, swe = raw["snowh"], raw["snow"]
    # --- NOAHMP_INIT :1032-1069
    for k in range(4):
        smk, tk = A["smois"][:, k, :], A["tslb"][:, k, :]
        smk = np.where(smk > smcmax, smcmax, smk)
        ok = (bb > 0.0) & (smcmax > 0.0) & (psisat > 0.0)
        with np.errstate(all="ignore"):
            fk = (((f(3.335E5) / (f(9.81) * (-psisat))) * ((tk - f(273.15)) / tk)) ** (f(-1.0) / bb)) * smcmax
        fk = np.maximum(fk.astype(np.float32), f(0.02))
        sh = np.where(ok & (tk < f(273.149)), np.minimum(fk, smk), smk)
        A["smois"][:, k, :] = np.where(glac, f(1.0), smk)
        A["sh2o"][:, k, :] = np.where(glac, f(0.0), sh)
        A["tslb"][:, k, :] = np.where(glac, np.minimum(tk, f(263.15)), tk)
    swe = np.where(glac, np.maximum(swe, f(10.0)), swe).astype(np.float32)
    snodep = np.where(glac, swe * f(0.01), snodep).astype(np.float32)
    A["snow"][...] = swe
    A["snowh"][...] = snodep
    A["tsk"][...] = tsk
    # --- :1073-1120
    cold = (swe > 0.0) & (tsk > f(273.15))
    tfix = np.where(cold, f(273.15), tsk)
    A["tvxy"][...] = tfix
    A["tgxy"][...] = tfix
    A["eahxy"][...] = 2000.0
    A["tahxy"][...] = tfix
    A["t2mvxy"][...] = tfix
    A["t2mbxy"][...] = tfix
    A["alboldxy"][...] = 0.65
    if cfg.opts["iopt_run"] != 5:
        A["waxy"][...] = 4900.0
        A["wtxy"][...] = 4900.0
        A["zwtxy"][...] = (f(25.0) + f(2.0)) - f(4900.0) / f(1000.0) / f(0.2)
    A["lfmassxy"][...] = 50.0
    A["stmassxy"][...] = 50.0
    A["rtmassxy"][...] = 500.0
    A["woodxy"][...] = 500.0
    A["stblcpxy"][...] = 1000.0
    A["fastcpxy"][...] = 1000.0
    A["xsaixy"][...] = 0.1
    A["xlaixy"][...] = 1.0  # the driver takes LAI from the forcing file when present
    isnow, tsno, snice, snliq, zsnso = _snow_init(swe, snodep, A["tgxy"], zsoil)
    A["isnowxy"][...] = isnow
    A["tsnoxy"][...] = tsno
    A["snicexy"][...] = snice
    A["snliqxy"][...] = snliq
    A["zsnsoxy"][...] = zsnso
    A["taussxy"][...] = 0.0
    # --- driver first-step guesses (module_hrldas_noahmp_driver.F90:374-384)
    A["eahxy"][...] = (frc1["p"] * frc1["qv"]) / (f(0.622) + frc1["qv"])
    A["tahxy"][...] = frc1["t"]
    A["chxy"][...] = 0.1
    A["cmxy"][...] = 0.1
    A["albedo"][...] = 0.2
    A["emiss"][...] = 0.95
    A["qsfc"][...] = frc1["qv"] / (f(1.0) + frc1["qv"])
    return A


def cold_start_device(model, cfg, st, frc1, xs=1, ys=1):
    """The same initial state through the library's own cold start: raw_initial_fields -> NoahMP.init (NOAHMP_INIT on
    the GPU) -> the driver's first-step guesses.  iopt_run != 5 (groundwater_fields supplies that state)."""
    f = np.float32
    g = st["_g"]
    nj, ni = g.shape
    A = {}
    for n in _capi.INOUT_NAMES + _capi.OUT_NAMES:
        A[n] = np.zeros(_capi.array_shape(n, ni, nj), _capi.array_dtype(n))
    raw = raw_initial_fields(cfg, st, frc1)
    for n in ("tsk", "tslb", "smois", "snow", "snowh"):
        A[n][...] = raw[n]
    I = {n: A[n] for n, k in _capi.INIT_SPEC if k in ("pf", "pi") and n in A}
    I.update(isltyp=st["isltyp"], ivgtyp=st["ivgtyp"], xice=st["xice"], tmn=st["tmn"], dzs=DZS,
             chstarxy=np.zeros((nj, ni), f))
    sc = dict(isurban=ISURBAN, isice=ISICE, iswater=ISWATER, fndsoilw=0, fndsnowh=1, nsoil=4, restart=0, allowed_to_read=1,
              iopt_run=cfg.opts["iopt_run"], dx=1000.0, dy=1000.0, wtddt=30.0, dt=float(cfg.dt),
              ids=xs, ide=xs + ni, jds=ys, jde=ys + nj, kds=1, kde=2, ims=xs, ime=xs + ni - 1, jms=ys, jme=ys + nj - 1,
              kms=1, kme=2, its=xs, ite=xs + ni - 1, jts=ys, jte=ys + nj - 1, kts=1, kte=2)
    model.init(I, sc)
    A["xlaixy"][...] = 1.0  # the driver takes LAI from the forcing file when present
    # driver first-step guesses (module_hrldas_noahmp_driver.F90:374-384)
    A["eahxy"][...] = (frc1["p"] * frc1["qv"]) / (f(0.622) + frc1["qv"])
    A["tahxy"][...] = frc1["t"]
    A["chxy"][...] = 0.1
    A["cmxy"][...] = 0.1
    A["albedo"][...] = 0.2
    A["emiss"][...] = 0.95
    A["qsfc"][...] = frc1["qv"] / (f(1.0) + frc1["qv"])
    return A


def args_from(cfg, st, frc, state, itimestep, nk=2):
    """Assemble (arrays, scalars) for _capi.make_args from static fields, one step's forcing and state."""
    nj, ni = st["ivgtyp"].shape
    arr = dict(state)
    for n in ("ivgtyp", "isltyp", "xland", "xice", "tmn", "xlatin", "vegfra", "vegmax"):
        arr[n] = np.ascontiguousarray(st[n])
    arr["coszin"] = np.ascontiguousarray(frc["coszin"], np.float32)
    arr["swdown"] = np.ascontiguousarray(frc["swdown"], np.float32)
    arr["glw"] = np.ascontiguousarray(frc["glw"], np.float32)
    arr["rainbl"] = np.ascontiguousarray(frc["rainbl"], np.float32)

    def lev2(x):  # levels 1 and 2 identical (driver :338-344)
        out = np.empty((nj, nk, ni), np.float32)
        out[:, :, :] = x[:, None, :]
        return out

    arr["t3d"], arr["qv3d"], arr["u_phy"], arr["v_phy"], arr["p8w3d"] = (
        lev2(frc["t"]), lev2(frc["qv"]), lev2(frc["u"]), lev2(frc["v"]), lev2(frc["p"]))
    arr["dz8w"] = np.full((nj, nk, ni), 60.0, np.float32)  # 2*zlvl, zlvl=30 (driver :344)
    arr["dzs"] = DZS
    sc = dict(cfg.opts)
    sc.update(itimestep=itimestep, yr=int(frc["yr"]), julian=float(frc["julian"]), dt=float(cfg.dt), nsoil=4,
              dx=1000.0, xice_thres=0.5, isice=ISICE, isurban=ISURBAN,
              ids=1, ide=ni, jds=1, jde=nj, kds=1, kde=nk, ims=1, ime=ni, jms=1, jme=nj, kms=1, kme=nk,
              its=1, ite=ni, jts=1, jte=nj, kts=1, kte=nk)
    return arr, sc


def _smooth_topo(cfg, g):
    """Gently rolling terrain (slopes of a few m per km) as a function of the GLOBAL cell index: the lateral-flow
    stencil needs a topography with physical gradients, unlike the white-noise HGT used for surface pressure."""
    gi = (g % cfg.ni).astype(np.float64)
    gj = (g // cfg.ni).astype(np.float64)
    z = 400.0 + 60.0 * np.sin(2 * np.pi * gi / 97.0) * np.cos(2 * np.pi * gj / 131.0) + 25.0 * np.sin(2 * np.pi * (gi + 2 * gj) / 41.0)
    return z.astype(np.float32)


def groundwater_fields(cfg, st, state, dx=1000.0):
    """Synthetic inputs of the opt_run=5 scheme for a tile, and MMF-consistent starting values
    of the state GROUNDWATER_INIT (noahmpdrv.F90:1286-1470) would otherwise provide.  Returns (wt_arrays, scalars);
    wt_arrays aliases state[...] for SMOIS, SH2O, SMCWTD, WTD (= ZWTXY), DEEPRECH, RECH so one dict serves both
    noahmplsm and WTABLE."""
    f = np.float32
    g = st["_g"]
    xp = backend()
    u = lambda fld: uniform(xp, g, -1, fld)
    nj, ni = g.shape
    eq = (-20.0 + 19.0 * u(F_EQWTD)).astype(f)
    A = {
        "fdepth": (50.0 + 450.0 * u(F_FDEPTH)).astype(f), "area": np.full((nj, ni), dx * dx, f), "topo": _smooth_topo(cfg, g),
        "rivercond": (1.0e-3 * u(F_RCOND)).astype(f), "riverbed": (eq - f(1.0)).astype(f), "eqwtd": eq,
        "pexp": np.ones((nj, ni), f),
        "qrf": np.zeros((nj, ni), f), "qspring": np.zeros((nj, ni), f), "qslat": np.zeros((nj, ni), f),
        "qrfs": np.zeros((nj, ni), f), "qsprings": np.zeros((nj, ni), f),
        "xland": st["xland"], "xice": st["xice"], "isltyp": st["isltyp"], "ivgtyp": st["ivgtyp"], "dzs": DZS,
    }
    state["zwtxy"][...] = (eq + (-3.0 + 6.0 * u(F_WTD0))).astype(f).clip(-40.0, -0.05)
    state["smcwtdxy"][...] = (0.15 + 0.2 * u(F_SMCWTD)).astype(f)
    state["smoiseq"][...] = (f(0.8) * state["smois"]).astype(f)
    state["waxy"][...] = 0.0
    for n, src in (("smois", "smois"), ("sh2oxy", "sh2o"), ("smcwtd", "smcwtdxy"), ("wtd", "zwtxy"),
                   ("deeprech", "deeprechxy"), ("rech", "rechxy"), ("smoiseq", "smoiseq")):
        A[n] = state[src]
    sc = dict(nsoil=4, xice_threshold=0.5, isice=ISICE, isurban=ISURBAN, wtddt=30.0,
              ids=1, ide=cfg.ni, jds=1, jde=cfg.nj, kds=1, kde=2, ims=1, ime=ni, jms=1, jme=nj, kms=1, kme=2,
              its=1, ite=ni, jts=1, jte=nj, kts=1, kte=2)
    return A, sc
