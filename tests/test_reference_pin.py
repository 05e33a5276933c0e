"""The pin of the oracle: the reference's own Fortran text, machine-translated to C++ (oracle/ref/f90cxx.py — a
translator of the language, it knows no physics) and compiled into oracle/_ref/libnoahmp_ref.so, against the
hand-written oracle, bit for bit, through the whole `noahmplsm` call (dispatcher, REDPRM, NOAHMP_SFLX, NOAHMP_GLACIER),
NOAHMP_INIT (with GROUNDWATER_INIT) and WTABLE_mmf_noahmp.

Two layers:
  * where the translated library exists (this container builds it from /root/reference; it travels to the GPU box as a
    built .so): oracle == translated reference on C1..C4 populations, snow / glacier / sea-ice / water cells, every
    accepted opt_* value, both math back ends (host libm and the portable nmp_math.h);
  * everywhere: the committed golden vectors the translated reference produced (tests/golden/reference_vectors.npz,
    made by tests/golden/gen_reference_vectors.py) are reproduced by the oracle — and by the CUDA PARITY build in
    tests/test_parity_gpu.py::test_reference_golden_vectors — without the reference in the loop.
"""
import importlib.util
import os

import numpy as np
import pytest

from noahmp_b200 import _capi, synthetic as S

from helpers import clone_state, diff_report, make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_reference_vectors",
                                                  os.path.join(ROOT, "tests", "golden", "gen_reference_vectors.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="module")
def R(built, tables_usgs_struct):
    from oracle.ref import refmodel
    so = refmodel.build()
    if so is None:
        pytest.skip("oracle/_ref/libnoahmp_ref.so is not built and there is no reference tree to build it from")
    r = refmodel.RefModel(so)
    missing, done = r.set_tables(tables_usgs_struct)
    # every table the physics reads has a module variable of the same name in the reference
    assert missing == ["nveg", "isurban_mp"], missing
    return r


def _both(O, R, cfg, tables, ts, nsteps, mode, prepare=None, skip=(), prepare_static=None):
    """oracle and translated reference side by side, each advancing its own state; -> final oracle state"""
    xp, st, state0 = make_case(cfg, tables)
    if prepare_static:
        prepare_static(st)
        state0 = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), tables)
    if prepare:
        prepare(state0)
    sa, sb = clone_state(state0), clone_state(state0)
    O.set_math_mode(mode)
    R.set_math_mode(mode)
    names = [n for n in _capi.INOUT_NAMES + _capi.OUT_NAMES if n not in skip]
    for step in range(1, nsteps + 1):
        frc = S.forcing(xp, cfg, step, st)
        arr, sc = S.args_from(cfg, st, frc, sa, step)
        status, _ = O.noahmplsm(arr, sc, ts, nthreads=4)
        assert status.code == 0, (step, status.code, status.i, status.j, status.value)
        arr2, sc2 = S.args_from(cfg, st, frc, sb, step)
        R.noahmplsm(arr2, sc2)
        rep = diff_report(sa, sb, names)
        assert not rep, "step %d: %r" % (step, rep)
    return st, sa


def _cfg(name, ni, nj, **opts):
    cfg = S.named_config(name)
    cfg.ni, cfg.nj = ni, nj
    cfg.opts.update(opts)
    return cfg


def test_the_translation_has_the_reference_argument_list(R):
    """`noahmplsm` as translated takes exactly the members of noahmp_lsm_args, in order (the boundary, once more),
    and the translated files hold the whole path."""
    sig = R.signature("NOAHMPLSM")
    assert [n for n, _ in sig] == [n.upper() for n, _ in _capi.NoahmpLsmArgs._fields_]
    for name in ("NOAHMP_SFLX", "NOAHMP_GLACIER", "REDPRM", "ENERGY", "WATER", "CARBON", "VEGE_FLUX", "BARE_FLUX",
                 "SFCDIF1", "SFCDIF2", "STOMATA", "CANRES", "TSNOSOI", "PHASECHANGE", "SNOWWATER", "SOILWATER",
                 "GROUNDWATER", "SHALLOWWATERTABLE", "CO2FLUX", "ENERGY_GLACIER", "WATER_GLACIER"):
        assert R.signature(name), name


@pytest.mark.parametrize("mode", [0, 1], ids=["libm", "portable"])
def test_c1_24_steps(O, R, tables_usgs, tables_usgs_struct, mode):
    """BASELINE config 0 (the reference's own CPU-runnable case): a diurnal cycle, every word of every array."""
    _both(O, R, _cfg("C1", 10, 10), tables_usgs, tables_usgs_struct, 24, mode)


def test_c2_nldas_tile(O, R, tables_usgs, tables_usgs_struct):
    """BASELINE config 1 population (default options, 10 % water) on a 116 x 112 tile, 24 steps."""
    st, s = _both(O, R, _cfg("C2", 116, 112), tables_usgs, tables_usgs_struct, 24, 0)
    assert (st["xland"] > 1.5).mean() > 0.03


@pytest.mark.parametrize("mode", [0, 1], ids=["libm", "portable"])
def test_c3_dynamic_vegetation_and_snow(O, R, tables_usgs, tables_usgs_struct, mode):
    """The headline physics (dveg = 2 carbon pools, 3-layer snow) on 96 x 64 columns, 18 steps; all snow-layer counts
    occur."""
    st, s = _both(O, R, _cfg("C3", 96, 64), tables_usgs, tables_usgs_struct, 18, mode)
    assert set(np.unique(s["isnowxy"])) >= {-3, -2, -1, 0}


def test_c4_glacier_seaice_water(O, R, tables_usgs, tables_usgs_struct):
    """Land + land-ice (NOAHMP_GLACIER) + water; 120 x 90 columns, 12 steps."""
    cfg = _cfg("C4", 120, 90)
    cfg.glacier_frac, cfg.snow_frac, cfg.t_base = 0.2, 0.4, 268.0
    st, s = _both(O, R, cfg, tables_usgs, tables_usgs_struct, 12, 0)
    assert (st["ivgtyp"] == S.ISICE).sum() > 200 and (st["xland"] > 1.5).sum() > 1000


OPTS = [
    dict(idveg=1, iopt_crs=2, iopt_btr=2, iopt_run=2, iopt_sfc=2, iopt_frz=2, iopt_inf=2, iopt_rad=1, iopt_alb=1,
         iopt_snf=2, iopt_tbot=1, iopt_stc=2),
    dict(idveg=3, iopt_crs=1, iopt_btr=3, iopt_run=3, iopt_sfc=1, iopt_frz=1, iopt_inf=1, iopt_rad=2, iopt_alb=2,
         iopt_snf=3, iopt_tbot=2, iopt_stc=1),
    dict(idveg=5, iopt_crs=2, iopt_btr=1, iopt_run=4, iopt_sfc=2, iopt_frz=2, iopt_inf=1, iopt_rad=3, iopt_alb=1,
         iopt_snf=1, iopt_tbot=2, iopt_stc=2),
    dict(idveg=2, iopt_crs=1, iopt_btr=1, iopt_run=5, iopt_sfc=1, iopt_frz=1, iopt_inf=2, iopt_rad=3, iopt_alb=2,
         iopt_snf=1, iopt_tbot=2, iopt_stc=1),
    dict(idveg=4, iopt_crs=1, iopt_btr=2, iopt_run=1, iopt_sfc=1, iopt_frz=2, iopt_inf=2, iopt_rad=1, iopt_alb=1,
         iopt_snf=2, iopt_tbot=1, iopt_stc=1),
]


@pytest.mark.parametrize("k", range(len(OPTS)))
def test_every_accepted_option_value(O, R, tables_usgs, tables_usgs_struct, k):
    """Each value of each opt_* switch the reference accepts offline appears in one of the combinations."""
    opts = OPTS[k]
    cfg = _cfg("C3", 64, 48, **opts)
    cfg.glacier_frac = 0.1

    def prepare(state0):
        if opts["iopt_run"] == 5:  # state of the MMF scheme (NOAHMP_INIT's groundwater block)
            state0["smoiseq"][...] = 0.8 * state0["smois"]
            state0["zwtxy"][...] = -3.0
            state0["smcwtdxy"][...] = 0.3

    _both(O, R, cfg, tables_usgs, tables_usgs_struct, 8, k % 2, prepare)


# other climates than the configurations' own: (name, base, t_base, start, latitudes, snow_frac, glacier_frac, options)
CLIMATES = [
    ("tropical_july", "C2", 301.0, (2017, 7, 15, 0), (-10.0, 15.0), 0.0, 0.0, dict(iopt_snf=2)),
    ("polar_winter", "C3", 243.0, (2017, 1, 10, 0), (60.0, 80.0), 0.9, 0.3, {}),
    ("spring_melt", "C3", 272.5, (2017, 4, 10, 6), (40.0, 55.0), 0.8, 0.1, dict(iopt_alb=1, iopt_snf=3)),
    ("summer_noon_dynveg", "C3", 295.0, (2017, 7, 1, 12), (30.0, 45.0), 0.0, 0.0, dict(iopt_rad=1, iopt_btr=2)),
    ("autumn_frost_koren", "C2", 271.0, (2017, 10, 20, 18), (45.0, 60.0), 0.2, 0.05,
     dict(iopt_frz=2, iopt_inf=2, iopt_stc=2, iopt_crs=2, iopt_run=3)),
]


@pytest.mark.parametrize("case", CLIMATES, ids=[c[0] for c in CLIMATES])
def test_other_climates(O, R, tables_usgs, tables_usgs_struct, case):
    """Branches the configurations' own winter / spring forcing does not reach: hot and humid, polar night, melting
    snowpack, strong insolation on growing vegetation, freezing soil with the Koren99 options.  16 steps each."""
    name, base, t_base, start, lat, snow, glac, opts = case
    cfg = _cfg(base, 72, 48, **opts)
    cfg.t_base, cfg.start, cfg.lat, cfg.snow_frac, cfg.glacier_frac = t_base, start, lat, snow, glac
    _both(O, R, cfg, tables_usgs, tables_usgs_struct, 16, len(name) % 2)


def test_long_melt_season(O, R, tables_usgs, tables_usgs_struct):
    """Eight days of a melting snowpack: layers thin out, combine, vanish (COMBINE / DIVIDE / SNOWH2O's rarer branches,
    PHASECHANGE's sign reversals), on vegetated and glacier columns."""
    cfg = _cfg("C3", 48, 32, iopt_snf=3)
    cfg.t_base, cfg.start, cfg.lat, cfg.snow_frac, cfg.glacier_frac = 278.0, (2017, 4, 20, 0), (38.0, 50.0), 0.9, 0.15
    def prepare(state0):   # a nearly empty top layer on some three-layer packs: COMBINE merges it downwards
        top = (state0["isnowxy"] == -3)
        top[:, ::2] = False
        removed = state0["snicexy"][:, 0, :][top] + state0["snliqxy"][:, 0, :][top] - np.float32(0.05)
        state0["snow"][top] = state0["snow"][top] - removed      # SNEQV stays the sum of the layers
        state0["sneqvoxy"][top] = state0["snow"][top]
        state0["snicexy"][:, 0, :][top] = 0.05
        state0["snliqxy"][:, 0, :][top] = 0.0
        one = (state0["isnowxy"] == -1)                          # a single, nearly empty layer: the pack vanishes
        one[:, 1::2] = False
        removed = state0["snicexy"][:, 2, :][one] + state0["snliqxy"][:, 2, :][one] - np.float32(0.05)
        state0["snow"][one] = state0["snow"][one] - removed
        state0["sneqvoxy"][one] = state0["snow"][one]
        state0["snicexy"][:, 2, :][one] = 0.05
        state0["snliqxy"][:, 2, :][one] = 0.0
        deep = (state0["isnowxy"] == -3)                         # more than 2000 mm of snow: the excess flows off
        deep[:, 1::3] = False
        deep[:, 2::3] = False
        deep &= ~top
        state0["snicexy"][:, 2, :][deep] += 2100.0
        state0["snow"][deep] += 2100.0
        state0["sneqvoxy"][deep] = state0["snow"][deep]

    st, s = _both(O, R, cfg, tables_usgs, tables_usgs_struct, 192, 0, prepare)
    assert (s["isnowxy"] == 0).mean() > 0.15 and set(np.unique(s["isnowxy"])) >= {-3, -2, -1, 0}


def test_glacier_in_summer(O, R, tables_usgs, tables_usgs_struct):
    """Land-ice columns whose upper layers cross the freezing point: the inter-layer heat / melt redistribution of
    PHASECHANGE_GLACIER.  Leap year (YEARLEN = 366)."""
    cfg = _cfg("C4", 40, 30)
    cfg.t_base, cfg.start, cfg.lat, cfg.snow_frac, cfg.glacier_frac, cfg.water_frac = 276.5, (2016, 7, 5, 0), (60.0, 72.0), 0.3, 0.6, 0.1
    def prepare(state0):   # thawed ice layers above and below frozen ones: the four redistribution sweeps
        state0["tslb"][::2, 0, :] = 274.2
        state0["tslb"][::2, 2, :] = 273.6
        state0["tslb"][1::2, 1, :] = 273.9

    _both(O, R, cfg, tables_usgs, tables_usgs_struct, 96, 1, prepare)


def test_sea_ice_points_soil_type_14_and_dry_soil(O, R, tables_usgs, tables_usgs_struct):
    """The dispatcher's special cases: sea-ice points (XICE >= XICE_THRES: skipped, SH2O = 1, LAI = 0.01, and the
    ITIMESTEP = 1 fills), water-type soil at a land point (reset to 7), a bone-dry top layer (RSURF's guard); year 2000
    (the 400-year leap rule)."""
    cfg = _cfg("C2", 40, 30)
    cfg.start = (2000, 2, 28, 12)

    def prepare_static(st):
        st["xice"][5:9, 3:30] = 1.0
        st["isltyp"][12:15, 4:20] = 14

    def prepare(state0):
        state0["smois"][20:24, :, :] = 0.02
        state0["sh2o"][20:24, :, :] = 0.004

    st, s = _both(O, R, cfg, tables_usgs, tables_usgs_struct, 30, 0, prepare, prepare_static=prepare_static)
    land_ice = (st["xice"] >= 0.5) & (st["xland"] < 1.5)
    assert land_ice.sum() > 50 and (s["xlaixy"][land_ice] == np.float32(0.01)).all()


def test_option_values_are_all_covered():
    seen = {}
    for o in OPTS:
        for n, v in o.items():
            seen.setdefault(n, set()).add(v)
    want = dict(idveg={1, 2, 3, 4, 5}, iopt_crs={1, 2}, iopt_btr={1, 2, 3}, iopt_run={1, 2, 3, 4, 5}, iopt_sfc={1, 2},
                iopt_frz={1, 2}, iopt_inf={1, 2}, iopt_rad={1, 2, 3}, iopt_alb={1, 2}, iopt_snf={1, 2, 3},
                iopt_tbot={1, 2}, iopt_stc={1, 2})
    assert seen == want


def test_rejected_inputs_stop_both(O, R, tables_usgs, tables_usgs_struct):
    """A soil type beyond the table: the reference calls wrf_error_fatal in REDPRM, the oracle reports the column."""
    cfg = _cfg("C1", 10, 10)
    xp, st, state = make_case(cfg, tables_usgs)
    st["isltyp"][3, 4] = 25
    arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, 1, st), state, 1)
    O.set_math_mode(0)
    status, _ = O.noahmplsm(arr, sc, tables_usgs_struct, nthreads=1)
    assert status.code != 0 and (status.i, status.j) == (5, 4)
    arr2, sc2 = S.args_from(cfg, st, S.forcing(xp, cfg, 1, st), clone_state(state), 1)
    with pytest.raises(RuntimeError, match="too many input soil types"):
        R.noahmplsm(arr2, sc2)
    # a snow pack whose layers do not add up to SNEQV: the model's own water-budget check (ERROR) stops both
    cfg = _cfg("C3", 12, 8)
    cfg.snow_frac = 1.0
    xp, st, state = make_case(cfg, tables_usgs)
    j, i = np.argwhere((state["isnowxy"] == -3) & (st["xland"] < 1.5) & (st["ivgtyp"] != S.ISICE))[0]
    state["snow"][j, i] += 5.0
    other = clone_state(state)
    arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, 1, st), state, 1)
    status, _ = O.noahmplsm(arr, sc, tables_usgs_struct, nthreads=1)
    assert status.code != 0 and (status.i, status.j) == (i + 1, j + 1)
    arr2, sc2 = S.args_from(cfg, st, S.forcing(xp, cfg, 1, st), other, 1)
    with pytest.raises(RuntimeError, match="Water budget problem"):
        R.noahmplsm(arr2, sc2)


def _same(x, y):
    return ((x == y) | (np.isnan(x) & np.isnan(y))).all() if x.dtype.kind == "f" else (x == y).all()


@pytest.mark.parametrize("name,ni,nj,run", [("C4", 96, 64, 1), ("C3", 64, 48, 1), ("C2", 60, 44, 5), ("C3", 130, 70, 5)])
def test_noahmp_init(O, R, tables_usgs_struct, name, ni, nj, run):
    """NOAHMP_INIT (SNOW_INIT; with iopt_run = 5 also GROUNDWATER_INIT, EQSMOISTURE and LATERALFLOW) on raw cold-start
    fields: every array it writes.  The groundwater cases contain columns whose Newton iteration for the deep soil
    moisture diverges to NaN, which the reference's MAX(SMC, 1.E-4) turns into 1.E-4 (gfortran's MAX)."""
    import test_init as TI
    for mode in (0, 1):
        O.set_math_mode(mode)
        R.set_math_mode(mode)
        A, sc = TI.init_case(name, ni, nj, run)[3:]
        A["snowh"][::3, ::2], A["snow"][::3, ::2] = 0.03, 6.0       # one snow layer
        A["snowh"][1::3, 1::2], A["snow"][1::3, 1::2] = 0.10, 25.0  # two
        A["snowh"][2::3, ::5], A["snow"][2::3, ::5] = 0.012, 2.0    # too thin for a layer
        A["snowh"][2::3, 1::5], A["snow"][2::3, 1::5] = 0.60, 150.0 # three
        A["snowh"][2::3, 3::5], A["snow"][2::3, 3::5] = 0.35, 80.0  # three, the thinner split
        B = TI.clone(A)
        rc, step_o = O.init(A, sc, tables_usgs_struct)
        step_r = R.init(B, sc)
        assert rc == 0 and step_o == step_r
        bad = [n for n in A if isinstance(A[n], np.ndarray) and not _same(A[n], B[n])]
        assert not bad, (mode, bad)


@pytest.mark.parametrize("name,ni,nj", [("C4", 40, 30), ("C3", 96, 64), ("C2", 80, 60)])
def test_wtable_coupled_with_the_column_physics(O, R, tables_usgs, tables_usgs_struct, name, ni, nj):
    """opt_run = 5: six steps of noahmplsm followed by WTABLE_mmf_noahmp (LATERALFLOW, UPDATEWTD), the oracle and the
    translated reference each on its own state."""
    for mode in (0, 1):
        O.set_math_mode(mode)
        R.set_math_mode(mode)
        cfg = _cfg(name, ni, nj, iopt_run=5)
        xp, st, sa = make_case(cfg, tables_usgs)
        sb = clone_state(sa)
        wa, wsc = S.groundwater_fields(cfg, st, sa)
        wb, _ = S.groundwater_fields(cfg, st, sb)
        # water tables from inside the top soil layer to far below the resolved soil, next to each other, so that
        # the lateral flow moves them both ways through the layers: the branches of UPDATEWTD
        jj, ii = np.meshgrid(np.arange(nj), np.arange(ni), indexing="ij")
        for s_ in (sa, sb):
            s_["zwtxy"][...] = (-0.05 - 9.0 * (((jj * 7 + ii * 3) % 23) / 22.0) ** 2).astype(np.float32)
            s_["smcwtdxy"][...] = 0.3
        for w_ in (wa, wb):
            w_["fdepth"][...] = 400.0
        for step in range(1, 7):
            frc = S.forcing(xp, cfg, step, st)
            arr, sc = S.args_from(cfg, st, frc, sa, step)
            O.noahmplsm(arr, sc, tables_usgs_struct, nthreads=4)
            arr2, sc2 = S.args_from(cfg, st, frc, sb, step)
            R.noahmplsm(arr2, sc2)
            O.wtable(wa, wsc, tables_usgs_struct)
            R.wtable(wb, wsc)
            bad = [n for n in wa if isinstance(wa[n], np.ndarray) and not _same(wa[n], wb[n])]
            bad += ["state." + n for n in sa if not _same(sa[n], sb[n])]
            assert not bad, (mode, step, bad)
        assert np.abs(wa["qslat"]).max() > 0


def test_wtable_rising_and_falling_through_the_layers(O, R, tables_usgs, tables_usgs_struct):
    """WTABLE_mmf_noahmp alone on flat terrain without rivers, water tables in a checkerboard of neighbouring depths at
    every level of the column (inside each soil layer, just below the soil, deep): half of the columns gain water and
    fill their layers upwards, the others drain — UPDATEWTD's rising branches, which the coupled runs reach rarely."""
    for mode in (0, 1):
        O.set_math_mode(mode)
        R.set_math_mode(mode)
        cfg = _cfg("C2", 48, 40, iopt_run=5)
        cfg.water_frac = 0.0
        xp, st, sa = make_case(cfg, tables_usgs)
        sb = clone_state(sa)
        wa, wsc = S.groundwater_fields(cfg, st, sa)
        wb, _ = S.groundwater_fields(cfg, st, sb)
        jj, ii = np.meshgrid(np.arange(cfg.nj), np.arange(cfg.ni), indexing="ij")
        level = np.array([-0.05, -0.25, -0.7, -1.5, -2.2, -3.5, -8.0], np.float32)[(jj // 6) % 7]
        wtd = (level - np.float32(0.6) * ((ii + jj) % 2) * np.minimum(np.float32(1.0), -level)).astype(np.float32)
        for s_, w_ in ((sa, wa), (sb, wb)):
            s_["zwtxy"][...] = wtd
            s_["smcwtdxy"][...] = 0.25
            s_["smois"][...] = (0.5 * s_["smoiseq"] / 0.8 + 0.05).astype(np.float32)   # room to fill
            # ... except in the eastern half, a hair below saturation: what arrives spills into the layer above
            smcmax = np.asarray(tables_usgs_struct.maxsmc, np.float32)[st["isltyp"] - 1]
            s_["smois"][:, :, cfg.ni // 2:] = (smcmax - np.float32(2e-4))[:, None, cfg.ni // 2:]
            s_["smcwtdxy"][:, cfg.ni // 2:] = (smcmax - np.float32(2e-4))[:, cfg.ni // 2:]
            s_["sh2o"][...] = s_["smois"]
            w_["topo"][...] = 100.0
            w_["fdepth"][...] = 300.0
            w_["rivercond"][...] = 0.0
        for step in range(1, 13):
            O.wtable(wa, wsc, tables_usgs_struct)
            R.wtable(wb, wsc)
            bad = [n for n in wa if isinstance(wa[n], np.ndarray) and not _same(wa[n], wb[n])]
            assert not bad, (mode, step, bad)
        assert (sa["zwtxy"] > wtd).sum() > 100 and (sa["zwtxy"] < wtd).sum() > 100


# ---- leaf routines over input ranges the synthetic forcing never visits ------------------------------------------------

def test_leaf_routines_over_wide_ranges(O, R, tables_usgs_struct):
    """ESAT from -90 to +70 C, TDFCND and FRH2O (Koren99 Newton iteration) over the soil classes and every wetness,
    ROSR12 on random diagonally dominant systems of 1..7 unknowns, COMBO on random layer pairs, STOMATA over light /
    temperature / humidity: the reference's routine, called by name, against the oracle's probe of the same routine."""
    import ctypes as C
    L = O.lib()
    rng = np.random.default_rng(7)
    L.nmo_tdfcnd.argtypes = [C.c_float] * 4
    L.nmo_tdfcnd.restype = C.c_float
    L.nmo_frh2o.argtypes = [C.c_float] * 6
    L.nmo_frh2o.restype = C.c_float
    for mode in (0, 1):
        O.set_math_mode(mode)
        R.set_math_mode(mode)
        out4 = (C.c_float * 4)()
        for t in np.linspace(-90.0, 70.0, 321, dtype=np.float32):
            L.nmo_esat(C.c_float(t), out4)
            assert [np.float32(v) for v in out4] == [np.float32(v) for v in R.call("ESAT", float(t), 0.0, 0.0, 0.0, 0.0)[1:]]
        T = tables_usgs_struct
        for k in range(400):
            s = int(rng.integers(0, 12))
            smcmax, quartz, bexp, psisat = T.maxsmc[s], T.qtz[s], T.bb[s], T.satpsi[s]
            smc = np.float32(rng.uniform(0.02, 1.0) * smcmax)
            sh2o = np.float32(rng.uniform(0.0, 1.0) * smc)
            R.var("NOAHMP_GLOBALS.SMCMAX")[0] = smcmax
            R.var("NOAHMP_GLOBALS.QUARTZ")[0] = quartz
            R.var("NOAHMP_GLOBALS.BEXP")[0] = bexp
            R.var("NOAHMP_GLOBALS.PSISAT")[0] = psisat
            assert np.float32(L.nmo_tdfcnd(smc, sh2o, smcmax, quartz)) == np.float32(R.call("TDFCND", 0.0, float(smc), float(sh2o))[0])
            tk = np.float32(rng.uniform(235.0, 274.5))
            assert np.float32(L.nmo_frh2o(tk, smc, sh2o, bexp, psisat, smcmax)) == \
                np.float32(R.call("FRH2O", 0.0, float(tk), float(smc), float(sh2o))[0]), (mode, k)
        for k in range(200):
            n = int(rng.integers(1, 8))
            a, c_ = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
            b = (np.abs(a) + np.abs(c_) + rng.uniform(0.1, 2, n)).astype(np.float32)
            d = rng.uniform(-5, 5, n).astype(np.float32)
            x = np.zeros(n, np.float32)
            L.nmo_rosr12(n, a.ctypes.data, b.ctypes.data, c_.ctypes.data, d.ctypes.data, x.ctypes.data)
            top = 4 - n + 1
            P, A, B, Cc, D, DEL = (np.zeros(7, np.float32) for _ in range(6))
            A[top + 2:], B[top + 2:], Cc[top + 2:], D[top + 2:] = a, b, c_, d
            R.call("ROSR12", P, A, B, Cc, D, DEL, top, 4, 3)
            assert (P[top + 2:] == x).all(), (mode, k)
        for k in range(200):
            v = np.array([rng.uniform(0.01, 0.3), rng.uniform(0, 20), rng.uniform(0, 60), rng.uniform(240, 273.16)], np.float32)
            w = np.array([rng.uniform(0.01, 0.3), rng.uniform(0, 20), rng.uniform(0, 60), rng.uniform(240, 273.16)], np.float32)
            want = R.call("COMBO", *[float(q) for q in v], *[float(q) for q in w])[:4]
            L.nmo_combo(v.ctypes.data, w.ctypes.data)
            assert [np.float32(q) for q in want] == list(v), (mode, k)


def test_stomata_and_twostream_over_wide_ranges(O, R, tables_usgs_struct):
    """STOMATA (Ball-Berry / Farquhar, the CI iteration) over light, leaf temperature, humidity and vegetation type;
    TWOSTREAM over sun angle, leaf area, wetness and the three canopy-gap options."""
    import ctypes as C
    L = O.lib()
    rng = np.random.default_rng(11)
    L.nmo_stomata.argtypes = [C.POINTER(_capi.NoahmpTables), C.c_int, C.c_void_p, C.c_void_p]
    L.nmo_twostream.argtypes = [C.POINTER(_capi.NoahmpTables), C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_float] * 9 + [C.c_void_p]
    for mode in (0, 1):
        O.set_math_mode(mode)
        R.set_math_mode(mode)
        for k in range(300):
            veg = int(rng.choice([2, 4, 5, 7, 10, 11, 13, 14, 15, 18, 21]))
            tv = rng.uniform(255.0, 315.0)
            ei = 611.0 * np.exp(17.3 * (tv - 273.16) / (tv - 35.9))
            x = np.array([rng.uniform(0.0, 400.0), rng.uniform(0.3, 1.0), tv, ei, ei * rng.uniform(0.1, 1.0),
                          tv + rng.uniform(-5, 5), rng.uniform(60000.0, 103000.0), 0.0, 0.0, float(rng.integers(0, 2)),
                          rng.uniform(0.0, 1.0), rng.uniform(5.0, 300.0)], np.float32)
            x[7], x[8] = np.float32(0.209) * x[6], np.float32(395e-6) * x[6]
            out = np.zeros(2, np.float32)
            L.nmo_stomata(C.byref(tables_usgs_struct), veg, x.ctypes.data, out.ctypes.data)
            r = R.call("STOMATA", veg, 1.0e-6, float(x[0]), float(x[1]), 1, 1, *[float(q) for q in x[2:]], 0.0, 0.0)
            assert [np.float32(r[-2]), np.float32(r[-1])] == list(out), (mode, k, veg)
        for k in range(300):
            veg = int(rng.choice([2, 4, 5, 7, 10, 11, 13, 14, 15, 18, 21]))
            opt_rad, ib, ic = int(rng.integers(1, 4)), int(rng.integers(1, 3)), int(rng.integers(0, 2))
            cosz, vai, fwet = rng.uniform(0.001, 1.0), rng.uniform(0.05, 7.0), rng.uniform(0.0, 1.0)
            tv, agd, agi = rng.uniform(250.0, 300.0), rng.uniform(0.05, 0.8), rng.uniform(0.05, 0.8)
            rho, tau, fveg = rng.uniform(0.05, 0.5), rng.uniform(0.01, 0.4), rng.uniform(0.05, 1.0)
            out = np.zeros(5, np.float32)
            L.nmo_twostream(C.byref(tables_usgs_struct), opt_rad, ib, ic, veg, cosz, vai, fwet, tv, agd, agi, rho, tau, fveg,
                            out.ctypes.data)
            R.var("NOAHMP_GLOBALS.OPT_RAD")[0] = opt_rad
            two = lambda v: np.full(2, v, np.float32)
            FAB, FRE, FTD, FTI, FREV, FREG = (np.zeros(2, np.float32) for _ in range(6))
            r = R.call("TWOSTREAM", ib, ic, veg, cosz, vai, fwet, tv, two(agd), two(agi), two(rho), two(tau), fveg, 1, 1, 1,
                       FAB, FRE, FTD, FTI, 0.0, FREV, FREG, 0.0, 0.0)
            got = [FAB[ib - 1], FRE[ib - 1], FTD[ib - 1], FTI[ib - 1], np.float32(r[19])]
            assert got == list(out), (mode, k, opt_rad, ib, ic)


def test_snow_routines_on_random_packs(O, R):
    """SNOWWATER (SNOWFALL, COMPACT, COMBINE, DIVIDE, SNOWH2O) and the glacier PHASECHANGE on random snow packs of 0-3
    layers with thin, nearly empty, very thick and melting layers -- the layer bookkeeping branches that a model run
    visits only now and then -- called by name in the translated reference and through the oracle's probes."""
    import ctypes as C
    L = C.CDLL(O.build())   # a handle of its own: other tests declare argument types on the shared one
    rng = np.random.default_rng(23)
    f = np.float32
    zsoil = np.array([-0.1, -0.4, -1.0, -2.0], f)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for mode in (0, 1):
        O.set_math_mode(mode)
        R.set_math_mode(mode)
        for k in range(600):
            isnow = -int(rng.integers(0, 4))
            dz = np.zeros(7, f)
            snice, snliq = np.zeros(3, f), np.zeros(3, f)
            for j in range(3 + isnow, 3):
                dz[j] = f(rng.choice([0.012, 0.03, 0.06, 0.2, 0.6]) * rng.uniform(0.7, 1.3))
                snice[j] = f(dz[j] * rng.uniform(50.0, 450.0)) if rng.random() > 0.25 else f(rng.uniform(0.0, 0.12))
                snliq[j] = f(rng.uniform(0.0, 0.1) * snice[j]) if rng.random() > 0.3 else f(0.0)
            dz[3:] = [0.1, 0.3, 0.6, 1.0]
            zsnso = (-np.cumsum(dz[3 + isnow:])).astype(f)
            z7 = np.zeros(7, f)
            z7[3 + isnow:] = zsnso
            if isnow < 0:
                sneqv, snowh = f((snice + snliq).sum()), f(dz[:3].sum())
            else:
                sneqv = f(rng.choice([0.0, 0.5, 8.0, 40.0]))
                snowh = f(sneqv / 120.0)
            stc = rng.uniform(255.0, 273.16, 7).astype(f)
            sh2o, sice = rng.uniform(0.05, 0.3, 4).astype(f), rng.uniform(0.0, 0.1, 4).astype(f)
            imelt = rng.integers(0, 3, 7).astype(np.int32)
            ficeold = rng.uniform(0.3, 1.0, 3).astype(f)
            sc = np.array([3600.0, rng.uniform(255, 278), 0.0, 0.0, rng.uniform(0, 2e-5), rng.uniform(0, 2e-5),
                           rng.choice([0.0, 1e-4, 1e-3])], f)
            if rng.random() < 0.5:      # snowing
                sc[3] = f(rng.uniform(1e-5, 2e-3))
                sc[2] = f(sc[3] / rng.uniform(50.0, 150.0))
            a = [x.copy() for x in (snice, snliq, sh2o, sice, stc, z7, dz)]
            b = [x.copy() for x in (snice, snliq, sh2o, sice, stc, z7, dz)]
            io2, out4, isn = np.array([snowh, sneqv], f), np.zeros(4, f), C.c_int(isnow)
            L.nmo_snowwater(vp(imelt), vp(sc), vp(zsoil), vp(ficeold), C.byref(isn), vp(io2), *[vp(x) for x in a], vp(out4))
            r = R.call("SNOWWATER", 3, 4, imelt.copy(), float(sc[0]), zsoil.copy(), float(sc[1]), float(sc[2]), float(sc[3]),
                       float(sc[4]), float(sc[5]), float(sc[6]), ficeold.copy(), 1, 1, isnow, float(snowh), float(sneqv),
                       *b, 0.0, 0.0, 0.0, 0.0)
            assert r[14] == isn.value, (mode, k, isnow, r[14], isn.value)
            assert [f(r[15]), f(r[16])] == list(io2) and [f(x) for x in r[24:28]] == list(out4), (mode, k, isnow)
            for x, y in zip(a, b):
                assert _same(x, y), (mode, k, isnow)
        for k in range(400):    # land ice: layers straddling the freezing point
            isnow = -int(rng.integers(0, 4))
            dz = np.array([0.05, 0.1, 0.3, 0.1, 0.3, 0.6, 1.0], f)
            fact = (3600.0 / (rng.uniform(1.0e6, 2.2e6, 7) * dz)).astype(f)
            stc = (273.16 + rng.choice([-3.0, -0.4, 0.0, 0.3, 1.5], 7) + rng.uniform(-0.05, 0.05, 7)).astype(f)
            snice, snliq = np.zeros(3, f), np.zeros(3, f)
            for j in range(3 + isnow, 3):
                snice[j], snliq[j] = f(rng.uniform(0.0, 60.0)), f(rng.uniform(0.0, 4.0))
            sneqv = f((snice + snliq).sum()) if isnow < 0 else f(rng.choice([0.0, 3.0]))
            snowh = f(dz[3 + isnow:3].sum()) if isnow < 0 else f(sneqv / 100.0)
            smc, sh2o = np.ones(4, f), rng.choice([0.0, 0.02, 0.4, 1.0], 4).astype(f)   # 1.0: a layer of melt water
            a = [x.copy() for x in (stc, snice, snliq)]
            b = [x.copy() for x in (stc, snice, snliq)]
            sa, sb = [smc.copy(), sh2o.copy()], [smc.copy(), sh2o.copy()]
            se, sh, qm, po = (C.c_float(sneqv), C.c_float(snowh), C.c_float(0.0), C.c_float(0.0))
            im_a, im_b = np.zeros(7, np.int32), np.zeros(7, np.int32)
            L.nmo_phasechange_glacier(isnow, C.c_float(3600.0), vp(fact), vp(dz), vp(a[0]), vp(a[1]), vp(a[2]), C.byref(se),
                                      C.byref(sh), vp(sa[0]), vp(sa[1]), C.byref(qm), vp(im_a), C.byref(po))
            r = R.call("PHASECHANGE_GLACIER", 3, 4, isnow, 3600.0, fact.copy(), dz.copy(), b[0], b[1], b[2], float(sneqv),
                       float(snowh), sb[0], sb[1], 0.0, im_b, 0.0)
            assert [f(r[9]), f(r[10]), f(r[13]), f(r[15])] == [f(se.value), f(sh.value), f(qm.value), f(po.value)], (mode, k)
            for x, y in zip(a + sa + [im_a], b + sb + [im_b]):
                assert _same(x, y), (mode, k, isnow)


def test_calc_declin_of_the_forcing_pipeline(O, R, tables_usgs):
    """Row f2: CALC_DECLIN of the HRLDAS driver (module_hrldas_noahmp_driver.F90:813-863), extracted as it stands with the
    date-string handling replaced by integer arguments (oracle/ref/build_ref.sh) and translated, against the COSZEN /
    JULIAN of the oracle's forcing preparation: every hour of four days of the year, a latitude-longitude net."""
    import test_forcing as TF
    cfg = _cfg("C1", 10, 10)
    _, st, _ = make_case(cfg, tables_usgs)
    A, B = TF._files(cfg, st, (1, 4))
    lat = np.linspace(-75.0, 80.0, 100, dtype=np.float32).reshape(10, 10)
    lon = np.linspace(-179.0, 179.5, 100, dtype=np.float32).reshape(10, 10).T.copy()
    for mode in (0, 1):
        O.set_math_mode(mode)
        R.set_math_mode(mode)
        for iday in (0, 79, 80, 171, 300, 364):
            for hour, minute, second in ((0, 0, 0), (5, 30, 0), (12, 0, 0), (17, 45, 30), (23, 59, 59)):
                out, julian = O.forcing(A, B, lat, lon, 1.0, iday, hour, minute, second, cfg.dt)
                for (j, i) in ((0, 0), (3, 7), (9, 9), (5, 2), (8, 1)):
                    r = R.call("CALC_DECLIN", iday, hour, minute, second, float(lat[j, i]), float(lon[j, i]), 0.0, 0.0)
                    assert np.float32(r[6]) == out[0][j, i], (mode, iday, hour, j, i)
                    assert np.float32(r[7]) == np.float32(julian)


# ---- the committed vectors: no reference needed ----------------------------------------------------------------------

def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


def check_against_golden(step_fn_factory, tables):
    """run every golden case through step_fn and compare with the stored arrays; -> number of arrays compared"""
    g, gen = golden(), _gen()
    n = 0
    for name in gen.CASES:
        got = gen.run(name, tables, step_fn_factory(name))
        for key, arr in got.items():
            want = g[key]
            same = (arr == want) | (np.isnan(arr) & np.isnan(want)) if arr.dtype.kind == "f" else (arr == want)
            assert same.all(), "%s: %d of %d words differ" % (key, int((~same).sum()), same.size)
            n += 1
    assert n == len(g.files)
    return n


def test_oracle_reproduces_the_reference_vectors(O, tables_usgs, tables_usgs_struct):
    """Vectors computed by the translated reference (portable math) == the oracle, on any machine."""
    O.set_math_mode(1)

    def factory(name):
        def step(arr, sc):
            status, _ = O.noahmplsm(arr, sc, tables_usgs_struct, nthreads=2)
            assert status.code == 0
        return step

    assert check_against_golden(factory, tables_usgs) > 1000


def test_vectors_are_current(R, tables_usgs):
    """Where the translated reference can be run, the committed file is what it produces now."""
    R.set_math_mode(1)
    assert check_against_golden(lambda name: R.noahmplsm, tables_usgs) > 1000
