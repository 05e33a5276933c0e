"""Cold start: NOAHMP_INIT / SNOW_INIT / GROUNDWATER_INIT / EQSMOISTURE.

CPU: the C++ oracle against the independent numpy restatement in noahmp_b200/synthetic.py, plus properties of the
groundwater initialisation.  GPU: noahmp_b200_init against the oracle, bit for bit."""
import numpy as np
import pytest

from noahmp_b200 import _capi, synthetic as S
from oracle import oracle as O

INIT_OUT = [n for n, k in _capi.INIT_SPEC if k in ("pf", "pi") and n not in
            ("isltyp", "ivgtyp", "dzs", "tsk", "tmn", "xice", "msftx", "msfty", "fdepthxy", "ht", "riverbedxy", "eqzwt",
             "rivercondxy", "pexpxy", "stepwtd")]


def _shape(n, ni, nj):
    L = _capi.INIT_LAYERS.get(n, 1)
    return (nj, ni) if L == 1 else (nj, L, ni)


def init_case(name, ni, nj, iopt_run=1, seed_fill=-777.0):
    """Raw cold-start inputs of a synthetic tile + every other NOAHMP_INIT array pre-filled with a sentinel."""
    cfg = S.named_config(name)
    cfg.ni, cfg.nj = ni, nj
    xp = S.backend()
    st = S.static_fields(xp, cfg)
    frc1 = S.forcing(xp, cfg, 1, st)
    raw = S.raw_initial_fields(cfg, st, frc1)
    A = {}
    for n, k in _capi.INIT_SPEC:
        if k == "pf" and n != "dzs":
            A[n] = np.full(_shape(n, ni, nj), seed_fill, np.float32)
    A["isnowxy"] = np.full((nj, ni), 9, np.int32)
    A["dzs"] = S.DZS.copy()
    for n in ("tsk", "tslb", "smois", "snow", "snowh"):
        A[n] = raw[n].copy()
    A["isltyp"], A["ivgtyp"] = st["isltyp"].copy(), st["ivgtyp"].copy()
    A["xice"], A["tmn"] = st["xice"].copy(), st["tmn"].copy()
    sc = dict(isurban=S.ISURBAN, isice=S.ISICE, iswater=S.ISWATER, fndsoilw=0, fndsnowh=1, nsoil=4, restart=0,
              allowed_to_read=1, iopt_run=iopt_run, dx=1000.0, dy=1000.0, wtddt=30.0, dt=float(cfg.dt),
              ids=1, ide=ni + 1, jds=1, jde=nj + 1, kds=1, kde=2, ims=1, ime=ni, jms=1, jme=nj, kms=1, kme=2,
              its=1, ite=ni, jts=1, jte=nj, kts=1, kte=2)
    if iopt_run == 5:
        dummy_state = {n: np.zeros(_shape(n if n != "smoiseq" else "smoiseq", ni, nj), np.float32)
                       for n in ("zwtxy", "smcwtdxy", "smoiseq", "waxy", "deeprechxy", "rechxy")}
        dummy_state["smois"], dummy_state["sh2o"] = A["smois"], A["smois"]
        gw, _ = S.groundwater_fields(cfg, st, dummy_state)
        A["fdepthxy"], A["ht"], A["rivercondxy"], A["pexpxy"] = gw["fdepth"], gw["topo"], gw["rivercond"], gw["pexp"]
        A["msftx"] = np.ones((nj, ni), np.float32)
        A["msfty"] = np.full((nj, ni), 1.02, np.float32)
        # a water table that follows the terrain smoothly (white noise here would make the Newton iteration for the
        # deep soil moisture chase fluxes no soil can carry), near its equilibrium depth
        wtd = (np.float32(-6.0) + np.float32(0.02) * (gw["topo"] - np.float32(400.0))).astype(np.float32)
        A["eqzwt"] = (wtd + np.float32(0.3) * (gw["fdepth"] / np.float32(500.0) - np.float32(0.5))).astype(np.float32)
        A["riverbedxy"] = (A["eqzwt"] - np.float32(1.0)).astype(np.float32)
        wtd[::7, ::5] = -0.7    # water table inside the resolved soil layers
        wtd[3::7, 2::5] = -2.4  # between the bottom of the soil and the deep layer
        A["zwtxy"] = wtd
    else:
        for n in _capi.INIT_GW:
            A.pop(n, None)
    return cfg, st, frc1, A, sc


def clone(A):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in A.items()}


def test_oracle_init_matches_numpy_restatement(built, tables_usgs_struct, tables_usgs):
    cfg, st, frc1, A, sc = init_case("C4", 96, 64)
    ref = S.cold_start(cfg, st, frc1, tables_usgs)
    rc, step = O.init(A, sc, tables_usgs_struct)
    assert rc == 0 and step is None
    exact = ["snow", "snowh", "smois", "tslb", "tvxy", "tgxy", "canwat", "canliqxy", "canicexy", "fwetxy", "sneqvoxy",
             "alboldxy", "qsnowxy", "wslakexy", "zwtxy", "waxy", "wtxy", "isnowxy", "tsnoxy", "snliqxy", "lfmassxy",
             "rtmassxy", "stmassxy", "woodxy", "stblcpxy", "fastcpxy", "xsaixy", "t2mvxy", "t2mbxy"]
    for n in exact:
        assert np.array_equal(A[n], ref[n]), n
    # snow ice: SWE/SNODEP then two multiplies, same order in both restatements
    assert np.array_equal(A["snicexy"], ref["snicexy"])
    # layer depths above the top snow layer are left alone by SNOW_INIT (sentinel here, 0 in the numpy version)
    act = np.arange(-2, 5)[None, :, None] >= (A["isnowxy"][:, None, :] + 1)
    assert np.array_equal(A["zsnsoxy"][act], ref["zsnsoxy"][act])
    assert np.all(A["zsnsoxy"][~act] == -777.0)
    # supercooled liquid water: x**y of numpy's float32 pow vs libm powf
    assert np.allclose(A["sh2o"], ref["sh2o"], rtol=2e-6, atol=0)
    # values the driver overwrites right after NOAHMP_INIT (cold_start applies that), here still the INIT ones
    assert np.all(A["eahxy"] == 2000.0) and np.all(A["chstarxy"] == np.float32(0.1))
    assert np.all(A["cmxy"] == 0.0) and np.all(A["chxy"] == 0.0)
    assert np.array_equal(A["tahxy"], A["tgxy"])
    assert (A["isnowxy"] < 0).any() and (A["ivgtyp"] == S.ISICE).any()


def test_oracle_init_restart_and_errors(built, tables_usgs_struct):
    _, _, _, A, sc = init_case("C1", 10, 10)
    B = clone(A)
    rc, _ = O.init(B, dict(sc, restart=1), tables_usgs_struct)
    assert rc == 0
    for n in A:
        assert np.array_equal(A[n], B[n]), n  # a restart run leaves everything to the restart file
    B["isltyp"][4, 5] = 0
    rc, _ = O.init(B, sc, tables_usgs_struct)
    assert rc == 9  # NOAHMP_ERR_ISLTYP
    C_ = clone(A)
    rc, _ = O.init(C_, dict(sc, fndsnowh=0), tables_usgs_struct)
    assert rc == 0
    land = ~((A["ivgtyp"] == S.ISICE) & (A["xice"] <= 0))
    assert np.array_equal(C_["snowh"][land], (A["snow"] * np.float32(0.005))[land])


def test_oracle_groundwater_init_properties(built, tables_usgs_struct, tables_usgs):
    cfg, st, frc1, A, sc = init_case("C2", 60, 44, iopt_run=5)
    wtd0 = A["zwtxy"].copy()
    sm0 = A["smois"].copy()
    rc, step = O.init(A, sc, tables_usgs_struct)
    assert rc == 0 and step == 1  # nint(30 min * 60 / 3600 s) = 1 ... max(.,1)
    soil, veg = A["isltyp"], A["ivgtyp"]
    smcmax = np.where(veg == S.ISURBAN, np.float32(0.45), tables_usgs["maxsmc"][soil - 1])
    assert np.array_equal(A["areaxy"], np.float32(1000.0 * 1000.0) / (A["msftx"] * A["msfty"]))
    for n in ("deeprechxy", "rechxy", "qslatxy", "qrfsxy", "qspringsxy", "waxy", "wtxy"):
        assert np.all(A[n] == 0.0), n
    # equilibrium soil moisture: a root of (SMC-SMCMAX)*DWSAT/DDZ + DKSAT*(SMC/SMCMAX)**(B+1) within Newton's tolerance
    eq = A["smoiseq"]
    ok = (tables_usgs["bb"][soil - 1] > 0) & (smcmax > 0) & (tables_usgs["satpsi"][soil - 1] > 0)
    assert ok.any() and (~ok).any()
    assert np.all(eq[:, 0, :][~ok] == smcmax[~ok]) and np.all(A["zwtxy"][~ok] == 0.0)  # e.g. the water soil class
    okl = np.broadcast_to(ok[:, None, :], eq.shape)
    assert np.all(eq[okl] >= np.float32(1e-4)) and np.all((eq <= (smcmax * np.float32(0.99))[:, None, :])[okl])
    zs = -np.cumsum(S.DZS)
    ddz = np.array([-zs[1] * 0.5, (zs[0] - zs[2]) * 0.5, (zs[1] - zs[3]) * 0.5, zs[2] - zs[3]])
    dw, dk, bb = (tables_usgs[k][soil - 1].astype(np.float64) for k in ("satdw", "satdk", "bb"))
    interior = (eq > 1.1e-4) & (eq < (smcmax * 0.989)[:, None, :]) & okl
    for k in range(4):
        x = eq[:, k, :].astype(np.float64)
        with np.errstate(all="ignore"):
            func = (x - smcmax) * dw / ddz[k] + dk * (x / smcmax) ** (bb + 1.0)
        scale = dw / ddz[k] * smcmax
        with np.errstate(all="ignore"):  # the non-soil classes (masked out below) divide by zero
            rel = np.abs(func / scale)
        assert np.all(rel[interior[:, k, :]] < 5e-4), k
    # deep soil moisture
    deep = (wtd0 < np.float32(zs[3] - S.DZS[3])) & ok
    # the reference's Newton iteration for the deep soil moisture has no safeguard: where the lateral flux asks for
    # more than the soil can carry it leaves NaN behind (a handful of cells here); everything else must be finite
    lost = np.isnan(A["smcwtdxy"])
    assert lost.mean() < 0.005
    assert not any(np.isnan(A[n]).any() for n in A if A[n].dtype == np.float32 and n != "smcwtdxy")
    deep &= ~lost
    assert deep.any() and np.all(A["smcwtdxy"][deep] >= np.float32(1e-4)) and np.all((A["smcwtdxy"] <= smcmax * 1.01)[~lost])
    assert np.array_equal(A["zwtxy"][deep], wtd0[deep])
    # water table inside the soil: layers wholly below it are saturated, the table is re-diagnosed
    shallow = (wtd0 >= np.float32(zs[3])) & ok
    assert shallow.any()
    assert np.all(A["smcwtdxy"][shallow] == smcmax[shallow])
    l4 = shallow & (wtd0 >= np.float32(zs[2]))
    assert l4.any() and np.all(A["smois"][:, 3, :][l4] == smcmax[l4])
    assert np.array_equal(A["smois"][:, 0, :], np.minimum(sm0[:, 0, :], tables_usgs["maxsmc"][soil - 1]))


def _gpu_model(tables, ni, nj):
    import noahmp_b200
    return noahmp_b200.NoahMP(tables, ni, nj)


@pytest.mark.gpu
@pytest.mark.parametrize("name,ni,nj,run", [("C4", 96, 64, 1), ("C2", 60, 44, 5), ("C3", 257, 130, 5)])
def test_gpu_init_bitexact(built, tables_usgs, tables_usgs_struct, name, ni, nj, run):
    _, _, _, A, sc = init_case(name, ni, nj, iopt_run=run)
    B = clone(A)
    O.set_math_mode(1)
    try:
        rc, step_o = O.init(A, sc, tables_usgs_struct)
    finally:
        O.set_math_mode(0)
    assert rc == 0
    m = _gpu_model(tables_usgs, ni, nj)
    step_g = m.init(B, sc)
    assert step_g == step_o
    for n in INIT_OUT:
        if n in A:
            assert np.array_equal(A[n], B[n], equal_nan=A[n].dtype == np.float32), n
    m.close()


@pytest.mark.gpu
def test_gpu_init_errors_and_restart(built, tables_usgs):
    import noahmp_b200
    _, _, _, A, sc = init_case("C1", 10, 10)
    m = _gpu_model(tables_usgs, 10, 10)
    B = clone(A)
    m.init(B, dict(sc, restart=1))
    for n in A:
        assert np.array_equal(A[n], B[n]), n
    B["isltyp"][2, 3] = 0
    with pytest.raises(noahmp_b200.NoahmpError) as e:
        m.init(B, sc)
    assert e.value.code == 9 and "ISLTYP" in str(e.value)
    assert np.all(B["tvxy"] == -777.0)  # nothing was written back
    with pytest.raises(noahmp_b200.NoahmpError):
        m.init(clone(A), dict(sc, iopt_run=5))  # groundwater arrays missing
    m.close()


@pytest.mark.gpu
def test_device_cold_start_equals_numpy_cold_start(built, tables_usgs):
    """bench.py starts from the library's own NOAHMP_INIT; the synthetic cases of the parity tests start from the numpy
    restatement.  Same state, up to the rounding of x**y in the supercooled-water formula."""
    cfg = S.named_config("C4")
    cfg.ni, cfg.nj = 120, 80
    xp = S.backend()
    st = S.static_fields(xp, cfg)
    frc1 = S.forcing(xp, cfg, 1, st)
    ref = S.cold_start(cfg, st, frc1, tables_usgs)
    m = _gpu_model(tables_usgs, cfg.ni, cfg.nj)
    dev = S.cold_start_device(m, cfg, st, frc1)
    m.close()
    for n in ref:
        if n == "sh2o":
            assert np.allclose(dev[n], ref[n], rtol=2e-6, atol=0), n
        else:
            assert np.array_equal(dev[n], ref[n]), n


def test_oracle_init_is_tile_independent(built, tables_usgs_struct):
    """Without the groundwater option the cold start is column-local: initialising two half tiles gives what the
    whole domain gives (the reference's MPI ranks each call NOAHMP_INIT on their own tile)."""
    _, _, _, A, sc = init_case("C4", 64, 40)
    whole = clone(A)
    assert O.init(whole, sc, tables_usgs_struct)[0] == 0
    nj, h = 40, 24
    for j0, j1 in ((0, h), (h, nj)):
        part = {}
        for n, v in A.items():
            part[n] = v.copy() if n == "dzs" else np.ascontiguousarray(v[j0:j1])
        s2 = dict(sc, jds=j0 + 1, jde=j1 + 1, jms=j0 + 1, jme=j1, jts=j0 + 1, jte=j1)
        assert O.init(part, s2, tables_usgs_struct)[0] == 0
        for n in INIT_OUT:
            if n in whole:
                assert np.array_equal(part[n], whole[n][j0:j1]), n


@pytest.mark.gpu
def test_groundwater_init_rejects_missing_soil_type_without_touching_the_tables(built, tables_usgs):
    """iopt_run=5 with a missing-field fill in ISLTYP (-9999): the reference stops in lsminit
    (noahmpdrv.F90:1008-1021); here NOAHMP_ERR_ISLTYP comes back with that message, the second init kernel does not
    index the soil tables with it, nothing is written back and the context stays usable."""
    import noahmp_b200
    cfg, st, frc1, A, sc = init_case("C2", 48, 32, iopt_run=5)
    m = noahmp_b200.NoahMP(tables_usgs, cfg.ni, cfg.nj, device=0)
    bad = clone(A)
    bad["isltyp"][5, 7] = -9999
    before = bad["tvxy"].copy()
    with pytest.raises(noahmp_b200.NoahmpError) as e:
        m.init(bad, sc)
    assert e.value.code == 9 and "ISLTYP" in str(e.value)
    assert np.array_equal(bad["tvxy"], before)
    good = clone(A)
    assert m.init(good, sc) == 1  # STEPWTD = max(nint(30*60/3600), 1): the context is still usable
    assert not np.array_equal(good["tvxy"], before)
    m.close()
