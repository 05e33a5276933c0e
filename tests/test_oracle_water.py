"""Known answers for the WATER / CARBON side of the oracle and for the glacier phase change, derived from the physics
(mass, carbon and energy budgets; the fixed point STOMATA solves) rather than from the code.  The reference ships no
golden vectors, so each routine is checked against a budget its inputs and outputs must close:

  CANWATER      precipitation = throughfall + drip + net canopy evaporation + change of canopy storage
  SNOWWATER     change of the pack's water = snowfall + frost - sublimation + rain - bottom outflow - glacier flow - ponding
                (COMPACT / COMBINE / DIVIDE move and merge layers without creating water), layer geometry consistent
  SOILWATER     change of the column's liquid water = infiltration - evaporation - transpiration - runoff - drainage,
                for every runoff option that closes the soil column itself
  CO2FLUX       change of all carbon pools = assimilation - every respiration and loss term
  STOMATA       the returned (RS, PSN) satisfy the Ball-Berry quadratic and the Farquhar minimum at the same CI
  PHASECHANGE_GLACIER   sensible heat lost = latent heat of the ice melted, layer by layer
"""
import ctypes as C

import numpy as np
import pytest

from noahmp_b200 import _capi

pf, pi = C.POINTER(C.c_float), C.POINTER(C.c_int)
f32 = np.float32


@pytest.fixture(scope="module")
def L(built):
    from oracle import oracle
    lib = oracle.lib()
    pt = C.POINTER(_capi.NoahmpTables)
    lib.nmo_canwater.argtypes = [pt, C.c_int, C.c_int, pf, pf, pf]
    lib.nmo_snowwater.argtypes = [pi, pf, pf, pf, pi, pf, pf, pf, pf, pf, pf, pf, pf, pf]
    lib.nmo_soilwater.argtypes = [pt, C.c_int, C.c_int, C.c_int, pf, pf, pf, pf, pf, pf, pf, pf]
    lib.nmo_soilwater.restype = C.c_int
    lib.nmo_co2flux.argtypes = [pt, C.c_int, pf, pf, pf]
    lib.nmo_stomata.argtypes = [pt, C.c_int, pf, pf]
    lib.nmo_phasechange_glacier.argtypes = [C.c_int, C.c_float, pf, pf, pf, pf, pf, pf, pf, pf, pf, pf, pi, pf]
    oracle.set_math_mode(0)
    return lib


def P(a):
    return a.ctypes.data_as(pi if a.dtype == np.int32 else pf)


@pytest.mark.parametrize("opt_snf", [1, 2, 3])
@pytest.mark.parametrize("seed", range(8))
def test_canwater_closes_the_canopy_water_budget(L, tables_usgs_struct, opt_snf, seed):
    rng = np.random.default_rng(100 + seed)
    dt = 3600.0
    sfctmp = rng.uniform(262.0, 285.0)
    frozen = float(rng.uniform() < 0.4)
    tv = rng.uniform(265.0, 272.9) if frozen else rng.uniform(273.3, 290.0)
    prec = rng.choice([0.0, 2e-4, 1.5e-3])
    inp = np.array([dt, sfctmp, rng.normal(0, 3), rng.normal(0, 3), rng.uniform(-30, 80), rng.uniform(0, 100), 0.1 * prec,
                    0.9 * prec, rng.uniform(0.5, 4.0), rng.uniform(0.1, 1.0), 275.0, rng.uniform(0.2, 0.95), frozen], f32)
    io = np.array([rng.uniform(0, 0.3), rng.uniform(0, 0.5), tv], f32)
    io0 = io.copy()
    out = np.zeros(8, f32)
    L.nmo_canwater(C.byref(tables_usgs_struct), opt_snf, int(rng.integers(2, 15)), P(inp), P(io), P(out))
    cmc, ecan, etran, qrain, qsnow, snowhin, fwet, fpice = out.astype(np.float64)
    assert cmc == pytest.approx(float(io[0]) + float(io[1]), abs=1e-6)
    storage = cmc - (float(io0[0]) + float(io0[1]))
    resid = prec * dt - (qrain + qsnow) * dt - ecan * dt - storage
    assert abs(resid) < 2e-4 * max(1.0, prec * dt), (resid, prec * dt, storage)
    assert 0.0 <= fwet <= 1.0 and 0.0 <= fpice <= 1.0 and qrain >= 0 and qsnow >= 0
    assert etran == pytest.approx(max(float(inp[5]), 0.0) / (2.8440e6 if frozen else 2.5104e6), rel=1e-5)
    if opt_snf == 3:
        assert fpice == (0.0 if sfctmp >= 273.16 else 1.0)
    assert io[0] >= 0 and io[1] >= 0


def _pack(rng, nlay):
    """a physically sensible nlay-layer pack (0 = none) in the oracle's array convention"""
    dz = np.zeros(7, f32); stc = np.zeros(7, f32); ice = np.zeros(3, f32); liq = np.zeros(3, f32)
    zsoil = np.array([-0.1, -0.4, -1.0, -2.0], f32)
    thick = {0: [], 1: [0.04], 2: [0.05, 0.2], 3: [0.05, 0.2, 0.6]}[nlay]
    for k, t in enumerate(thick):
        j = 3 - nlay + k
        dz[j] = t
        stc[j] = rng.uniform(262.0, 272.5)
        rho = rng.uniform(120.0, 350.0)
        ice[j] = rho * t
        liq[j] = rng.uniform(0.0, 0.02) * 1000.0 * t
    dz[3:] = [0.1, 0.3, 0.6, 1.0]
    stc[3:] = rng.uniform(270.0, 276.0, 4)
    zsn = np.zeros(7, f32)
    acc = 0.0
    for j in range(3 - nlay, 7):
        acc -= dz[j]
        zsn[j] = acc
    return dz, stc, ice, liq, zsoil, zsn


@pytest.mark.parametrize("nlay", [0, 1, 2, 3])
@pytest.mark.parametrize("seed", range(6))
def test_snowwater_conserves_water_and_keeps_the_layer_geometry(L, nlay, seed):
    rng = np.random.default_rng(7 * nlay + seed)
    dz, stc, ice, liq, zsoil, zsn = _pack(rng, nlay)
    dt = 3600.0
    snowing = rng.uniform() < 0.6
    qsnow = rng.uniform(1e-4, 3e-3) if snowing else 0.0
    snowhin = qsnow / rng.uniform(70.0, 120.0)
    qrain = 0.0 if snowing else rng.choice([0.0, 5e-4])
    qsnfro, qsnsub = rng.uniform(0, 2e-5), rng.uniform(0, 2e-5)
    sc = np.array([dt, 268.0, snowhin, qsnow, qsnfro, qsnsub, qrain], f32)
    imelt = np.zeros(7, np.int32)
    ficeold = np.where(ice + liq > 0, ice / np.maximum(ice + liq, 1e-20), 0).astype(f32)
    isnow = np.array([-nlay], np.int32)
    sneqv0 = float(ice.sum() + liq.sum())
    snowh0 = float(dz[:3].sum())
    io2 = np.array([snowh0, sneqv0], f32)
    sh2o = rng.uniform(0.15, 0.3, 4).astype(f32); sice = rng.uniform(0.0, 0.05, 4).astype(f32)
    soil1_0 = (float(sh2o[0]) + float(sice[0])) * 0.1 * 1000.0
    dzs = dz.copy()
    out = np.zeros(4, f32)
    L.nmo_snowwater(P(imelt), P(sc), P(zsoil), P(ficeold), P(isnow), P(io2), P(ice), P(liq), P(sh2o), P(sice), P(stc), P(zsn),
                    P(dzs), P(out))
    qsnbot, snoflow, pond1, pond2 = out.astype(np.float64)
    n1 = -int(isnow[0])
    assert 0 <= n1 <= 3
    sneqv1 = float(io2[1])
    if n1 > 0:
        assert sneqv1 == pytest.approx(float(ice.sum() + liq.sum()), rel=2e-6)
        assert np.all(ice[:3 - n1] == 0) and np.all(liq[:3 - n1] == 0) and np.all(dzs[:3 - n1] == 0)
        assert np.all(ice[3 - n1:] > 0) and np.all(dzs[3 - n1:3] > 0)
        z = -np.cumsum(dzs[3 - n1:])  # ZSNSO: depth of every layer bottom below the snow surface (negative)
        assert np.allclose(zsn[3 - n1:], z, rtol=1e-5, atol=1e-6)
    soil1_1 = (float(sh2o[0]) + float(sice[0])) * 0.1 * 1000.0
    rain_in = qrain * dt if n1 > 0 else 0.0  # rain reaches the pack only while a layer exists
    gained = (qsnow + qsnfro - qsnsub) * dt + rain_in
    lost = qsnbot * dt + snoflow * dt + pond1 + pond2 + (soil1_1 - soil1_0)
    if nlay > 0 and n1 == 0 and qrain > 0:
        pytest.skip("the pack vanished inside the call: rain is routed by WATER, not by SNOWWATER")
    assert (sneqv1 - sneqv0) == pytest.approx(gained - lost, abs=2e-3 + 2e-6 * sneqv0), (sneqv0, sneqv1, gained, lost)
    assert qsnbot >= 0 and pond1 >= 0 and pond2 >= 0


@pytest.mark.parametrize("opt_run", [1, 3, 4])
@pytest.mark.parametrize("soiltyp", [2, 6, 10])
@pytest.mark.parametrize("seed", range(4))
def test_soilwater_closes_the_column_water_budget(L, tables_usgs_struct, opt_run, soiltyp, seed):
    rng = np.random.default_rng(31 * opt_run + seed)
    dt = 3600.0
    zsoil = np.array([-0.1, -0.4, -1.0, -2.0], f32)
    dz = np.array([0.1, 0.3, 0.6, 1.0])
    qinsur = float(rng.choice([0.0, 2e-7, 2e-6]))          # m/s at the soil surface
    qseva = float(rng.uniform(0, 3e-8))
    etrani = (rng.uniform(0, 1.5e-8, 4) * np.array([1, 1, 1, 0])).astype(f32)
    frozen = rng.uniform() < 0.3
    sice = (rng.uniform(0.0, 0.08, 4) if frozen else np.zeros(4)).astype(f32)
    sh2o = rng.uniform(0.18, 0.30, 4).astype(f32)
    smc = (sh2o + sice).astype(f32)
    w0 = float((sh2o.astype(np.float64) * dz).sum() * 1000.0)
    io3 = np.array([2.5, 0.3, 0.0], f32)
    out = np.zeros(4, f32)
    sc = np.array([dt, qinsur, qseva], f32)
    rc = L.nmo_soilwater(C.byref(tables_usgs_struct), opt_run, 1, soiltyp, P(sc), P(zsoil), P(etrani), P(sice), P(sh2o), P(smc),
                         P(io3), P(out))
    assert rc == 0
    runsrf, qdrain, runsub, fcrmax = out.astype(np.float64)
    w1 = float((sh2o.astype(np.float64) * dz).sum() * 1000.0)
    src = (qinsur - qseva - float(etrani.astype(np.float64).sum())) * dt * 1000.0
    resid = (w1 - w0) - (src - runsrf * dt - qdrain * dt - runsub * dt)
    assert abs(resid) < 5e-3 + 2e-6 * w0, (resid, w0, w1, runsrf, qdrain, runsub)
    assert runsrf >= 0 and 0 <= fcrmax <= 1
    assert np.all(sh2o > 0)
    if qinsur == 0.0:
        assert runsrf == pytest.approx(0.0, abs=1e-9)
    if not frozen:
        assert fcrmax == 0.0


@pytest.mark.parametrize("vegtyp", [2, 5, 7, 11, 14])
@pytest.mark.parametrize("seed", range(4))
def test_co2flux_closes_the_carbon_budget(L, tables_usgs, tables_usgs_struct, vegtyp, seed):
    """All pools together change by assimilation minus every respiration / loss term.  AUTORS leaves out stem
    maintenance and growth respiration and FASTCP never receives the dying stem mass (the reference's own omissions,
    noahmplsm.F90:9066, :9048): both are recomputed here from the inputs, as is everything that does not depend on
    the routine's outputs."""
    rng = np.random.default_rng(17 * vegtyp + seed)
    T = tables_usgs
    v = vegtyp - 1
    dt = 3600.0
    igs, stc1, psn, tv = 1.0, rng.uniform(276, 295), rng.uniform(18.0, 32.0), rng.uniform(285.0, 298.0)
    wroot, wstres, foln = rng.uniform(0.3, 0.8), rng.uniform(0.0, 0.5), 1.0
    lapm = float(T["sla"][v]) / 1000.0
    lfmass, rtmass, stmass = rng.uniform(12, 30), rng.uniform(200, 600), rng.uniform(20, 60)
    fastcp, stblcp, wood = rng.uniform(500, 1500), rng.uniform(500, 1500), rng.uniform(100, 800)
    xlai = max(lfmass * lapm, 0.05)
    sc = np.array([igs, dt, stc1, psn, tv, wroot, wstres, foln, lapm], f32)
    pools = np.array([xlai, 0.1, lfmass, rtmass, stmass, fastcp, stblcp, wood], f32)
    p0 = pools.astype(np.float64).copy()
    out = np.zeros(7, f32)
    L.nmo_co2flux(C.byref(tables_usgs_struct), vegtyp, P(sc), P(pools), P(out))
    gpp, npp, nee, autors, heters, totsc, totlb = out.astype(np.float64)
    p1 = pools.astype(np.float64)
    wdpool = float(T["wdpool"][v])
    assert gpp == pytest.approx(psn * 12e-6, rel=1e-6)
    assert nee == pytest.approx((autors + heters - gpp) * 44.0 / 12.0, rel=1e-4, abs=1e-12)
    assert totsc == pytest.approx(p1[5] + p1[6], rel=1e-6) and totlb == pytest.approx(p1[2] + p1[3] + p1[7], rel=1e-6)
    assert pools[0] == pytest.approx(max(p1[2] * lapm, 0.05), rel=1e-6)
    # terms AUTORS / FASTCP leave out, from the inputs (noahmplsm.F90:8905-8990)
    tf = float(T["arm"][v]) ** ((tv - 298.16) / 10.0)
    rsstem = float(T["rms25"][v]) * (stmass * 1e-3) * tf * 1.0 * 12e-6
    stempt = xlai / 10.0
    carbfx = psn * 12e-6
    grstem = max(0.0, float(T["fragr"][v]) * (stempt * carbfx - rsstem))
    sc_ = np.exp(-0.3 * max(0.0, tv - float(T["tdlef"][v]))) * (stmass * 0 + lfmass / 120.0)
    sd_ = np.exp((wstres - 1.0) * 100.0)
    diest = stmass * 1e-6 * (float(T["dilefw"][v]) * sd_ + float(T["dilefc"][v]) * sc_)
    # the budget closes only while no allocation is clamped at zero (ADDNPPLF / ADDNPPST = MAX(0, ...), :8963-8964): the
    # inputs above keep the routine in that regime, checked here from the inputs
    leafpt = np.exp(0.01 * (1.0 - np.exp((0.50 if vegtyp == int(T["eblforest"]) else 0.75) * xlai)) * xlai) - stempt
    fnf = min(foln / max(1e-6, float(T["folnmx"][v])), 1.0)
    rsleaf = min(lfmass / dt, float(T["rmf25"][v]) * tf * fnf * xlai * 1.0 * (1.0 - wstres) * 12e-6)
    grleaf = max(0.0, float(T["fragr"][v]) * (leafpt * carbfx - rsleaf))
    assert leafpt * carbfx - grleaf - rsleaf > 0 and stempt * carbfx - grstem - rsstem > 0, "test inputs left the unclamped regime"
    if wdpool == 1.0 or wdpool == 0.0:
        d_pools = (p1[2] + p1[3] + p1[4] + p1[5] + p1[6] + (p1[7] if wdpool else 0.0)) - \
                  (p0[2] + p0[3] + p0[4] + p0[5] + p0[6] + (p0[7] if wdpool else 0.0))
        want = (gpp - autors - rsstem - grstem - diest - heters) * dt
        if wdpool == 0.0:
            # no wood pool: what the allocation sends to wood leaves the budget (WOOD = (...) * WDPOOL = 0)
            pytest.skip("vegetation type without a wood pool")
        assert d_pools == pytest.approx(want, abs=3e-3 + 1e-6 * abs(p0[2:].sum())), (d_pools, want)
    assert np.all(p1[2:] >= 0)


@pytest.mark.parametrize("vegtyp", [2, 7, 11, 14])
@pytest.mark.parametrize("apar", [20.0, 120.0, 400.0])
def test_stomata_returns_a_fixed_point_of_ci(L, tables_usgs, tables_usgs_struct, vegtyp, apar):
    T = tables_usgs
    v = vegtyp - 1
    tv, sfctmp, sfcprs = 293.0, 291.0, 95000.0
    ei, ea = 2330.0, 1400.0
    o2, co2 = 0.209 * sfcprs, 395e-6 * sfcprs
    rb, btran, igs, foln = 25.0, 0.8, 1.0, 1.0
    inp = np.array([apar, foln, tv, ei, ea, sfctmp, sfcprs, o2, co2, igs, btran, rb], f32)
    out = np.zeros(2, f32)
    L.nmo_stomata(C.byref(tables_usgs_struct), vegtyp, P(inp), P(out))
    rs, psn = float(out[0]), float(out[1])
    assert psn > 0 and rs > 0
    cf = sfcprs / (8.314 * sfctmp) * 1e6
    rlb, rs_ = rb / cf, rs / cf
    mp, bp, c3 = float(T["mp"][v]), float(T["bp"][v]), float(T["c3psn"][v])
    cs = max(co2 - 1.37 * rlb * sfcprs * psn, 1e-6)
    # (1) Ball-Berry with the leaf-boundary layer: RS solves  A r^2 + B r + C = 0
    A = mp * psn * sfcprs * ea / (cs * ei) + bp
    B = (mp * psn * sfcprs / cs + bp) * rlb - 1.0
    Cq = -rlb
    assert abs(A * rs_ * rs_ + B * rs_ + Cq) < 2e-4 * max(abs(B * rs_), abs(Cq))
    # (2) the CI this RS implies reproduces PSN through the Farquhar minimum (bisection stops at 0.05 Pa)
    ci = max(cs - psn * sfcprs * 1.65 * rs_, 0.0)
    tc = tv - 273.16
    kc = float(T["kc25"][v]) * float(T["akc"][v]) ** ((tc - 25) / 10)
    ko = float(T["ko25"][v]) * float(T["ako"][v]) ** ((tc - 25) / 10)
    awc, cp = kc * (1 + o2 / ko), 0.5 * kc / ko * o2 * 0.21
    fnf = min(foln / max(1e-6, float(T["folnmx"][v])), 1.0)
    vcmx = float(T["vcmx25"][v]) / (1 + np.exp((-2.2e5 + 710 * (tc + 273.16)) / (8.314 * (tc + 273.16)))) * fnf * btran * \
        float(T["avcmx"][v]) ** ((tc - 25) / 10)
    j = 4.6 * apar * float(T["qe25"][v])
    wj = max(ci - cp, 0) * j / (ci + 2 * cp) * c3 + j * (1 - c3)
    wc = max(ci - cp, 0) * vcmx / (ci + awc) * c3 + vcmx * (1 - c3)
    we = 0.5 * vcmx * c3 + 4000 * vcmx * ci / sfcprs * (1 - c3)
    assert min(wj, wc, we) * igs == pytest.approx(psn, rel=5e-3)
    # no light: the minimum conductance and no assimilation
    inp[0] = 0.0
    L.nmo_stomata(C.byref(tables_usgs_struct), vegtyp, P(inp), P(out))
    assert out[1] == 0.0 and float(out[0]) == pytest.approx(cf / bp, rel=1e-6)


@pytest.mark.parametrize("isnow", [0, -2, -3])
@pytest.mark.parametrize("seed", range(5))
def test_glacier_phasechange_trades_latent_for_sensible_heat(L, isnow, seed):
    """PHASECHANGE_GLACIER (glacier.F90:1635-1922): total water per layer unchanged up to the redistribution sweeps of
    the ice column, and over the whole column the sensible heat that disappears is the latent heat of the ice melted."""
    rng = np.random.default_rng(50 + seed - isnow)
    dt = f32(3600.0)
    dz = np.array([0.05, 0.2, 0.5, 0.1, 0.3, 0.6, 1.0], f32)
    nl = -isnow
    hcap = rng.uniform(0.4e6, 1.9e6, 7).astype(f32)
    fact = (dt / (hcap * dz)).astype(f32)
    stc = rng.uniform(270.5, 275.0, 7).astype(f32)
    stc[:3 - nl] = 0
    snice, snliq = np.zeros(3, f32), np.zeros(3, f32)
    snice[3 - nl:] = (rng.uniform(150, 350, nl) * dz[3 - nl:3]).astype(f32)
    snliq[3 - nl:] = (rng.uniform(0, 10, nl) * dz[3 - nl:3]).astype(f32)
    smc = np.ones(4, f32)                               # glacier "soil" = ice with liquid fraction SH2O
    sh2o = rng.uniform(0.0, 0.05, 4).astype(f32)
    stc0, ice0, liq0, sh0 = stc.copy(), snice.copy(), snliq.copy(), sh2o.copy()
    sneqv = C.c_float(float(snice.sum() + snliq.sum()) if nl else 30.0)
    snowh = C.c_float(float(dz[3 - nl:3].sum()) if nl else 0.1)
    sneqv0 = sneqv.value
    qmelt, ponding = C.c_float(0), C.c_float(0)
    imelt = np.zeros(7, np.int32)
    L.nmo_phasechange_glacier(isnow, dt, P(fact), P(dz), P(stc), P(snice), P(snliq), C.byref(sneqv), C.byref(snowh), P(smc),
                              P(sh2o), C.byref(qmelt), P(imelt), C.byref(ponding))
    LF = 0.3336e6
    # snow layers: mass conserved layer by layer
    if nl:
        assert np.allclose(snice + snliq, ice0 + liq0, rtol=2e-6, atol=1e-5)
    d_ice_snow = float((snice.astype(np.float64) - ice0).sum())
    if nl == 0:  # bulk snow on the glacier melts with QMELT
        d_ice_snow = float(sneqv.value) - sneqv0
        assert qmelt.value * float(dt) == pytest.approx(-d_ice_snow, rel=1e-4, abs=1e-4)
    ice_soil0 = ((1.0 - sh0.astype(np.float64)) * dz[3:] * 1000.0).sum()
    ice_soil1 = ((smc.astype(np.float64) - sh2o) * dz[3:] * 1000.0).sum()
    act = np.arange(7) >= 3 - nl
    sensible = float((hcap.astype(np.float64) * dz * (stc.astype(np.float64) - stc0))[act].sum())
    latent = LF * (d_ice_snow + (ice_soil1 - ice_soil0))
    assert sensible == pytest.approx(latent, rel=5e-3, abs=2000.0), (sensible, latent)
    assert np.all(smc == 1.0) and np.all(sh2o >= 0) and np.all(sh2o <= 1.0)
