"""The CPU oracle against everything that can pin it without the reference binary: analytic known
answers, the model's own conservation checks on every synthetic configuration, and the portable math it shares
with the GPU parity build.  (The pin against the reference's own text is tests/test_reference_pin.py.)"""
import ctypes as C

import numpy as np
import pytest

from noahmp_b200 import _capi, synthetic as S

from helpers import clone_state, diff_report, make_case, run_oracle


@pytest.fixture(scope="module")
def O(built):
    from oracle import oracle
    oracle.lib()
    return oracle


def test_esat_known_answer(O):
    """ESAT at 0 C: the constant terms of the polynomials (noahmplsm.F90:5296-5314) -> 610.78 / 610.92 Pa."""
    out = (C.c_float * 4)()
    O.lib().nmo_esat(0.0, out)
    assert abs(out[0] - 610.7799961) < 1e-3 and abs(out[1] - 610.9177956) < 1e-3
    assert abs(out[2] - 44.38099984) < 1e-4 and abs(out[3] - 50.30305237) < 1e-4
    O.lib().nmo_esat(20.0, out)  # 20 C over water: 2338-2340 Pa in every standard table
    assert 2330.0 < out[0] < 2345.0


@pytest.mark.parametrize("zlvl,zpd,z0m,z0h,ur", [(30.0, 0.0, 0.01, 0.01, 5.0), (30.0, 13.0, 1.1, 1.1, 3.0),
                                                  (10.0, 0.65, 0.06, 0.006, 8.0)])
def test_sfcdif1_neutral_pass_is_the_log_law(O, zlvl, zpd, z0m, z0h, ur):
    """First pass of SFCDIF1 (no stability correction): the textbook neutral drag laws
    CM = k^2 / ln((z-d)/z0m)^2,  CH = k^2 / (ln((z-d)/z0m) ln((z-d)/z0h)),  u* = U sqrt(CM)."""
    L = O.lib()
    L.nmo_sfcdif1_neutral.argtypes = [C.c_float] * 5 + [C.POINTER(C.c_float)]
    out = (C.c_float * 4)()
    L.nmo_sfcdif1_neutral(zlvl, zpd, z0m, z0h, ur, out)
    k = 0.40
    lm, lh = np.log((zlvl - zpd) / z0m), np.log((zlvl - zpd) / z0h)
    assert out[0] == pytest.approx(k * k / lm ** 2, rel=2e-6)
    assert out[1] == pytest.approx(k * k / (lm * lh), rel=2e-6)
    assert out[2] == pytest.approx(ur * k / lm, rel=2e-6)
    # the 2-m exchange coefficient: k u* / ln((2+z0h)/z0h)
    assert out[3] == pytest.approx(k * (ur * k / lm) / np.log((2.0 + z0h) / z0h), rel=2e-6)


@pytest.mark.parametrize("t,smc,bexp,psisat,smcmax", [(268.0, 0.30, 5.33, 0.3548, 0.439), (272.5, 0.42, 4.74, 0.1413, 0.434),
                                                       (255.0, 0.20, 10.39, 0.4677, 0.468), (263.0, 0.35, 2.79, 0.0692, 0.339)])
def test_frh2o_solves_the_freezing_point_depression(O, t, smc, bexp, psisat, smcmax):
    """FRH2O (Koren et al. 1999): below freezing the liquid water left in the soil is the root of
    ln[(psisat g / Lf) (1 + 8 ice)^2 (smcmax / liq)^b] = ln[-(T - T0) / T], found here independently by bisection in
    fp64; the routine stops its Newton iteration at |d ice| <= 0.005.  At and above freezing all water is liquid."""
    L = O.lib()
    L.nmo_frh2o.argtypes = [C.c_float] * 6
    L.nmo_frh2o.restype = C.c_float
    free = L.nmo_frh2o(t, smc, smc, bexp, psisat, smcmax)
    g, lf, t0, ck = 9.80616, 0.3336e6, 273.16, 8.0
    bx = min(bexp, 5.5)

    def resid(ice):
        return (np.log((psisat * g / lf) * (1.0 + ck * ice) ** 2 * (smcmax / (smc - ice)) ** bx) - np.log(-(t - t0) / t))
    lo, hi = 0.0, smc - 0.02
    if resid(lo) * resid(hi) > 0:  # no root inside: the routine clamps to the nearer end
        exact = lo if abs(resid(lo)) < abs(resid(hi)) else hi
    else:
        for _ in range(80):
            mid = 0.5 * (lo + hi)
            if resid(lo) * resid(mid) <= 0:
                hi = mid
            else:
                lo = mid
        exact = 0.5 * (lo + hi)
    assert abs((smc - free) - exact) <= 0.006
    assert 0.0 < free <= smc
    assert L.nmo_frh2o(273.2, smc, smc, bexp, psisat, smcmax) == np.float32(smc)
    # colder soil keeps less liquid water
    assert L.nmo_frh2o(t - 5.0, smc, smc, bexp, psisat, smcmax) <= free + 1e-6


def test_esat_against_the_magnus_formula(O):
    """Saturation vapour pressure over water and over ice between -40 and +40 C: within 0.6 % of the Magnus /
    Alduchov-Eskridge formulas (an independent fit of the same physical curve)."""
    out = (C.c_float * 4)()
    for t in np.arange(-40.0, 40.1, 5.0):
        O.lib().nmo_esat(C.c_float(t), out)
        ew = 610.94 * np.exp(17.625 * t / (t + 243.04))
        assert out[0] == pytest.approx(ew, rel=6e-3), t
        if t <= 0.0:
            ei = 611.21 * np.exp(22.587 * t / (t + 273.86))
            assert out[1] == pytest.approx(ei, rel=6e-3), t
        # d(es)/dT is a separate polynomial fit of the slope: within 1.5 % of a central difference of the first
        O.lib().nmo_esat(C.c_float(t + 0.5), out); hi = out[0]
        O.lib().nmo_esat(C.c_float(t - 0.5), out); lo = out[0]
        O.lib().nmo_esat(C.c_float(t), out)
        assert out[2] == pytest.approx(hi - lo, rel=1.5e-2), t


@pytest.mark.parametrize("smcmax,quartz", [(0.339, 0.92), (0.439, 0.40), (0.468, 0.25), (0.476, 0.10)])
def test_tdfcnd_limits_of_the_johansen_conductivity(O, smcmax, quartz):
    """TDFCND (Peters-Lidard et al. 1998 / Johansen 1975): the dry limit is (0.135 rho_d + 64.7) / (2700 - 0.947 rho_d),
    the saturated unfrozen limit k_s^(1-n) k_w^n with k_s = 7.7^q 2.0^(1-q), the saturated frozen limit k_s^(1-n) k_ice^n,
    and conductivity grows with wetness in between."""
    L = O.lib()
    L.nmo_tdfcnd.argtypes = [C.c_float] * 4
    L.nmo_tdfcnd.restype = C.c_float
    rho_d = (1.0 - smcmax) * 2700.0
    dry = (0.135 * rho_d + 64.7) / (2700.0 - 0.947 * rho_d)
    ks = 7.7 ** quartz * 2.0 ** (1.0 - quartz)
    assert L.nmo_tdfcnd(0.02 * smcmax, 0.02 * smcmax, smcmax, quartz) == pytest.approx(dry, rel=1e-5)   # Ke = 0 below 10 %
    assert L.nmo_tdfcnd(smcmax, smcmax, smcmax, quartz) == pytest.approx(ks ** (1 - smcmax) * 0.57 ** smcmax, rel=1e-5)
    assert L.nmo_tdfcnd(smcmax, 1e-9, smcmax, quartz) == pytest.approx(ks ** (1 - smcmax) * 2.2 ** smcmax, rel=1e-4)
    w = [L.nmo_tdfcnd(f * smcmax, f * smcmax, smcmax, quartz) for f in (0.15, 0.3, 0.5, 0.7, 0.9, 1.0)]
    assert all(b > a for a, b in zip(w, w[1:])) and dry < w[0]
    assert 0.1 < dry < 0.4 and 0.8 < w[-1] < 3.0   # W/m/K: the range soils have


@pytest.mark.parametrize("ic", [0, 1])
@pytest.mark.parametrize("vegtyp,cosz,vai,rho,tau,alb", [(2, 0.8, 3.0, 0.11, 0.07, 0.15), (14, 0.3, 1.2, 0.07, 0.05, 0.6),
                                                         (7, 0.55, 6.0, 0.45, 0.34, 0.25), (11, 0.05, 0.4, 0.10, 0.10, 0.9)])
def test_twostream_closed_form_solves_the_two_stream_equations(O, tables_usgs, ic, vegtyp, cosz, vai, rho, tau, alb):
    """TWOSTREAM (noahmplsm.F90:2711-2957) evaluates the closed-form solution (coefficients H1..H10) of the canopy
    two-stream equations of Dickinson / Sellers.  Independent check: integrate the same boundary-value problem
      -mu dIup/dx + b Iup - c Idn = d exp(-K x),    mu dIdn/dx + b Idn - c Iup = f exp(-K x),   0 <= x <= VAI,
    numerically in fp64 (scipy.solve_bvp) and compare albedo and transmitted diffuse flux."""
    from scipy.integrate import solve_bvp
    ts = _capi.tables_from_dict(tables_usgs)
    L = O.lib()
    L.nmo_twostream.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_float] * 9 + [C.POINTER(C.c_float)]
    out = (C.c_float * 5)()
    L.nmo_twostream(C.addressof(ts), 2, 1, ic, vegtyp, cosz, vai, 0.0, 290.0, alb, alb, rho, tau, 1.0, out)
    fab, fre, ftd, fti, gdir = (float(v) for v in out)
    # coefficients of the equations (CLM technical note, section 3.1)
    mu = max(0.001, cosz)
    chil = min(max(float(tables_usgs["xl"][vegtyp - 1]), -0.4), 0.6)
    if abs(chil) <= 0.01:
        chil = 0.01
    phi1 = 0.5 - 0.633 * chil - 0.330 * chil * chil
    phi2 = 0.877 * (1.0 - 2.0 * phi1)
    g = phi1 + phi2 * mu
    K = g / mu
    avmu = (1.0 - phi1 / phi2 * np.log((phi1 + phi2) / phi1)) / phi2
    om = rho + tau
    asu = 0.5 * om * g / (g + phi2 * mu) * (1.0 - phi1 * mu / (g + phi2 * mu) * np.log((phi1 * mu + g + phi2 * mu) / (phi1 * mu)))
    beta0 = (1.0 + avmu * K) / (om * avmu * K) * asu
    beta = 0.5 * (rho + tau + (rho - tau) * ((1.0 + chil) / 2.0) ** 2) / om
    b, c = 1.0 - om + om * beta, om * beta
    d, f = (avmu * K * om * beta0, avmu * K * om * (1.0 - beta0)) if ic == 0 else (0.0, 0.0)
    assert gdir == pytest.approx(g, rel=1e-6)

    def rhs(x, y):
        e = np.exp(-K * x)
        return np.vstack([(b * y[0] - c * y[1] - d * e) / avmu, (-b * y[1] + c * y[0] + f * e) / avmu])

    def bc(ya, yb):
        if ic == 0:
            return np.array([ya[1], yb[0] - alb * (yb[1] + np.exp(-K * vai))])
        return np.array([ya[1] - 1.0, yb[0] - alb * yb[1]])
    x = np.linspace(0.0, vai, 400)
    sol = solve_bvp(rhs, bc, x, np.zeros((2, x.size)) + 0.1, tol=1e-10, max_nodes=200000)
    assert sol.success
    assert fre == pytest.approx(sol.y[0, 0], abs=1e-4)  # fp32 closed form with cancellations vs fp64 integration
    assert fti == pytest.approx(sol.y[1, -1], abs=1e-4)
    assert ftd == pytest.approx(np.exp(-K * vai) if ic == 0 else 0.0, abs=1e-6)
    # what is neither reflected nor absorbed by the ground is absorbed by the canopy
    assert fab == pytest.approx(1.0 - fre - (1.0 - alb) * (ftd + fti), abs=1e-6) and -1e-6 <= fab <= 1.0


@pytest.mark.parametrize("isnow", [0, -2, -3])
@pytest.mark.parametrize("opt_tbot", [1, 2])
def test_tsnosoi_conserves_heat(O, isnow, opt_tbot):
    """One TSNOSOI step (HRT + HSTEP, noahmplsm.F90:5711-5977) is a conservative implicit diffusion step: the heat
    content of the snow/soil column changes by DT * (ground heat flux in at the top + flux through the bottom), the
    bottom flux being zero (opt_tbot=1) or the explicit -DF (T_N - TBOT) / (z_N - ZBOT) of opt_tbot=2.  A uniform
    column with no flux stays as it is."""
    L = O.lib()
    pf = C.POINTER(C.c_float)
    L.nmo_tsnosoi.argtypes = [C.c_int] * 3 + [C.c_float, pf, C.c_float, pf, pf, C.c_float, C.c_float, C.c_float, pf]
    f32 = np.float32
    dzsnow = {0: [], -2: [0.06, 0.11], -3: [0.05, 0.20, 0.31]}[isnow]
    dzs = np.array(dzsnow + [0.1, 0.3, 0.6, 1.0], np.float64)
    n = dzs.size
    # ZSNSO is the depth of every layer bottom below the snow surface (negative downward)
    z = np.zeros(7, f32); z[7 - n:] = -np.cumsum(dzs)
    rng = np.random.default_rng(5 + isnow)
    df = np.zeros(7, f32); df[7 - n:] = rng.uniform(0.1, 2.5, n)
    hc = np.zeros(7, f32); hc[7 - n:] = rng.uniform(0.5e6, 3.0e6, n)
    stc0 = np.zeros(7, f32); stc0[7 - n:] = rng.uniform(255.0, 285.0, n)
    snowh, zbot, tbot, ssoil, dt = f32(sum(dzsnow)), f32(-8.0), f32(279.0), f32(37.5), f32(1800.0)
    stc = stc0.copy()
    L.nmo_tsnosoi(1, opt_tbot, isnow, tbot, z.ctypes.data_as(pf), ssoil, df.ctypes.data_as(pf), hc.ctypes.data_as(pf),
                  zbot, dt, snowh, stc.ctypes.data_as(pf))
    dh = float(np.sum(hc[7 - n:].astype(np.float64) * dzs * (stc[7 - n:].astype(np.float64) - stc0[7 - n:])))
    bot = 0.0
    if opt_tbot == 2:
        zmid = 0.5 * (float(z[5]) + float(z[6]))
        bot = -float(df[6]) * (float(stc0[6]) - float(tbot)) / (zmid - (float(zbot) - float(snowh)))
    want = float(dt) * (float(ssoil) + bot)
    assert dh == pytest.approx(want, rel=2e-4, abs=50.0)   # J/m2; fp32 tridiagonal solve
    assert np.all(stc[:7 - n] == 0.0)
    # no gradients, no fluxes: nothing moves
    flat = np.zeros(7, f32); flat[7 - n:] = 271.5
    L.nmo_tsnosoi(1, 2, isnow, f32(271.5), z.ctypes.data_as(pf), f32(0.0), df.ctypes.data_as(pf), hc.ctypes.data_as(pf), zbot,
                  dt, snowh, flat.ctypes.data_as(pf))
    assert np.allclose(flat[7 - n:], 271.5, atol=2e-4)


@pytest.mark.parametrize("opt_frz", [1, 2])
@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_phasechange_trades_latent_for_sensible_heat_exactly(O, opt_frz, seed):
    """PHASECHANGE (noahmplsm.F90:6039-6245) on a snow-free soil column: water mass per layer is unchanged, liquid water
    stays within [0, SMC], and in every layer the sensible heat that disappears is the latent heat of the ice that
    melts:  C dz (T_new - T_old) = Lf (ice_new - ice_old),  with C dz = DT / FACT."""
    L = O.lib()
    pf, pi = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.nmo_phasechange.argtypes = [C.c_int, C.c_int, C.c_float, pf, pf, pf, pf, pf, pf, pf, pf, pf, C.c_float, C.c_float,
                                  C.c_float, pf, pi, pf]
    f32 = np.float32
    rng = np.random.default_rng(seed)
    dz = np.array([0, 0, 0, 0.1, 0.3, 0.6, 1.0], f32)
    hcap = rng.uniform(1.2e6, 3.0e6, 7).astype(f32)
    dt = f32(1800.0)
    fact = np.zeros(7, f32); fact[3:] = dt / (hcap[3:] * dz[3:])
    stc0 = np.zeros(7, f32); stc0[3:] = rng.uniform(271.0, 275.5, 4)
    smc0 = rng.uniform(0.15, 0.42, 4).astype(f32)
    sh2o0 = (smc0 * rng.uniform(0.2, 1.0, 4)).astype(f32)
    stc, smc, sh2o = stc0.copy(), smc0.copy(), sh2o0.copy()
    snice, snliq = np.zeros(3, f32), np.zeros(3, f32)
    sneqv, snowh, qmelt, ponding = (C.c_float(0.0) for _ in range(4))
    imelt = np.zeros(7, np.int32)
    L.nmo_phasechange(opt_frz, 0, dt, fact.ctypes.data_as(pf), dz.ctypes.data_as(pf), stc.ctypes.data_as(pf),
                      snice.ctypes.data_as(pf), snliq.ctypes.data_as(pf), C.byref(sneqv), C.byref(snowh),
                      smc.ctypes.data_as(pf), sh2o.ctypes.data_as(pf), 5.33, 0.3548, 0.439, C.byref(qmelt),
                      imelt.ctypes.data_as(pi), C.byref(ponding))
    assert np.allclose(smc, smc0, rtol=3e-7, atol=0)           # total water of a layer does not change
    assert np.all(sh2o >= 0.0) and np.all(sh2o <= smc * (1 + 1e-6))
    ice0 = (smc0.astype(np.float64) - sh2o0) * dz[3:] * 1000.0   # kg/m2
    ice1 = (smc.astype(np.float64) - sh2o) * dz[3:] * 1000.0
    sensible = hcap[3:].astype(np.float64) * dz[3:] * (stc[3:].astype(np.float64) - stc0[3:])
    latent = 0.3336e6 * (ice1 - ice0)
    assert np.allclose(sensible, latent, rtol=2e-3, atol=300.0), (sensible, latent)   # J/m2, fp32 column
    moved = np.abs(ice1 - ice0) > 1e-2  # kg/m2; below that it is the fp32 round trip through layer masses
    assert moved.any()
    assert np.all(imelt[3:][moved] > 0)
    # melting needs T >= TFRZ, freezing T < TFRZ
    assert np.all(stc0[3:][(ice1 - ice0) < -1e-2] >= np.float32(273.16))
    assert np.all(stc0[3:][(ice1 - ice0) > 1e-2] < np.float32(273.16))
    assert qmelt.value == 0.0 and ponding.value == 0.0


@pytest.mark.parametrize("n", [4, 5, 6, 7])
def test_rosr12_matches_dense_solve(O, n):
    """ROSR12 (noahmplsm.F90:5979-6036) against numpy's dense solver on diagonally dominant systems."""
    rng = np.random.default_rng(n)
    for _ in range(20):
        a = rng.uniform(-1, 0, n).astype(np.float32); c = rng.uniform(-1, 0, n).astype(np.float32)
        a[0] = 0; c[-1] = 0
        b = (1 + np.abs(a) + np.abs(c) + rng.uniform(0, 1, n)).astype(np.float32)
        d = rng.uniform(-5, 5, n).astype(np.float32)
        x = np.zeros(n, np.float32)
        O.lib().nmo_rosr12(n, a.ctypes.data, b.ctypes.data, c.ctypes.data, d.ctypes.data, x.ctypes.data)
        M = np.diag(b.astype(np.float64)) + np.diag(a[1:].astype(np.float64), -1) + np.diag(c[:-1].astype(np.float64), 1)
        assert np.allclose(x, np.linalg.solve(M, d.astype(np.float64)), rtol=2e-5, atol=2e-5)


def test_combo_conserves_mass_and_follows_enthalpy_rule(O):
    """COMBO (noahmplsm.F90:7375-7424): layer thickness, liquid and ice add; the merged temperature follows the
    reference's three-branch enthalpy rule (HC < 0: sensible heat only; 0 <= HC <= HFUS*WLIQ: TFRZ; else warm)."""
    CICE, CWAT, HFUS, TFRZ = 2.094e6, 4.188e6, 0.3336e6, 273.16
    rng = np.random.default_rng(0)
    seen = set()
    for _ in range(400):
        v = np.array([rng.uniform(.01, .3), rng.uniform(0, 8), rng.uniform(0.1, 60), 273.16 - 10 ** rng.uniform(-4, 1)], np.float32)
        w = np.array([rng.uniform(.01, .3), rng.uniform(0, 8), rng.uniform(0.1, 60), 273.16 - 10 ** rng.uniform(-4, 1)], np.float32)
        h = lambda q: (CICE * q[2] + CWAT * q[1]) * (q[3] - TFRZ) + HFUS * q[1]
        hc = h(v.astype(np.float64)) + h(w.astype(np.float64))
        m0 = v[:3].astype(np.float64) + w[:3]
        O.lib().nmo_combo(v.ctypes.data, w.ctypes.data)
        assert np.allclose(v[:3], m0, rtol=1e-6)
        cap = CICE * m0[2] + CWAT * m0[1]
        if hc < 0:
            want = TFRZ + hc / cap; seen.add("cold")
        elif hc <= HFUS * m0[1]:
            want = TFRZ; seen.add("mixed")
        else:
            want = TFRZ + (hc - HFUS * m0[1]) / cap; seen.add("warm")
        assert abs(v[3] - want) < 2e-3, (v[3], want)
    assert {"cold", "mixed"} <= seen


FN = dict(exp=0, log=1, log10=2, pow=3, atan=4, tan=5, cos=6, acos=7, tanh=8)


@pytest.mark.parametrize("name,lo,hi", [("exp", -80, 80), ("log", 1e-30, 1e30), ("log10", 1e-30, 1e30), ("atan", -50, 50),
                                        ("tanh", -12, 12), ("cos", -6.3, 6.3), ("acos", -1, 1), ("tan", -1.5, 1.5)])
def test_portable_math_within_two_ulp_of_libm(O, name, lo, hi):
    """csrc/nmp_math.h (what the GPU parity build and the oracle's M1 mode evaluate) vs glibc."""
    rng = np.random.default_rng(1)
    if lo > 0:
        x = np.exp(rng.uniform(np.log(lo), np.log(hi), 200000)).astype(np.float32)
    else:
        x = rng.uniform(lo, hi, 200000).astype(np.float32)
    O.set_math_mode(1); a = O.math_array(FN[name], x)
    O.set_math_mode(0); b = O.math_array(FN[name], x)
    ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    ok = np.isfinite(b)
    # glibc's own log10f / tanf are 2-ulp functions; the portable versions are the correctly rounded value
    assert ulp[ok].max() <= 2, (name, int(ulp[ok].max()))
    assert (ulp[ok] > 1).mean() < 1e-3 and (ulp[ok] > 0).mean() < 0.25


def test_portable_pow_within_one_ulp_of_libm(O):
    rng = np.random.default_rng(2)
    x = np.exp(rng.uniform(np.log(1e-6), np.log(1e4), 300000)).astype(np.float32)
    y = rng.uniform(-12, 30, 300000).astype(np.float32)
    O.set_math_mode(1); a = O.math_array(FN["pow"], x, y)
    O.set_math_mode(0); b = O.math_array(FN["pow"], x, y)
    ok = np.isfinite(b) & (b > 1e-37)
    ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    assert ulp[ok].max() <= 1


def _cfg(name, ni, nj):
    cfg = S.named_config(name); cfg.ni, cfg.nj = ni, nj
    return cfg


@pytest.mark.parametrize("name,ni,nj,steps", [("C1", 10, 10, 24), ("C2", 116, 56, 72), ("C3", 96, 64, 48), ("C4", 120, 90, 48)])
@pytest.mark.parametrize("math_mode", [0, 1])
def test_conservation_checks_hold(O, tables_usgs, name, ni, nj, steps, math_mode):
    """ERRSW / ERRENG / ERRWAT (noahmplsm.F90:1164-1226, glacier.F90:2932-2970) stay under the reference's fatal
    thresholds on every column-step of every synthetic configuration, and the state stays finite."""
    cfg = _cfg(name, ni, nj)
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state = make_case(cfg, tables_usgs)
    err = run_oracle(cfg, ts, st, state, steps, math_mode=math_mode, nthreads=4)
    assert err is None, err
    for n in ("tsk", "tslb", "smois", "sh2o", "snow", "snowh", "hfx", "lh", "xlaixy"):
        assert np.isfinite(state[n]).all(), n
    land = st["xland"] < 1.5
    assert 200 < state["tsk"][land].min() and state["tsk"][land].max() < 340


def test_oracle_thread_count_invariance(O, tables_usgs):
    """Columns are independent: 1 thread and 5 threads give identical bits."""
    cfg = _cfg("C4", 64, 40)
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, s0 = make_case(cfg, tables_usgs)
    a, b = clone_state(s0), clone_state(s0)
    run_oracle(cfg, ts, st, a, 4, nthreads=1)
    run_oracle(cfg, ts, st, b, 4, nthreads=5)
    assert not diff_report(a, b)


def test_math_mode_sensitivity_defines_tolerances(O, tables_usgs):
    """oracle(host libm) vs oracle(portable math) — two <=1 ulp libms — already differ
    by this much after 24 steps; the FAST-build tolerances of tests/test_parity_gpu.py sit above these numbers."""
    cfg = _cfg("C2", 116, 112)
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, s0 = make_case(cfg, tables_usgs)
    a, b = clone_state(s0), clone_state(s0)
    run_oracle(cfg, ts, st, a, 24, math_mode=0)
    run_oracle(cfg, ts, st, b, 24, math_mode=1)
    land = st["xland"] < 1.5
    d = np.abs(a["tsk"] - b["tsk"])[land]
    assert np.quantile(d, 0.999) < 0.05          # branch flips are rare ...
    assert (a["isnowxy"] != b["isnowxy"]).mean() < 2e-3
    assert np.abs(a["smois"] - b["smois"]).max() < 5e-3


def test_oracle_state_matches_committed_fixture(O, tables_usgs):
    """tests/golden/oracle_state.json (made by tests/golden/gen_oracle_state.py): the portable-math oracle reproduces
    the committed CRC-32 of every state array; a change of the oracle's arithmetic has to be deliberate."""
    import importlib.util
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("gen_oracle_state", os.path.join(here, "gen_oracle_state.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(here, "oracle_state.json")))
    got = mod.state_crcs()
    assert got.keys() == want.keys()
    for case in want:
        bad = [n for n in want[case] if want[case][n] != got[case][n]]
        assert not bad, (case, bad)
