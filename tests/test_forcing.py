"""Row f2 — driver-side forcing preparation (hrldas_input_interpolate, fills, CALC_DECLIN): oracle known answers on
CPU; the device pipeline against the oracle and against the plain noahmplsm path on GPU."""
import numpy as np
import pytest

from noahmp_b200 import _capi, synthetic as S

from helpers import clone_state, diff_report, make_case


def _files(cfg, st, steps):
    """Forcing 'files' at the given model steps, as hrldas_input_read would hold them (VEGFRA as a fraction)."""
    xp = S.backend()
    out = []
    for k in steps:
        f = S.forcing(xp, cfg, k, st)
        out.append({"t": f["t"], "q": f["qv"], "u": f["u"], "v": f["v"], "p": f["p"], "lw": f["glw"], "sw": f["swdown"],
                    "pcp": (f["rainbl"] / np.float32(cfg.dt)).astype(np.float32),
                    "fpar": (st["vegfra"] / np.float32(100.0)).astype(np.float32)})
        for n in out[-1]:
            out[-1][n] = np.ascontiguousarray(out[-1][n], np.float32)
    return out


def test_calc_declin_matches_numpy_restatement(built, tables_usgs):
    """COSZEN / JULIAN of the C++ oracle equal the independent numpy restatement in noahmp_b200/synthetic.py (both
    follow module_hrldas_noahmp_driver.F90:813-863) to fp32 rounding."""
    from oracle import oracle as O
    cfg = S.named_config("C2"); cfg.ni, cfg.nj = 58, 28
    _, st, _ = make_case(cfg, tables_usgs)
    A, B = _files(cfg, st, (1, 4))
    O.set_math_mode(0)
    for step in (1, 7, 14, 20):
        yr, julian, hour = S.clock(cfg, step)
        iday = int(julian)
        out, j = O.forcing(A, B, st["xlatin"], st["xlong"], 1.0, iday, int(hour), 0, 0, cfg.dt)
        cosz, _, jn = S.cosz_julian(S.backend(), cfg, step, st["xlatin"], st["xlong"])
        assert abs(j - float(jn)) < 1e-6
        assert np.abs(out[0] - cosz).max() < 2e-6
        assert np.array_equal(out[1], A["t"]) and np.array_equal(out[9], (A["pcp"] * np.float32(cfg.dt)))
        assert np.array_equal(out[7], out[8]) and (out[11] == 60.0).all()
        assert np.array_equal(out[10], A["fpar"] * np.float32(100.0))


def test_calc_declin_astronomical_known_answers(built, tables_usgs):
    """Where the sun is overhead at local noon the cosine of the zenith angle is 1: on the June solstice at the Tropic
    of Cancer (declination +23.44 deg), on the December solstice at the Tropic of Capricorn, at the equinoxes on the
    equator; and at the equator at local midnight the sun is at the nadir.  The driver's formula for the declination is
    good to a few tenths of a degree, i.e. 1 - cosz < 1e-3."""
    from oracle import oracle as O
    cfg = S.named_config("C1")
    _, st, _ = make_case(cfg, tables_usgs)
    A, B = _files(cfg, st, (1, 4))
    lat = np.zeros_like(st["xlatin"]); lon = np.zeros_like(st["xlong"])
    O.set_math_mode(0)
    cases = [(172, 23.44, 12, 1.0), (355, -23.44, 12, 1.0), (80, 0.0, 12, 1.0), (266, 0.0, 12, 1.0), (80, 0.0, 0, -1.0)]
    for iday, la, hour, want in cases:
        lat[...] = la
        out, _ = O.forcing(A, B, lat, lon, 1.0, iday, hour, 0, 0, cfg.dt)
        assert abs(float(out[0][0, 0]) - want) < 1.5e-3, (iday, la, hour, float(out[0][0, 0]))
    # 90 degrees of longitude = 6 hours of local time: sunrise/sunset geometry at the equinox on the equator
    lat[...] = 0.0; lon[...] = 90.0
    out, _ = O.forcing(A, B, lat, lon, 1.0, 80, 12, 0, 0, cfg.dt)
    assert abs(float(out[0][0, 0])) < 2e-2


def test_interpolation_weights(built, tables_usgs):
    from oracle import oracle as O
    cfg = S.named_config("C1")
    _, st, _ = make_case(cfg, tables_usgs)
    A, B = _files(cfg, st, (1, 4))
    out, _ = O.forcing(A, B, st["xlatin"], st["xlong"], np.float32(2.0 / 3.0), 120, 1, 0, 0, cfg.dt)
    fr = np.float32(2.0 / 3.0)
    want = (A["t"] * fr) + (B["t"] * (np.float32(1.0) - fr))
    assert np.array_equal(out[1], want.astype(np.float32))
    assert np.array_equal(out[9], A["pcp"] * np.float32(cfg.dt))  # precipitation held from the earlier file


@pytest.mark.gpu
def test_device_forcing_pipeline_bitexact(built, tables_usgs):
    """Forcing files every 3 h, hourly steps: device planes == oracle planes bit for bit (PARITY build), and the
    model stepped from them == the model stepped through noahmplsm with host-prepared arrays."""
    import noahmp_b200
    from oracle import oracle as O
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = 96, 72
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state0 = make_case(cfg, tables_usgs)
    files = _files(cfg, st, (1, 4, 7))
    a, b = clone_state(state0), clone_state(state0)
    m1 = noahmp_b200.NoahMP(tables_usgs, cfg.ni, cfg.nj, sync=noahmp_b200.SYNC_FULL, math=noahmp_b200.MATH_PARITY)
    m2 = noahmp_b200.NoahMP(tables_usgs, cfg.ni, cfg.nj, sync=noahmp_b200.SYNC_RESIDENT, math=noahmp_b200.MATH_PARITY)
    m2.set_chunks(3)
    m2.set_fetch(["tsk", "hfx"])
    m2.forcing_static(st["xlatin"], st["xlong"], 30.0)
    xp = S.backend()
    O.set_math_mode(1)
    arr0, sc0 = S.args_from(cfg, st, S.forcing(xp, cfg, 1, st), b, 1)
    m2.upload(arr0, sc0)
    m2.forcing_upload(0, files[0]); m2.forcing_upload(1, files[1])
    bracket = 0
    for step in range(1, 7):
        if step == 4:  # model time reached file B: B becomes A, the next file is read
            m2.forcing_swap(); m2.forcing_upload(1, files[2]); bracket = 1
        k = (step - 1) % 3
        fraction = np.float32(np.float32(3 - k) / np.float32(3))
        yr, julian, hour = S.clock(cfg, step)
        planes, j = O.forcing(files[bracket], files[bracket + 1], st["xlatin"], st["xlong"], fraction, int(julian),
                              int(hour), 0, 0, cfg.dt)
        # path 1: host-prepared arrays through the plain call
        frc = {"coszin": planes[0], "t": planes[1], "qv": planes[2], "u": planes[3], "v": planes[4], "swdown": planes[5],
               "glw": planes[6], "p": planes[7], "rainbl": planes[9], "yr": yr, "julian": j}
        st1 = dict(st); st1["vegfra"] = planes[10]
        arr1, sc = S.args_from(cfg, st1, frc, a, step)
        s1 = m1.noahmplsm(arr1, sc)
        # path 2: device pipeline
        j2 = m2.forcing_apply(float(fraction), int(julian), int(hour), 0, 0, float(cfg.dt))
        assert j2 == j
        import torch
        dev = [torch.as_tensor(p, device="cuda:0").cpu().numpy() for p in m2.device_forcing()]
        for idx in range(12):
            assert np.array_equal(dev[idx].view(np.int32), planes[idx].view(np.int32)), (step, idx)
        arr2, sc2 = S.args_from(cfg, st1, frc, b, step)
        sc2 = dict(sc2); sc2["julian"] = j2
        s2 = m2.noahmplsm_device_forcing(arr2, sc2)
        assert (s1.code, s2.code) == (0, 0)
        assert np.array_equal(a["tsk"], b["tsk"]) and np.array_equal(a["hfx"], b["hfx"]), step
    m2.sync_host(arr2, sc2)
    assert not diff_report(a, b)
    m1.close(); m2.close()
